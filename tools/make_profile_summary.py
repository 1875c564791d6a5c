"""Turns the raw ncu outputs of tools/gpu_profiles.sh (gpurun_out/<tag>_*) into the committed evidence under profiles/."""
import collections, csv, json, os, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = "gpurun_out", "profiles"
os.makedirs(P, exist_ok=True)
U_B = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
U_T = {'ns': 1e-3, 'us': 1, 'ms': 1e3, 'usecond': 1, 'nsecond': 1e-3, 'msecond': 1e3, 'second': 1e6}

def long_csv(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
    per = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if not r or not r[0].isdigit(): continue
        d = per.setdefault(int(r[0]), {"kernel": r[ix['Kernel Name']], "grid": r[ix['Grid Size']], "block": r[ix['Block Size']]})
        d[r[ix['Metric Name']]] = (float(r[ix['Metric Value']].replace(',', '')), r[ix['Metric Unit']])
    return per

def short(name):
    n = name.replace('void ', '').replace('yb::', '')
    return n.split('(')[0]

# ---- (1) launch list of `python bench.py --steps 2 --warmup 1`
per = long_csv(f"{G}/{tag}_launches.csv")
with open(f"{P}/{tag}_launches_yolo11n_b256.csv", "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none: python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras\n")
    f.write("id,kernel,grid,block,duration_us\n")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, d in per.items():
        t = d['gpu__time_duration.sum']; us = t[0] * U_T[t[1]]
        f.write(f"{k},{short(d['kernel'])},\"{d['grid']}\",\"{d['block']}\",{us:.2f}\n")
        a = agg[short(d['kernel'])]; a[0] += 1; a[1] += us
tot = sum(v[1] for v in agg.values())
with open(f"{P}/{tag}_launch_share_yolo11n_b256.txt", "w") as f:
    f.write(f"# share of the summed kernel time by kernel ({len(per)} launches captured; cold-cache serialised ncu replays)\n")
    for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{n:60s} {c:4d} launches {us:10.1f} us {100 * us / tot:5.1f} %\n")

# ---- (2) per-launch DRAM / L2 bytes of every conv launch of one forward
per = long_csv(f"{G}/{tag}_conv_dram.csv")
rows = []
for k, d in per.items():
    g = lambda m, U: d[m][0] * U[d[m][1]]
    rows.append(dict(launch=len(rows), kernel=short(d['kernel']), grid=d['grid'],
                     us=g('gpu__time_duration.sum', U_T), dram_rd=g('dram__bytes_read.sum', U_B),
                     dram_wr=g('dram__bytes_write.sum', U_B), l2=g('lts__t_bytes.sum', U_B),
                     tc_pct=d['sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active'][0]))
with open(f"{P}/{tag}_conv_dram_yolo11n_b256.csv", "w") as f:
    f.write("# every conv_gemm_tcgen05_kernel launch of one YOLO11n B=256 forward (ncu, --clock-control none)\n")
    f.write("launch,kernel,grid,duration_us,dram_read_MB,dram_write_MB,l2_MB,dram_GBps,tensor_pipe_active_pct\n")
    for r in rows:
        f.write(f"{r['launch']},{r['kernel']},\"{r['grid']}\",{r['us']:.2f},{r['dram_rd'] / 1e6:.1f},{r['dram_wr'] / 1e6:.1f},"
                f"{r['l2'] / 1e6:.1f},{(r['dram_rd'] + r['dram_wr']) / r['us'] / 1e3:.0f},{r['tc_pct']:.1f}\n")
traffic = dict(model="n", batch=256, size=640, kernel="conv_gemm_tcgen05_kernel", launches=len(rows),
               dram_bytes_per_step=sum(r['dram_rd'] + r['dram_wr'] for r in rows),
               dram_read_bytes=sum(r['dram_rd'] for r in rows), dram_write_bytes=sum(r['dram_wr'] for r in rows),
               l2_bytes_per_step=sum(r['l2'] for r in rows), ncu_time_us=sum(r['us'] for r in rows),
               source=f"profiles/{tag}_conv_dram_yolo11n_b256.csv")
json.dump(traffic, open(f"{P}/{tag}_traffic.json", "w"), indent=1)

# ---- (3) whole step: DRAM bytes by kernel
if os.path.exists(f"{G}/{tag}_all_dram.csv"):
    per = long_csv(f"{G}/{tag}_all_dram.csv")
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for d in per.values():
        a = agg[short(d['kernel'])]
        a[0] += 1
        a[1] += d['gpu__time_duration.sum'][0] * U_T[d['gpu__time_duration.sum'][1]]
        a[2] += sum(d[m][0] * U_B[d[m][1]] for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
    with open(f"{P}/{tag}_step_dram_by_kernel_yolo11n_b256.txt", "w") as f:
        f.write("# one forward + NMS at B=256: launches, summed duration, DRAM bytes, achieved DRAM GB/s (ncu, cold)\n")
        for n, (c, us, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{n:40s} {c:4d} launches {us:9.1f} us {by / 1e6:10.1f} MB {by / us / 1e3:7.0f} GB/s\n")

# ---- (4) full captures: keep the raw metric page (one row per launch) and a stall summary
KEEP = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct', 'sm__inst_executed.sum', 'sm__cycles_elapsed.max',
        'sm__inst_executed_pipe_tma.sum.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active']
names = {0: "net.p2.0 (3x3 s2 16->32 @160x160, im2col gather)" if tag in ("r01", "r02") else "net.p2.0 (3x3 s2 16->32 @160x160, halo patches over the space-to-depth stem output)",
         3: "net.p2.1.res_m.0.conv2 (3x3 8->16 +res @160x160, halo patch)" if tag in ("r01", "r02") else "net.p2.1.res_m.0.conv2 (3x3 8->16 @160x160, halo patch, residual folded into the consumer's weights)",
         61: "head.box.0.0 (3x3 64->64 @80x80, halo patch, resident weights)", 64: "head.cls.0.1 (dw3x3 + 1x1 64->80 @80x80, fused depthwise)",
         66: "head.cls.0.4 (1x1 80->80 + sigmoid, fp32 planes)"}
with open(f"{P}/{tag}_ncu_full_conv_gemm_yolo11n_b256.csv", "w") as f:
    w = csv.writer(f)
    first = True
    for idx in sorted(names):
        path = f"{G}/{tag}_conv{idx}_raw.csv"
        if not os.path.exists(path): continue
        rows = list(csv.reader(open(path)))
        hdr, units, d = rows[0], rows[1], rows[2]
        ix = {h: i for i, h in enumerate(hdr)}
        cols = [k for k in KEEP if k in ix]
        if first:
            w.writerow(["layer"] + cols); w.writerow(["(unit)"] + [units[ix[k]] for k in cols]); first = False
        w.writerow([names[idx]] + [d[ix[k]] for k in cols])
with open(f"{P}/{tag}_ncu_sass_evidence.txt", "w") as f:
    for idx in sorted(names):
        path = f"{G}/{tag}_conv{idx}_src.csv"
        if not os.path.exists(path): continue
        rows = list(csv.reader(open(path)))
        hi = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'][0]
        hdr = rows[hi + 1]; data = rows[hi + 2:]
        ix = {h: i for i, h in enumerate(hdr)}
        ops = collections.Counter()
        for d in data:
            src = d[ix['Source']].split()
            op = src[1] if src and src[0].startswith('@') and len(src) > 1 else (src[0] if src else '')
            base = op.split('.')[0]
            if base in ('UTCHMMA', 'UTMALDG', 'UTMASTG', 'LDTM', 'UTCBAR', 'LDGSTS', 'SYNCS', 'ARRIVES', 'MUFU', 'FFMA2', 'HFMA2', 'UBLKCP'):
                ops[base] += int(float(d[ix['Instructions Executed']] or 0))
        f.write(f"{names[idx]}\n    executed warp-instructions by class: " + ", ".join(f"{k}={v}" for k, v in sorted(ops.items())) + "\n")
print("wrote", sorted(os.listdir(P)))
