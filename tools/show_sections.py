"""Per-launch table from an ncu `--page raw --csv` section capture (tools/ncu_sections.sh)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]; data = rows[hi + 2:]
ix = {h: i for i, h in enumerate(hdr)}
def g(r, k):
    try: return float(r[ix[k]].replace(',', ''))
    except Exception: return float('nan')
cols = [('gpu__time_duration.sum', 'us'), ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2%'), ('l1tex__throughput.avg.pct_of_peak_sustained_active', 'l1%'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%'), ('dram__bytes_read.sum', 'rdMB'), ('dram__bytes_write.sum', 'wrMB'),
        ('lts__t_sector_hit_rate.pct', 'l2hit'), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'),
        ('launch__grid_size', 'grid'), ('launch__registers_per_thread', 'regs')]
flt = sys.argv[2] if len(sys.argv) > 2 else ''
print('  #  ' + 'kernel'.ljust(28) + ' '.join(n.rjust(8) for _, n in cols))
for j, r in enumerate(data):
    name = r[ix['Kernel Name']]
    if flt and flt not in name: continue
    short = name.split('(')[0].replace('yb::', '')[:27]
    vals = []
    for k, n in cols:
        v = g(r, k)
        u = rows[hi + 1][ix[k]] if k in ix else ''
        if n in ('rdMB', 'wrMB'):
            v = v / 1e6 if u == 'byte' else v * (1e3 if u == 'Gbyte' else 1 if u == 'Mbyte' else 1e-3 if u == 'Kbyte' else 1)
        if n == 'us': v = v / 1e3 if u == 'ns' else v * (1e3 if u == 'ms' else 1)
        vals.append(f'{v:8.1f}')
    print(f'{j:3d}  ' + short.ljust(28) + ' '.join(vals))
