"""Per-layer table: operand mode, tiling, measured time (event between ops) vs the layer's roofline.
usage: layer_table.py <size> <batch> [steps]   (GPU)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from yolo_infer_pt_b200 import synth
from yolo_infer_pt_b200.nets import nn
size, B = sys.argv[1], int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
m = getattr(nn, f"yolo_v11_{size}")(80); synth.load_synth(m, 0, "survey"); m = m.fuse().eval().cuda()
x = torch.randint(0, 255, (B, 3, 640, 640), dtype=torch.uint8, device="cuda")
for _ in range(3): m(x)
eng = m._engine_for(x)
eng.profile(True)
for _ in range(steps): m(x)
torch.cuda.synchronize()
ms, n = eng.profile_read()
eng.profile(False)
desc = eng.describe()
work = bench.conv_algorithmic_work(desc, B)
peaks = bench.read_peaks()
tot = gap = 0.0
print(f"{'name':34s} k s    C    N  BN  HxW mode res occ st   ms   roof  ratio")
for i, op in enumerate(desc["ops"]):
    t = float(ms[i]); tot += t
    if op["kind"] != 1:
        print(f"{op['name']:34s} kind{op['kind']} {'':48s}{t:6.3f}")
        continue
    fl, by = work[i]
    if op.get("fused_away"):
        continue
    roof = max(fl / (peaks["tf_sustained"] * 1e12), by / (peaks["hbm"] * 1e9)) * 1e3
    mode = "DW" if op["dw_fused"] else "P2" if op["pair"] else "P" if op["patch"] else "T" if op["a_tma"] else "G"
    C = sum(s["C"] for s in op["src"])
    gap += t - roof
    print(f"{op['name']:34s} {op['k']} {op['stride']} {C:4d} {op['N_pad']:4d} {op['BN']:3d} {op['Hout']:4d} {mode:>3s} {op['resident']:3d} {op['occ']:3d} {op['stages']:2d} "
          f"{t:6.3f} {roof:6.3f} {t / max(roof, 1e-9):5.2f}")
print(f"total {tot:.3f} ms, conv gap to per-layer roofline {gap:.3f} ms")
