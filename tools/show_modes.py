"""Operand mode chosen per conv op (needs a GPU: the plan is prepared at bind time)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yolo_infer_pt_b200 import synth
from yolo_infer_pt_b200.nets import nn
size, B = sys.argv[1], int(sys.argv[2])
m = getattr(nn, f"yolo_v11_{size}")(80); synth.load_synth(m, 0, "survey"); m = m.fuse().eval().cuda()
x = torch.zeros(B, 3, 640, 640, dtype=torch.uint8, device="cuda")
m(x)
eng = m._engine_for(x)
for op in eng.describe()["ops"]:
    if op["kind"] == 1 and op["k"] == 3:
        print(f"{op['name']:36s} s{op['stride']} C{op['src'][0]['C']:4d} N{op['N_pad']:4d}/BN{op['BN']:3d} {op['Hout']:3d}x{op['Wout']:<3d} "
              f"patch{op['patch']} pair{op['pair']} res{op['resident']} occ{op['occ']} st{op['stages']} K_pad{op['K_pad']}")
