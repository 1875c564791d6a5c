"""Determinism probe: a batch made of 8 pictures repeated must give identical predictions for identical pictures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yolo_infer_pt_b200 import synth
from yolo_infer_pt_b200.nets import nn
recipe = sys.argv[1] if len(sys.argv) > 1 else "survey"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
m = nn.yolo_v11_n(80); synth.load_synth(m, 0, recipe); m = m.fuse().eval().cuda()
base = (synth.synth_images(8, 640, 640, seed=0) * 255).round().to(torch.uint8)
x = base.repeat(B // 8, 1, 1, 1).contiguous().cuda()
for trial in range(3):
    with torch.no_grad():
        y = m(x)
    torch.cuda.synchronize()
    ref = y[:8]
    bad = []
    for i in range(8, B):
        d = (y[i] - ref[i % 8]).abs()
        if d.max() > 0:
            rows = (d.amax(1) > 0).nonzero().flatten().tolist()
            cols = (d.amax(0) > 0).nonzero().flatten()
            bad.append((i, float(d.max()), len(rows), rows[:6], int(cols.min()), int(cols.max()), int(cols.numel())))
    print(f"trial {trial}: {len(bad)} of {B - 8} images differ from their twin", bad[:6])
