"""Minimal forward for ncu: YOLO11<model> at BxSxS, `--iters` forwards (first one is warm-up)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yolo_infer_pt_b200 import synth
from yolo_infer_pt_b200.nets import nn
from yolo_infer_pt_b200.utils import util
ap = argparse.ArgumentParser()
ap.add_argument("--model", default="n"); ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--size", type=int, default=640); ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--nms", type=int, default=0)
a = ap.parse_args()
m = getattr(nn, f"yolo_v11_{a.model}")(80)
synth.load_synth(m, 0, "survey")
m = m.fuse().eval().cuda()
x = (synth.synth_images(min(a.batch, 4), a.size, a.size) * 255).round().to(torch.uint8)
x = x.repeat((a.batch + 3) // 4, 1, 1, 1)[:a.batch].contiguous().cuda()
for _ in range(a.iters):
    y = m(x)
    if a.nms:
        util.nms_padded(y, 0.001, 0.65)
torch.cuda.synchronize()
print("ok", tuple(y.shape))
