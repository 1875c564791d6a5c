"""Batch-1 latency (BASELINE config 3): p50/p90 of forward(+NMS) for YOLO11 n/s/m at 640x640, CUDA graph replay."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from yolo_infer_pt_b200 import synth
from yolo_infer_pt_b200.nets import nn
from yolo_infer_pt_b200.utils import util
ap = argparse.ArgumentParser(); ap.add_argument("--models", default="n,s,m"); ap.add_argument("--iters", type=int, default=500)
a = ap.parse_args()
out = {}
for size in a.models.split(","):
    m = getattr(nn, f"yolo_v11_{size}")(80); synth.load_synth(m, 0, "survey"); m = m.fuse().eval().cuda()
    x = (synth.synth_images(1, 640, 640) * 255).round().to(torch.uint8).cuda()
    for _ in range(20): y = m(x); util.nms_padded(y)
    torch.cuda.synchronize()
    res = {}
    for name, fn in (("forward", lambda: m(x)), ("forward+nms", lambda: util.nms_padded(m(x)))):
        ts = []
        for _ in range(a.iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
        ws = []
        for _ in range(a.iters):
            t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ws.append((time.perf_counter() - t0) * 1e3)
        res[name] = dict(p50_ms=float(np.percentile(ts, 50)), p90_ms=float(np.percentile(ts, 90)),
                         wall_p50_ms=float(np.percentile(ws, 50)))
    out[size] = res
print(json.dumps(out))
