"""Forward time per batch size with one stream lane and with four (YB_LANES): python tools/lanes_vs_batch.py [model]"""
import os, sys, json, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 2:   # child: lanes batch model
    import numpy as np, torch
    from yolo_infer_pt_b200 import synth
    from yolo_infer_pt_b200.nets import nn
    b, size = int(sys.argv[2]), sys.argv[3]
    m = getattr(nn, f"yolo_v11_{size}")(80); synth.load_synth(m, 0, "survey"); m = m.fuse().eval().cuda()
    x = (synth.synth_images(min(b, 4), 640, 640) * 255).round().to(torch.uint8).repeat((b + 3) // 4, 1, 1, 1)[:b].contiguous().cuda()
    for _ in range(10): m(x)
    torch.cuda.synchronize()
    ts = []
    for _ in range(100):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); m(x); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    print(float(np.percentile(ts, 50)))
    sys.exit(0)
size = sys.argv[1] if len(sys.argv) > 1 else "n"
for b in (1, 2, 4, 8, 16, 32, 64):
    r = {}
    for lanes in (1, 4):
        env = dict(os.environ, YB_LANES=str(lanes))
        r[lanes] = float(subprocess.run([sys.executable, __file__, str(lanes), str(b), size], env=env, capture_output=True, text=True).stdout.strip().splitlines()[-1])
    print(f"{size} B={b}: 1 lane {r[1]:.3f} ms, 4 lanes {r[4]:.3f} ms ({100 * (r[4] / r[1] - 1):+.1f} %)", flush=True)
