"""NMS-only timing on the bench workload's predictions (YOLO11n, survey weights), B from argv."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yolo_infer_pt_b200 import synth
from yolo_infer_pt_b200.nets import nn
from yolo_infer_pt_b200.utils import util
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
m = nn.yolo_v11_n(80); synth.load_synth(m, 0, "survey"); m = m.fuse().eval().cuda()
x = (synth.synth_images(8, 640, 640) * 255).round().to(torch.uint8).repeat((B + 7) // 8, 1, 1, 1)[:B].contiguous().cuda()
y = m(x).clone()
for _ in range(3): util.nms_padded(y, 0.001, 0.65)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): util.nms_padded(y, 0.001, 0.65)
e1.record(); torch.cuda.synchronize()
print(f"B={B} first={os.environ.get('YB_NMS_FIRST_BAND','-')} next={os.environ.get('YB_NMS_NEXT_BAND','-')}: {e0.elapsed_time(e1)/20*1e3:.1f} us per NMS call")
