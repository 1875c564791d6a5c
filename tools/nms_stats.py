"""Work counters of the per-image NMS kernel on the bench workload (library built with NVCC_EXTRA=-DYB_NMS_STATS)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yolo_infer_pt_b200 import synth
from yolo_infer_pt_b200.nets import nn
from yolo_infer_pt_b200.utils import util
m = nn.yolo_v11_n(80); synth.load_synth(m, 0, "survey"); m = m.fuse().eval().cuda()
B = 64
x = (synth.synth_images(8, 640, 640) * 255).round().to(torch.uint8).repeat(B // 8, 1, 1, 1).contiguous().cuda()
y = m(x)
det, cnt = util.nms_padded(y, 0.001, 0.65)
torch.cuda.synchronize()
ws = list(util._workspaces.values())[0]
hdr = ws[:B * 32].view(torch.int32).view(B, 8).cpu().float()
names = ["hist", "walk/select", "compaction scan", "sort", "prefetch+A+compaction", "B", "C", "D"]
for i, n in enumerate(names):
    print(f"{n:22s} mean {hdr[:, i].mean().item():12.0f}")
