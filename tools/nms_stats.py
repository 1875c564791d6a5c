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
# header words in a stats build: cand_count, sel_count, sel2_count, selected, pad[0..] (struct ImgHdr order may differ: print all)
names = ["cycles: histogram", "cycles: select + compact", "bands", "chunks of 128", "cycles: load + test vs kept + compaction",
         "cycles: B (pairwise masks)", "cycles: C (resolution)", "cycles: D (append)"]
print("raw header means:", [round(hdr[:, i].mean().item()) for i in range(8)])
