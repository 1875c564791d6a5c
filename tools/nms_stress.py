"""BASELINE config 5: NMS-only timing at eval settings (conf 0.001, IoU 0.7, max_det 300), B = 64,
8400 / 33600 anchors x 80 classes, sparse (~8 k candidates / image) and dense (every slot a candidate)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from yolo_infer_pt_b200 import synth
from yolo_infer_pt_b200.utils import util
for anchors, img in ((8400, 640), (33600, 1280)):
    for mode in ("sparse", "dense"):
        pred = torch.from_numpy(synth.synth_predictions(64, 80, anchors, img=img, mode=mode, seed=1)).cuda()
        for _ in range(3): det, cnt = util.nms_padded(pred, 0.001, 0.7)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): det, cnt = util.nms_padded(pred, 0.001, 0.7)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        gb = pred.numel() * 4 / 1e9
        print(f"B=64 A={anchors} {mode:6s}: {ms:7.3f} ms per call = {64 / ms * 1e3:9.0f} img/s, scores read at {gb / ms * 1e3:6.0f} GB/s, "
              f"mean kept {cnt.float().mean().item():.0f}")
