"""GPU parity of the device-side compute_metric (yb_compute_metric) against the oracle and the reference
fixtures, exact (boolean matrices), through the C ABI."""
import os

import numpy as np
import pytest
import torch

from oracle import metric_oracle as mo
from yolo_infer_pt_b200.utils import util

pytestmark = pytest.mark.gpu


def test_metric_matches_reference_fixtures(golden_dir):
    g = np.load(os.path.join(golden_dir, "metric_cases.npz"))
    iou_v = torch.from_numpy(g["iou_v"]).cuda()
    for i in range(int(g["n"])):
        c = util.compute_metric(torch.from_numpy(g[f"det{i}"]).cuda(), torch.from_numpy(g[f"gt{i}"]).cuda(), iou_v)
        assert c.dtype == torch.bool and c.is_cuda
        assert np.array_equal(c.cpu().numpy(), g[f"correct{i}"]), f"case {i}"


def test_metric_batch_vs_oracle_random():
    rng = np.random.default_rng(5)
    B, max_det, max_t, nc = 16, 300, 64, 6
    iou_v = np.linspace(0.5, 0.95, 10, dtype=np.float32)
    det = np.zeros((B, max_det, 6), np.float32)
    tgt = np.zeros((B, max_t, 5), np.float32)
    counts = rng.integers(0, max_det + 1, B).astype(np.int32)
    tcounts = rng.integers(0, max_t + 1, B).astype(np.int32)
    counts[0], tcounts[1] = 0, 0            # no detections / no labels
    for b in range(B):
        m, n = int(tcounts[b]), int(counts[b])
        c = rng.uniform(60, 580, (m, 2)); s = rng.uniform(20, 200, (m, 2))
        tgt[b, :m, 0] = rng.integers(0, nc, m)
        tgt[b, :m, 1:3] = c - s / 2
        tgt[b, :m, 3:5] = c + s / 2
        if n and m:
            src = rng.integers(0, m, n)
            det[b, :n, :4] = tgt[b, src, 1:5] + rng.normal(0, 1, (n, 4)) * rng.uniform(1, 25, (n, 1))
            det[b, :n, 5] = np.where(rng.random(n) < 0.85, tgt[b, src, 0], rng.integers(0, nc, n))
        elif n:
            det[b, :n, :4] = rng.uniform(0, 640, (n, 4))
        det[b, :n, 4] = np.sort(rng.uniform(0.001, 1, n))[::-1]
    got = util.compute_metric_batch(torch.from_numpy(det).cuda(), torch.from_numpy(counts).cuda(),
                                    torch.from_numpy(tgt).cuda(), torch.from_numpy(tcounts).cuda(),
                                    torch.from_numpy(iou_v).cuda()).cpu().numpy()
    for b in range(B):
        n, m = int(counts[b]), int(tcounts[b])
        want = mo.compute_metric(det[b, :n], tgt[b, :m], iou_v) if m else np.zeros((n, 10), bool)
        assert np.array_equal(got[b, :n], want), f"image {b}"
        assert not got[b, n:].any()


def test_metric_end_to_end_with_nms():
    """nms_padded -> compute_metric_batch without leaving the device."""
    from yolo_infer_pt_b200 import synth
    pred = torch.from_numpy(synth.synth_predictions(4, 80, 2100, img=320, mode="sparse", seed=9)).cuda()
    det, counts = util.nms_padded(pred, 0.001, 0.65)
    tgt = torch.zeros((4, 8, 5), device="cuda")
    tgt[:, :, 0] = det[:, :8, 5]
    tgt[:, :, 1:5] = det[:, :8, :4]         # labels = the top detections themselves
    tc = torch.full((4,), 8, dtype=torch.int32, device="cuda")
    correct = util.compute_metric_batch(det, counts, tgt, tc, torch.linspace(0.5, 0.95, 10))
    assert correct[:, :8].all()             # every label is matched by its own detection at every threshold
