"""GPU parity tests of non_max_suppression: bit-exact against the reference's recorded outputs
(tests/golden) and against the C oracle, through the C ABI (yb_nms)."""
import os

import numpy as np
import pytest
import torch

from oracle import nms_oracle
from yolo_infer_pt_b200 import synth
from yolo_infer_pt_b200.utils import util
from test_oracle import NMS_CASES, load_nms_case

pytestmark = pytest.mark.gpu


def _run(pred, conf, iou):
    out = util.non_max_suppression(torch.from_numpy(pred).to("cuda:0"), conf, iou)
    return [o.cpu().numpy() for o in out]


@pytest.mark.parametrize("name", NMS_CASES)
def test_nms_bit_exact_vs_reference_golden(name, golden_dir):
    g, pred = load_nms_case(golden_dir, name)
    out = _run(pred, float(g["conf"]), float(g["iou"]))
    assert [len(o) for o in out] == g["counts"].tolist()
    for b, o in enumerate(out):
        assert o.dtype == np.float32 and o.shape[1] == 6
        assert np.array_equal(o, g["det"][b, :len(o)]), f"image {b}: rows differ from the reference"


@pytest.mark.parametrize("batch,anchors,img,mode,seed", [(64, 8400, 640, "sparse", 20), (4, 8400, 640, "dense", 21),
                                                        (8, 33600, 1280, "sparse", 22), (2, 33600, 1280, "dense", 23),
                                                        (5, 84, 64, "dense", 24)])
def test_nms_bit_exact_vs_oracle_stress(batch, anchors, img, mode, seed):
    """BASELINE config 5: conf 0.001, IoU 0.7, max_det 300, up to 33600 anchors x 80 classes."""
    pred = synth.synth_predictions(batch, 80, anchors, img=img, mode=mode, seed=seed)
    out = _run(pred, 0.001, 0.7)
    ref = nms_oracle.non_max_suppression(pred, 0.001, 0.7)
    for b in range(batch):
        assert np.array_equal(out[b], ref[b]), f"image {b}"


def test_nms_ties_break_by_candidate_index():
    """Documented contract for tied scores: ascending (anchor, class) index — same as the oracle."""
    pred = synth.synth_predictions(2, 80, 2100, img=320, mode="sparse", seed=30)
    s = pred[:, 4:]
    s[s > 0] = np.round(s[s > 0] * 64) / 64 + np.float32(1 / 128)   # heavy ties
    out = _run(pred, 0.001, 0.65)
    ref = nms_oracle.non_max_suppression(pred, 0.001, 0.65)
    for a, b in zip(out, ref):
        assert np.array_equal(a, b)


def _clustered(rng, B, nc, A, img, centres):
    pred = np.zeros((B, 4 + nc, A), dtype=np.float32)
    c = rng.uniform(60, img - 60, size=(B, centres, 2)).astype(np.float32)
    pick = rng.integers(0, centres, size=(B, A))
    for b in range(B):
        pred[b, 0] = c[b, pick[b], 0] + rng.normal(0, 6, A)
        pred[b, 1] = c[b, pick[b], 1] + rng.normal(0, 6, A)
    pred[:, 2] = rng.uniform(90, 110, size=(B, A))
    pred[:, 3] = rng.uniform(90, 110, size=(B, A))
    return pred


def test_nms_fat_histogram_bins_and_many_bands():
    """Candidates share two score values (score bins with far more than 4096 keys: the band select has
    to split them exactly on the index bits) and the boxes are clustered, so most are suppressed and
    the lazy walk runs through many bands.  Image 0 fits the key list, image 1 (> 32768 candidates)
    is banded from the raw scores and hits the max_nms = 30000 cut (util.py:157)."""
    rng = np.random.default_rng(50)
    B, nc, A = 2, 80, 8400
    pred = _clustered(rng, B, nc, A, 640, 12)
    for b, ncls in enumerate((3, 6)):          # few classes -> far fewer than max_det survivors
        on = rng.random((ncls, A)) < 0.8
        pred[b, 4:4 + ncls] = np.where(on, rng.choice(np.float32([0.5, 0.25]), size=(ncls, A)), 0)
    out = _run(pred, 0.001, 0.65)
    ref = nms_oracle.non_max_suppression(pred, 0.001, 0.65)
    for b in range(B):
        assert len(ref[b]) < 300
        assert np.array_equal(out[b], ref[b]), f"image {b}"


def test_nms_many_bands_distinct_scores():
    """Clustered boxes with distinct scores: fewer than max_det survive, every candidate is consumed."""
    rng = np.random.default_rng(51)
    B, nc, A = 2, 80, 8400
    pred = _clustered(rng, B, nc, A, 640, 6)
    n = 3 * A
    ladder = (rng.permutation(n).astype(np.float32) + 1) / np.float32(n + 2)      # tie-free
    on = rng.random(n) < 0.8
    pred[:, 4:7] = np.where(on, ladder, 0).reshape(3, A).astype(np.float32)
    out = _run(pred, 0.001, 0.65)
    ref = nms_oracle.non_max_suppression(pred, 0.001, 0.65)
    for b in range(B):
        assert len(ref[b]) < 300
        assert np.array_equal(out[b], ref[b]), f"image {b}"


def test_nms_scores_outside_unit_interval():
    """Negative threshold, scores on both sides of [0, 1] (clamped histogram bins at both ends)."""
    rng = np.random.default_rng(52)
    pred = synth.synth_predictions(1, 80, 2100, img=320, mode="sparse", seed=53)
    n = 80 * 2100
    vals = np.concatenate([np.linspace(-0.9, 3.0, n // 2, dtype=np.float32),
                           -np.geomspace(1e-30, 0.5, n - n // 2).astype(np.float32)])
    vals = np.unique(vals)
    sc = np.full(n, -5.0, dtype=np.float32)
    sc[:len(vals)] = vals
    pred[0, 4:] = rng.permutation(sc).reshape(80, 2100)
    out = _run(pred, -1.0, 0.65)
    ref = nms_oracle.non_max_suppression(pred, -1.0, 0.65)
    assert np.array_equal(out[0], ref[0])


def test_nms_edge_cases():
    # threshold compared in double: fl32(1/3) > 1/3 suppresses
    pred = np.zeros((1, 5, 2), dtype=np.float32)
    pred[0, :, 0] = [1.0, 0.5, 2.0, 1.0, 0.9]
    pred[0, :, 1] = [2.0, 0.5, 2.0, 1.0, 0.8]
    assert len(_run(pred, 0.1, 1.0 / 3.0)[0]) == 1
    assert len(_run(pred, 0.1, float(np.float32(1.0 / 3.0)))[0]) == 2
    # zero-area identical boxes: IoU is NaN, both kept
    pred = np.zeros((1, 5, 2), dtype=np.float32)
    pred[0, :, 0] = [5.0, 5.0, 0.0, 0.0, 0.9]
    pred[0, :, 1] = [5.0, 5.0, 0.0, 0.0, 0.8]
    assert len(_run(pred, 0.1, 0.65)[0]) == 2
    # NaN scores are never candidates; empty result is (0, 6) fp32
    pred = np.full((2, 84, 84), np.nan, dtype=np.float32)
    out = _run(pred, 0.001, 0.65)
    assert all(o.shape == (0, 6) for o in out)


def test_nms_properties_at_full_size():
    """Size-independent properties on the B=64 stress tensor: sorted scores, <= 300 rows, every kept
    pair of one class has IoU <= thr, and re-running NMS on the kept boxes keeps all of them."""
    pred = synth.synth_predictions(64, 80, 8400, img=640, mode="sparse", seed=40)
    out = _run(pred, 0.001, 0.7)
    for det in out:
        assert len(det) <= 300
        assert np.all(np.diff(det[:, 4]) <= 0)
        for c in np.unique(det[:, 5]):
            d = det[det[:, 5] == c]
            if len(d) < 2:
                continue
            x1 = np.maximum(d[:, None, 0], d[None, :, 0]); y1 = np.maximum(d[:, None, 1], d[None, :, 1])
            x2 = np.minimum(d[:, None, 2], d[None, :, 2]); y2 = np.minimum(d[:, None, 3], d[None, :, 3])
            inter = np.clip(x2 - x1, 0, None) * np.clip(y2 - y1, 0, None)
            area = (d[:, 2] - d[:, 0]) * (d[:, 3] - d[:, 1])
            iou = inter / (area[:, None] + area[None, :] - inter)
            np.fill_diagonal(iou, 0)
            assert iou.max() <= 0.7 + 1e-4
