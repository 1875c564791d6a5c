"""Records outputs of the reference's own `utils.util.compute_metric` on seeded detections / labels
-> tests/golden/metric_cases.npz.  Run in the build container (needs /root/reference)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, "/root/reference")
from utils import util as ref_util  # noqa: E402

rng = np.random.default_rng(11)
iou_v = torch.linspace(0.5, 0.95, 10)
out = {"iou_v": iou_v.numpy(), "n": 0}
cases = [(40, 6, 3), (300, 25, 5), (120, 60, 2), (7, 1, 1), (300, 100, 80), (5, 12, 4)]
for ci, (n, m, nc) in enumerate(cases):
    gt = np.zeros((m, 5), np.float32)
    gt[:, 0] = rng.integers(0, nc, m)
    c = rng.uniform(60, 580, (m, 2)); s = rng.uniform(20, 200, (m, 2))
    gt[:, 1:3] = c - s / 2; gt[:, 3:5] = c + s / 2
    det = np.zeros((n, 6), np.float32)
    src = rng.integers(0, m, n)                      # detections are jittered copies of labels
    jit = rng.normal(0, 1, (n, 4)) * rng.uniform(1, 25, (n, 1))
    det[:, :4] = gt[src, 1:5] + jit
    det[:, 4] = np.sort(rng.uniform(0.001, 1, n))[::-1]
    det[:, 5] = np.where(rng.random(n) < 0.85, gt[src, 0], rng.integers(0, nc, n))
    correct = ref_util.compute_metric(torch.from_numpy(det), torch.from_numpy(gt), iou_v).numpy()
    out[f"det{ci}"] = det; out[f"gt{ci}"] = gt; out[f"correct{ci}"] = correct
    out["n"] = ci + 1
np.savez_compressed(os.path.join(os.path.dirname(__file__), "metric_cases.npz"), **out)
print("wrote metric_cases.npz", [int(out[f"correct{i}"].sum()) for i in range(out["n"])])
