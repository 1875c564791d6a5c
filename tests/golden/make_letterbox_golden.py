"""Records outputs of the reference's own pre-processing functions (utils/dataset.py load_image logic +
resize) on small seeded images -> tests/golden/letterbox_cases.npz.  Run in the build container (needs
/root/reference and cv2)."""
import os, sys
import numpy as np
import cv2
sys.path.insert(0, "/root/reference")
from utils import dataset as ref_ds  # noqa: E402

S = 96
rng = np.random.default_rng(7)
shapes = [(60, 96), (96, 60), (37, 50), (200, 120), (96, 96), (11, 7), (150, 301)]
out = {"input_size": S, "n": len(shapes)}
for i, (h, w) in enumerate(shapes):
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    # Dataset.load_image (dataset.py:95-103), augment = False
    r = S / max(h, w)
    im = img
    if r != 1:
        im = cv2.resize(img, dsize=(int(w * r), int(h * r)), interpolation=cv2.INTER_LINEAR)
    im, ratio, pad = ref_ds.resize(im, S, False)                     # dataset.py:292-313
    sample = np.ascontiguousarray(im.transpose((2, 0, 1))[::-1])     # dataset.py:86-88
    out[f"img{i}"] = img
    out[f"out{i}"] = sample
    out[f"meta{i}"] = np.array([ratio[0] * r, pad[0], pad[1]], dtype=np.float64)
np.savez_compressed(os.path.join(os.path.dirname(__file__), "letterbox_cases.npz"), **out)
print("wrote letterbox_cases.npz")
