"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's eval-time image
pre-processing, the step immediately in front of YOLO.forward.

Follows, operation by operation:
  * utils/dataset.py:95-103  `Dataset.load_image`: r = input_size / max(h, w); if r != 1:
        cv2.resize(image, dsize=(int(w * r), int(h * r)), interpolation=cv2.INTER_LINEAR)
  * utils/dataset.py:292-313 `resize(image, input_size, augment=False)`: letterbox to a square of
        input_size with cv2.copyMakeBorder(..., cv2.BORDER_CONSTANT) (value 0), borders
        top/bottom = round(h -/+ 0.1), left/right = round(w -/+ 0.1)
  * utils/dataset.py:86-88   HWC -> CHW, BGR -> RGB
The bilinear resampler is OpenCV's 8-bit INTER_LINEAR (third-party dependency, not under /root/reference;
de-facto version: opencv 4.13.0 as installed): 11-bit fixed-point coefficients from float32 fractions,
horizontal pass in int32, vertical pass ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2; the
x direction clamps the fraction at the image edge, the y direction only clamps the row index.
Pinned in tests/test_oracle.py against cv2 itself (53+ shapes, bit-exact) and against fixtures recorded
from the reference's own functions (tests/golden/letterbox_*.npz).
"""
import numpy as np


def _coeffs(dn, sn, clamp_frac):
    inv = np.float64(dn) / np.float64(sn)
    scale = np.float64(1.0) / inv
    d = np.arange(dn, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp_frac:
        lo = s < 0
        f[lo] = 0
        s[lo] = 0
        hi = s >= sn - 1
        f[hi] = 0
        s[hi] = sn - 1
    a1 = np.rint(f * np.float32(2048)).astype(np.int32)
    a0 = np.rint((np.float32(1.0) - f) * np.float32(2048)).astype(np.int32)
    return s, a0, a1


def resize_linear_u8(src, dw, dh):
    """cv2.resize(src, dsize=(dw, dh), interpolation=cv2.INTER_LINEAR) for uint8 HWC images."""
    sh, sw = src.shape[:2]
    xo, xa0, xa1 = _coeffs(dw, sw, True)
    yo, ya0, ya1 = _coeffs(dh, sh, False)
    s = src.astype(np.int32)
    x1 = np.minimum(xo + 1, sw - 1)
    rows = s[:, xo] * xa0[None, :, None] + s[:, x1] * xa1[None, :, None]
    y0 = np.clip(yo, 0, sh - 1)
    y1 = np.clip(yo + 1, 0, sh - 1)
    b0 = ya0[:, None, None]
    b1 = ya1[:, None, None]
    out = (((b0 * (rows[y0] >> 4)) >> 16) + ((b1 * (rows[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def geometry(h, w, input_size):
    """(dh, dw, top, left, ratio, pad_w, pad_h) exactly as load_image + resize compute them."""
    r = input_size / max(h, w)
    dh, dw = (int(h * r), int(w * r)) if r != 1 else (h, w)
    r2 = min(min(input_size / dh, input_size / dw), 1.0)
    pad = int(round(dw * r2)), int(round(dh * r2))
    assert pad == (dw, dh), "second resize of dataset.resize() is unreachable after load_image"
    pw = (input_size - pad[0]) / 2
    ph = (input_size - pad[1]) / 2
    top, left = int(round(ph - 0.1)), int(round(pw - 0.1))
    return dh, dw, top, left, r, pw, ph


def letterbox(image_bgr, input_size):
    """HWC uint8 BGR image -> (3, S, S) uint8 RGB letterboxed sample, (ratio, pad_w, pad_h)."""
    h, w = image_bgr.shape[:2]
    dh, dw, top, left, r, pw, ph = geometry(h, w, input_size)
    img = image_bgr if (dh, dw) == (h, w) else resize_linear_u8(image_bgr, dw, dh)
    out = np.zeros((input_size, input_size, 3), dtype=np.uint8)
    out[top:top + dh, left:left + dw] = img
    return np.ascontiguousarray(out.transpose(2, 0, 1)[::-1]), (r, pw, ph)
