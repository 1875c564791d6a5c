"""Device-side stand-in for the eval-time image pre-processing of the reference's `utils.dataset`
(t0saki/YOLO-Infer-pt utils/dataset.py): `Dataset.load_image` (:95-103, cv2.resize INTER_LINEAR),
`resize` (:292-313, letterbox with a constant-0 border) and the HWC->CHW / BGR->RGB shuffle of
`__getitem__` (:86-88), for a whole batch in one kernel of libyolob200.so (`yb_letterbox`), bit-exact
with OpenCV's 8-bit bilinear resampler.  The output is the (B, 3, S, S) uint8 tensor `YOLO.forward`
takes directly (the stem kernel folds main.py:266-267's `/ 255`).  No CPU fallback: inputs must be CUDA
tensors.  Training-time augmentation (mosaic, HSV, flips) is out of scope.
"""
import ctypes

import torch

from yolo_infer_pt_b200 import _lib


def letterbox_batch(images, input_size, out=None, stream=None):
    """images: list of HWC uint8 BGR CUDA tensors (any sizes, as cv2.imread yields them).
    Returns (samples, meta): samples (B, 3, S, S) uint8 RGB on the device; meta (B, 3) float64 on the
    device = (ratio, pad_w, pad_h) per image, the values `resize()` returns for mapping boxes back."""
    if not images:
        raise ValueError("letterbox_batch: empty batch")
    dev = images[0].device
    if dev.type != "cuda":
        raise RuntimeError("letterbox_batch runs on the GPU only (no CPU fallback); got a CPU tensor")
    rows = []
    keep = []
    for im in images:
        if im.device != dev or im.dtype != torch.uint8 or im.dim() != 3 or im.shape[2] != 3:
            raise ValueError("letterbox_batch: every image must be an HWC uint8 tensor with 3 channels on one device")
        h, w = int(im.shape[0]), int(im.shape[1])
        r = int(input_size) / max(h, w)
        if h < 1 or w < 1 or (r != 1 and min(int(h * r), int(w * r)) < 1):
            raise ValueError(f"letterbox_batch: a {h}x{w} image has no pixels left at input size {input_size} "
                             "(cv2.resize raises here in the reference, utils/dataset.py:99)")
        im = im.contiguous()
        keep.append(im)
        rows.append((im.data_ptr(), im.shape[0], im.shape[1]))
    B, S = len(rows), int(input_size)
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    # the descriptor upload and the kernel that reads it are ordered on the SAME stream (`stream=` may differ
    # from the current one)
    with torch.cuda.stream(st):
        desc = torch.tensor(rows, dtype=torch.int64).pin_memory().to(dev, non_blocking=True)
        if out is None:
            out = torch.empty((B, 3, S, S), dtype=torch.uint8, device=dev)
        meta = torch.empty((B, 3), dtype=torch.float64, device=dev)
    L = _lib.lib()
    with torch.cuda.device(dev):
        _lib.check(L.yb_letterbox(ctypes.c_void_p(desc.data_ptr()), B, S, ctypes.c_void_p(out.data_ptr()),
                                  ctypes.c_void_p(meta.data_ptr()), ctypes.c_void_p(st.cuda_stream)), "yb_letterbox")
    out._yb_keepalive = (keep, desc)   # the kernel reads these asynchronously
    return out, meta
