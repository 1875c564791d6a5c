"""Drop-in for the hot-path functions of the reference's `utils.util`
(t0saki/YOLO-Infer-pt utils/util.py:76-96, 123-169): `wh2xy`, `make_anchors`, `non_max_suppression`.

`non_max_suppression` keeps the reference signature and return type but runs entirely on the
device through libyolob200.so (`yb_nms`): one D2H copy of the per-image counts at the end instead of
several host synchronisations per image.  It raises for CPU tensors — there is no CPU fallback.
"""
import ctypes
import os
import sys

import numpy
import torch

MAX_WH = 7680      # utils/util.py:124
MAX_DET = 300      # utils/util.py:125
MAX_NMS = 30000    # utils/util.py:126


def _lib_module():
    try:
        from yolo_infer_pt_b200 import _lib
    except ImportError:
        root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        sys.path.insert(0, root)
        from yolo_infer_pt_b200 import _lib
    return _lib


def wh2xy(x):
    """(cx, cy, w, h) rows -> (x1, y1, x2, y2) rows; tensor or numpy (reference util.py:76-82)."""
    y = x.clone() if isinstance(x, torch.Tensor) else numpy.copy(x)
    half_w, half_h = x[:, 2] / 2, x[:, 3] / 2
    y[:, 0] = x[:, 0] - half_w
    y[:, 1] = x[:, 1] - half_h
    y[:, 2] = x[:, 0] + half_w
    y[:, 3] = x[:, 1] + half_h
    return y


def make_anchors(x, strides, offset=0.5):
    """Anchor centres (A,2) as (x+offset, y+offset) and strides (A,1) for a list of feature maps
    (reference util.py:85-96). The CUDA forward computes these from the anchor index instead;
    this helper remains for the training loss (util.py:864)."""
    assert x is not None
    points, scales = [], []
    for fmap, stride in zip(x, strides):
        h, w = fmap.shape[-2:]
        ys = torch.arange(h, device=fmap.device, dtype=fmap.dtype) + offset
        xs = torch.arange(w, device=fmap.device, dtype=fmap.dtype) + offset
        gy, gx = torch.meshgrid(ys, xs, indexing="ij")
        points.append(torch.stack((gx, gy), -1).view(-1, 2))
        scales.append(torch.full((h * w, 1), float(stride), dtype=fmap.dtype, device=fmap.device))
    return torch.cat(points), torch.cat(scales)


_workspaces = {}


def nms_workspace(batch, num_classes, num_anchors, device, max_nms=MAX_NMS):
    """A private NMS workspace (zeroed headers) for `Engine.forward(..., nms_sink=)` + `nms_padded(...,
    workspace=, prefiltered=True)`: the forward's class-score epilogues fill its candidate lists."""
    L = _lib_module().lib()
    return torch.zeros(L.yb_nms_workspace_bytes(batch, num_classes, num_anchors, max_nms), dtype=torch.uint8,
                       device=device)


def nms_padded(outputs, confidence_threshold=0.001, iou_threshold=0.65, max_det=MAX_DET, max_nms=MAX_NMS,
               max_wh=MAX_WH, workspace=None, prefiltered=False):
    """Device-resident NMS: returns (det, counts) with det (B, max_det, 6) fp32 rows
    [x1, y1, x2, y2, score, class] and counts (B,) int32 — no host synchronisation.
    `prefiltered`: `outputs` came from `Engine.forward(x, nms_sink=(workspace, conf, max_nms))` with the same
    threshold, so the candidate lists in `workspace` are already filled and the score pass is skipped."""
    if not (isinstance(outputs, torch.Tensor) and outputs.is_cuda):
        raise RuntimeError("yolo_infer_pt_b200.non_max_suppression needs a CUDA tensor; "
                           "there is no CPU fallback on the inference path")
    if outputs.dim() != 3 or outputs.shape[1] < 5:
        raise RuntimeError(f"expected (B, 4+nc, A) predictions, got {tuple(outputs.shape)}")
    _lib = _lib_module()
    L = _lib.lib()
    pred = outputs if outputs.dtype == torch.float32 else outputs.float()
    pred = pred.contiguous()
    B, no, A = pred.shape
    nc = no - 4
    dev = pred.device
    key = (dev.index, B, nc, A, max_nms)
    if prefiltered and workspace is None:
        raise ValueError("prefiltered NMS needs the workspace the forward appended its candidates to")
    ws = workspace if workspace is not None else _workspaces.get(key)
    if ws is None:
        nbytes = L.yb_nms_workspace_bytes(B, nc, A, max_nms)
        ws = torch.zeros(nbytes, dtype=torch.uint8, device=dev)   # zero headers once; the kernel keeps them zero
        _workspaces.clear()
        _workspaces[key] = ws
    # the kernel writes every row (zeros past the last detection) and every count: no fill kernels
    det = torch.empty(B, max_det, 6, dtype=torch.float32, device=dev)
    counts = torch.empty(B, dtype=torch.int32, device=dev)
    conf32 = float(numpy.float32(confidence_threshold))  # torch compares an fp32 tensor in fp32
    if prefiltered and getattr(ws, "_yb_sink_tag", None) != (conf32, int(max_nms), pred.data_ptr()):
        raise ValueError("prefiltered NMS: the workspace was not filled by a forward of these predictions with "
                         "the same confidence threshold and max_nms")
    stream = torch.cuda.current_stream(dev).cuda_stream
    fn = L.yb_nms_prefiltered if prefiltered else L.yb_nms_clean
    with torch.cuda.device(dev):
        _lib.check(fn(pred.data_ptr(), B, nc, A, conf32, float(iou_threshold), max_det, max_nms,
                      float(max_wh), det.data_ptr(), counts.data_ptr(), ws.data_ptr(), ws.numel(),
                      ctypes.c_void_p(stream)), "yb_nms")
    if prefiltered:
        ws._yb_sink_pending = False     # consumed: the per-image kernel left the headers zeroed
    return det, counts


def non_max_suppression(outputs, confidence_threshold=0.001, iou_threshold=0.65):
    """Reference signature and result (util.py:123-169): list of B tensors (k<=300, 6) fp32
    [x1, y1, x2, y2, confidence, class] on outputs.device, score-descending."""
    det, counts = nms_padded(outputs, confidence_threshold, iou_threshold)
    counts = counts.cpu().tolist()  # the single host synchronisation of the call
    return [det[i, :k] for i, k in enumerate(counts)]


def compute_metric_batch(det, counts, targets, target_counts, iou_v):
    """Batched, device-resident `compute_metric` (reference util.py:99-120) over the padded output of
    `nms_padded`: det (B, max_det, 6), counts (B,), targets (B, max_t, 5) rows [class, x1, y1, x2, y2] in
    pixels, target_counts (B,), iou_v (T,).  Returns correct (B, max_det, T) bool; no host round trip."""
    if not det.is_cuda:
        raise RuntimeError("compute_metric runs on the GPU only (no CPU fallback)")
    _lib = _lib_module()
    L = _lib.lib()
    dev = det.device
    det = det.float().contiguous()
    counts = counts.to(dev, torch.int32).contiguous()
    targets = targets.to(dev, torch.float32).contiguous()
    target_counts = target_counts.to(dev, torch.int32).contiguous()
    iou_v = iou_v.to(dev, torch.float32).contiguous()
    B, max_det = det.shape[0], det.shape[1]
    max_t = targets.shape[1]
    correct = torch.empty((B, max_det, iou_v.numel()), dtype=torch.uint8, device=dev)
    tptr = targets.data_ptr() if max_t else det.data_ptr()   # never dereferenced when there are no labels
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        _lib.check(L.yb_compute_metric(det.data_ptr(), counts.data_ptr(), tptr, target_counts.data_ptr(), B, max_det,
                                       max_t, iou_v.data_ptr(), iou_v.numel(), correct.data_ptr(),
                                       ctypes.c_void_p(stream)), "yb_compute_metric")
    return correct.bool()


def compute_metric(output, target, iou_v):
    """Reference signature (util.py:99): output (N, 6) detections of one image, target (M, 5) labels
    [class, x1, y1, x2, y2], iou_v (T,) -> correct (N, T) bool on output.device."""
    n, m = output.shape[0], target.shape[0]
    dev = output.device
    cnt = torch.tensor([n], dtype=torch.int32, device=dev)
    tcnt = torch.tensor([m], dtype=torch.int32, device=dev)
    det = output[None, :, :6] if n else torch.zeros((1, 1, 6), device=dev)
    tgt = target[None] if m else torch.zeros((1, 0, 5), device=dev)
    return compute_metric_batch(det, cnt, tgt, tcnt, iou_v)[0, :n]


def load_weight(model, ckpt):
    """Reference util.py:345-355: copy every tensor of the checkpoint's `model` whose key and shape match
    into `model` (strict=False).  `ckpt` is a path (or an already loaded dict); its 'model' entry may be a
    state_dict or a pickled module - the reference's own class works when `yolo_infer_pt_b200/` provides
    `nets.nn` on sys.path."""
    dst = model.state_dict()
    obj = ckpt if isinstance(ckpt, dict) else torch.load(ckpt, map_location="cpu", weights_only=False)
    src = obj["model"]
    src = src.float().cpu().state_dict() if isinstance(src, torch.nn.Module) else src
    keep = {k: v.float() for k, v in src.items() if k in dst and v.shape == dst[k].shape}
    model.load_state_dict(state_dict=keep, strict=False)
    return model


# ---- Ultralytics checkpoint import (reference util.py:358-516, corrected) -----------------------------------
# Ultralytics YOLO11 numbers its layers 0..23 (yolo11.yaml): 0-10 backbone (Conv, Conv, C3k2, Conv, C3k2, Conv, C3k2,
# Conv, C3k2, SPPF, C2PSA), 11-22 neck (13/16/19/22 C3k2, 17/20 Conv; the rest are Upsample / Concat without
# weights), 23 Detect.  The reference's table (util.py:372-475) gets three families of keys wrong and silently skips
# them ("Skipping ... missing key"):
#   * util.py:406-409,425-429  nested C3k bottlenecks `6.m.0.m.0.cv1` are mapped to `...res_m.0.m.0.conv1`; the module
#     is `res_m.0.res_m.0.conv1` (nets/nn.py:52-63), and C2PSA `10.m.0.attn.qkv / proj / pe / ffn.N` are mapped to
#     `net.p5.3.m.0.attn.*`; the modules are `net.p5.3.res_m.0.conv1.{qkv, conv2, conv1}` / `.conv2.N` (nn.py:97-148);
#   * util.py:454-477  Detect: Ultralytics `cv2` is the BOX branch (Conv, Conv, Conv2d) and `cv3` the CLASS branch
#     (two DWConv+Conv pairs, Conv2d); the table sends cv2 -> head.cls and cv3 -> head.box with indices that do not
#     exist, so the whole head keeps its random initialisation;
#   * every direct table hit skips the `.bn.` -> `.norm.` rename (only the fallback path applies it), so the
#     BatchNorm statistics of all C3k2 / SPPF / C2PSA convs are dropped as well.
# Here the map is derived from the structure instead of a table, for every model size.
_ULTRA_LAYERS = {0: "net.p1.0", 1: "net.p2.0", 2: "net.p2.1", 3: "net.p3.0", 4: "net.p3.1", 5: "net.p4.0", 6: "net.p4.1",
                 7: "net.p5.0", 8: "net.p5.1", 9: "net.p5.2", 10: "net.p5.3", 13: "fpn.h1", 16: "fpn.h2", 17: "fpn.h3",
                 19: "fpn.h4", 20: "fpn.h5", 22: "fpn.h6"}
_ULTRA_TOKENS = {"cv1": "conv1", "cv2": "conv2", "cv3": "conv3", "m": "res_m", "bn": "norm"}
_ULTRA_ATTN = {"qkv": "qkv", "proj": "conv2", "pe": "conv1"}


def ultralytics_key(key):
    """Ultralytics YOLO11 state_dict key ('model.6.m.0.m.1.cv2.bn.weight') -> this package's / the reference's key
    ('net.p4.1.res_m.0.res_m.1.conv2.norm.weight'), or None for keys that have no counterpart."""
    parts = key.split(".")
    if parts[0] == "model":
        parts = parts[1:]
    if len(parts) < 2 or not parts[0].isdigit():
        return None
    idx, rest = int(parts[0]), parts[1:]
    if idx == 23:                                            # Detect
        if rest[0] == "dfl":
            return "head." + ".".join(rest)
        if rest[0] == "cv2" and len(rest) >= 4:              # box branch: Conv, Conv, Conv2d
            lvl, j, tail = rest[1], int(rest[2]), rest[3:]
            return f"head.box.{lvl}.{j}." + ".".join("norm" if t == "bn" else t for t in tail)
        if rest[0] == "cv3" and len(rest) >= 4:              # class branch: (DWConv, Conv), (DWConv, Conv), Conv2d
            lvl, a = rest[1], int(rest[2])
            if a == 2:
                return f"head.cls.{lvl}.4." + ".".join(rest[3:])
            if len(rest) < 5:
                return None
            return f"head.cls.{lvl}.{2 * a + int(rest[3])}." + ".".join("norm" if t == "bn" else t for t in rest[4:])
        return None
    if idx not in _ULTRA_LAYERS:
        return None
    out, i = [_ULTRA_LAYERS[idx]], 0
    while i < len(rest):
        t = rest[i]
        if t == "attn" and i + 1 < len(rest):                # PSABlock.attn.{qkv,proj,pe} -> conv1.{qkv,conv2,conv1}
            out += ["conv1", _ULTRA_ATTN.get(rest[i + 1], rest[i + 1])]
            i += 2
        elif t == "ffn":                                     # PSABlock.ffn.N -> conv2.N
            out.append("conv2")
            i += 1
        else:
            out.append(_ULTRA_TOKENS.get(t, t))
            i += 1
    return ".".join(out)


def load_ultralytics_weight(model, ckpt, strict=True, verbose=False):
    """Reference util.py:358-516 with the key map corrected (see above): load an Ultralytics-format YOLO11
    checkpoint (`yolo11n.pt`: {'model': DetectionModel}) or a plain Ultralytics state_dict into `model`.
    strict: raise unless every tensor of `model` was found with the right shape (the reference prints and goes on,
    which leaves whole branches at their random initialisation)."""
    obj = ckpt if isinstance(ckpt, dict) else torch.load(ckpt, map_location="cpu", weights_only=False)
    src = obj["model"] if "model" in obj and not isinstance(obj["model"], torch.Tensor) else obj
    src = src.float().state_dict() if isinstance(src, torch.nn.Module) else src
    dst = model.state_dict()
    new, skipped = {}, []
    for k, v in src.items():
        dk = ultralytics_key(k)
        if dk is None and k in dst:
            dk = k
        if dk in dst and tuple(v.shape) == tuple(dst[dk].shape):
            new[dk] = v.to(dst[dk].dtype) if v.is_floating_point() else v
        else:
            skipped.append((k, dk))
    missing = [k for k in dst if k not in new]
    if verbose:
        for k, dk in skipped:
            print(f"load_ultralytics_weight: skipped {k} -> {dk}")
    if strict and missing:
        raise KeyError(f"load_ultralytics_weight: {len(missing)} tensors of the model were not found in the checkpoint "
                       f"(first: {missing[:4]}); {len(skipped)} checkpoint tensors had no target (first: {skipped[:4]})")
    model.load_state_dict(new, strict=False)
    return model
