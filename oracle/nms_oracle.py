"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes wrapper of oracle/nms_oracle.c (plain-C restatement of
reference utils/util.py:123-169 + torchvision CPU nms).  Inputs/outputs are numpy arrays."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libnms_oracle.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "nms_oracle.c")
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, "libnms_oracle.so"], stdout=subprocess.DEVNULL)
        _lib = ctypes.CDLL(_SO)
        _lib.nms_oracle_batch.restype = None
        _lib.nms_oracle_batch.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_float, ctypes.c_double, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_float, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    return _lib


def non_max_suppression(pred, conf=0.001, iou=0.65, max_det=300, max_nms=30000, max_wh=7680.0, full_scan=False):
    """pred: (B, 4+nc, A) float32 array -> list of B arrays (k, 6) [x1,y1,x2,y2,score,cls]."""
    pred = np.ascontiguousarray(pred, dtype=np.float32)
    B, no, A = pred.shape
    out = np.zeros((B, max_det, 6), dtype=np.float32)
    counts = np.zeros(B, dtype=np.int32)
    _load().nms_oracle_batch(pred.ctypes.data, B, no - 4, A, float(np.float32(conf)), float(iou), max_det,
                             max_nms, float(max_wh), int(full_scan), out.ctypes.data, counts.ctypes.data)
    return [out[b, :counts[b]].copy() for b in range(B)]
