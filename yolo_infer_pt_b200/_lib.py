"""ctypes binding of libyolob200.so (C ABI declared in include/yolob200.h).

The library is loaded lazily on first use, never at import time: the reference forks DataLoader
workers (utils/util.py:33), so importing this package must not touch CUDA.  There is no CPU
implementation behind these symbols; if the shared object is missing the caller gets an error.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libyolob200.so")

YB_F32, YB_F16, YB_BF16, YB_U8 = 0, 1, 2, 3


class ArchDesc(ctypes.Structure):
    _fields_ = [
        ("width", ctypes.c_int * 6),
        ("depth", ctypes.c_int * 6),
        ("csp", ctypes.c_int * 2),
        ("num_classes", ctypes.c_int),
    ]


class ConvInfo(ctypes.Structure):
    _fields_ = [
        ("name", ctypes.c_char * 96),
        ("cout", ctypes.c_int),
        ("cin", ctypes.c_int),
        ("ksize", ctypes.c_int),
        ("stride", ctypes.c_int),
        ("groups", ctypes.c_int),
        ("act", ctypes.c_int),
        ("wrapped", ctypes.c_int),
        ("kind", ctypes.c_int),
        ("blob_offset", ctypes.c_size_t),
        ("blob_bytes", ctypes.c_size_t),
    ]


# name -> (restype, argtypes); every symbol include/yolob200.h declares
SYMBOLS = {
    "yb_plan_create": (ctypes.c_int, [ctypes.POINTER(ArchDesc), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "yb_plan_destroy": (None, [ctypes.c_void_p]),
    "yb_plan_workspace_bytes": (ctypes.c_size_t, [ctypes.c_void_p]),
    "yb_plan_weight_bytes": (ctypes.c_size_t, [ctypes.c_void_p]),
    "yb_plan_num_anchors": (ctypes.c_int, [ctypes.c_void_p]),
    "yb_plan_num_outputs": (ctypes.c_int, [ctypes.c_void_p]),
    "yb_plan_num_convs": (ctypes.c_int, [ctypes.c_void_p]),
    "yb_plan_conv_info": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ConvInfo)]),
    "yb_plan_num_launches": (ctypes.c_int, [ctypes.c_void_p]),
    "yb_plan_pack_conv": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p]),
    "yb_plan_bind": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "yb_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                  ctypes.c_void_p]),
    "yb_forward_nms": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_float,
                                      ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "yb_forward_raw": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                      ctypes.c_void_p]),
    "yb_plan_use_graph": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "yb_plan_profile": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "yb_plan_profile_read": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "yb_plan_set_conv_impl": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "yb_plan_debug_read": (ctypes.c_longlong, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p,
                                               ctypes.c_size_t, ctypes.POINTER(ctypes.c_int),
                                               ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
    "yb_plan_debug_write": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t]),
    "yb_plan_run_op": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                      ctypes.c_void_p, ctypes.c_void_p]),
    "yb_plan_describe": (ctypes.c_longlong, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t]),
    "yb_nms_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "yb_nms": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                              ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_void_p,
                              ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "yb_nms_clean": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                    ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_void_p,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "yb_nms_prefiltered": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                          ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "yb_nms_workspace_init": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "yb_compute_metric": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                         ctypes.c_void_p]),
    "yb_letterbox": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_void_p]),
    "yb_last_error": (ctypes.c_char_p, []),
    "yb_launch_count": (ctypes.c_ulonglong, []),
    "yb_version": (ctypes.c_int, []),
}

_lib = None
_lock = threading.Lock()


def lib():
    """Return the loaded library, raising if the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} not found: build the CUDA extension first "
                        "(python -c 'import __graft_entry__ as g; g.build()' or ./build.sh). "
                        "yolo_infer_pt_b200 has no CPU or PyTorch fallback for the inference path.")
                handle = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in SYMBOLS.items():
                    fn = getattr(handle, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = handle
    return _lib


def check(rc, what=""):
    if rc < 0:
        msg = lib().yb_last_error()
        raise RuntimeError(f"libyolob200 {what} failed ({rc}): {msg.decode() if msg else ''}")
    return rc
