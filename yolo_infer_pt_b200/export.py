"""Engine export: the deployment artefact of the fused network.

The reference's `export_onnx` (utils/util.py:47-71) traces the fused model into an ONNX file for an external
runtime.  The equivalent here is the engine's own input: the architecture lists, the plan geometry and the packed
weight blob (BatchNorm folded as `fuse_conv` does, nets/nn.py:8-25; every conv in the K-major tensor-core layout of
its kernel).  `export_engine` writes them to one `.npz`; `load_engine` rebuilds a bound `Engine` from that file
alone - no `nets.nn` model, no checkpoint, no BatchNorm folding at load time.  Exporting needs no GPU (host-only
plan); loading for execution needs the sm_100 device like everything else on the path.
"""
import numpy as np
import torch

from .engine import Engine

FORMAT = 1


def export_engine(model, path, batch, height, width, act_dtype=torch.float16):
    """Pack `model` (a nets.nn.YOLO, fused or not) for a (batch, 3, height, width) input and write the artefact."""
    width_l, depth, csp, nc = model._arch
    eng = Engine(width_l, depth, csp, nc, batch, height, width, host_only=True, act_dtype=act_dtype)
    blob = eng.pack_from_model(model)
    np.savez_compressed(path, format=np.int64(FORMAT), width=np.asarray(width_l, np.int64),
                        depth=np.asarray(depth, np.int64), csp=np.asarray([int(bool(c)) for c in csp], np.int64),
                        num_classes=np.int64(nc), batch=np.int64(batch), height=np.int64(height), input_width=np.int64(width),
                        act_f16=np.int64(1 if act_dtype == torch.float16 else 0), weight_bytes=np.int64(eng.weight_bytes),
                        conv_names=np.asarray([c["name"] for c in eng.convs]), blob=blob)
    return path


def load_engine(path, device=None, host_only=False):
    """Rebuild the Engine an artefact was exported for and bind its packed weights.  The plan is re-derived from the
    architecture lists by the library, and the blob is only accepted if that plan has the same layout."""
    g = np.load(path if str(path).endswith(".npz") else str(path) + ".npz", allow_pickle=False)
    if int(g["format"]) != FORMAT:
        raise RuntimeError(f"engine artefact format {int(g['format'])} is not supported (expected {FORMAT})")
    act = torch.float16 if int(g["act_f16"]) else torch.bfloat16
    eng = Engine([int(v) for v in g["width"]], [int(v) for v in g["depth"]], [bool(v) for v in g["csp"]],
                 int(g["num_classes"]), int(g["batch"]), int(g["height"]), int(g["input_width"]), device,
                 host_only=host_only, act_dtype=act)
    names = [c["name"] for c in eng.convs]
    if eng.weight_bytes != int(g["weight_bytes"]) or names != [str(n) for n in g["conv_names"]]:
        raise RuntimeError("engine artefact does not match this library's plan for its architecture "
                           "(weight layout changed: re-export from the model)")
    eng.host_blob = np.ascontiguousarray(g["blob"], dtype=np.uint8)
    if not host_only:
        eng._bind()
    return eng
