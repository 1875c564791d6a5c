"""yolo_infer_pt_b200 — B200 (sm_100a) implementation of the YOLO-Infer-pt inference hot path.

Two ways in:

  * drop-in, with the reference's own import lines (main.py:12-13) — also what un-pickling a
    reference checkpoint (`nets.nn.YOLO`) needs:

        import sys; sys.path.insert(0, "<repo>/yolo_infer_pt_b200")
        from nets import nn
        from utils import util

  * as a package: `from yolo_infer_pt_b200.nets import nn`, `from yolo_infer_pt_b200.utils import util`.

Importing never touches CUDA or loads the shared library (the reference forks DataLoader
workers, utils/util.py:33); libyolob200.so is opened on the first forward / NMS call.
"""
__version__ = "0.1.0"
