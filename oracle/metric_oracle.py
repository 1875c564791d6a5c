"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's detection consumer
`compute_metric` (utils/util.py:99-120), the step immediately behind non_max_suppression in `test()`
(main.py:271-299).  Follows the reference statement by statement: fp32 IoU matrix
with the same operation order (util.py:101-105), then per IoU threshold the [label, detect, iou] match
list, sorted by IoU descending (argsort()[::-1]), reduced with numpy.unique over detections and then
over labels (util.py:109-119).  Pinned in tests/test_oracle.py against fixtures recorded from the
reference function itself (tests/golden/metric_cases.npz).
"""
import numpy as np


def iou_matrix(output, target):
    """(M labels, N detections) fp32 IoU, util.py:101-105."""
    a1, a2 = target[:, None, 1:3].astype(np.float32), target[:, None, 3:5].astype(np.float32)
    b1, b2 = output[None, :, 0:2].astype(np.float32), output[None, :, 2:4].astype(np.float32)
    wh = np.clip(np.minimum(a2, b2) - np.maximum(a1, b1), 0, None)
    inter = wh[..., 0] * wh[..., 1]
    da, db = a2 - a1, b2 - b1
    return inter / (da[..., 0] * da[..., 1] + db[..., 0] * db[..., 1] - inter + np.float32(1e-7))


def compute_metric(output, target, iou_v):
    """output (N, 6) [x1, y1, x2, y2, conf, cls], target (M, 5) [cls, x1, y1, x2, y2], iou_v (T,) ->
    correct (N, T) bool."""
    output = np.asarray(output, dtype=np.float32)
    target = np.asarray(target, dtype=np.float32)
    iou_v = np.asarray(iou_v, dtype=np.float32)
    iou = iou_matrix(output, target)
    correct = np.zeros((output.shape[0], iou_v.shape[0]), dtype=bool)
    same = target[:, 0:1] == output[None, :, 5]
    for i in range(len(iou_v)):
        x = np.nonzero((iou >= iou_v[i]) & same)
        if x[0].shape[0]:
            matches = np.concatenate((np.stack(x, 1).astype(np.float32), iou[x[0], x[1]][:, None]), 1)
            if x[0].shape[0] > 1:
                matches = matches[matches[:, 2].argsort()[::-1]]
                matches = matches[np.unique(matches[:, 1], return_index=True)[1]]
                matches = matches[np.unique(matches[:, 0], return_index=True)[1]]
            correct[matches[:, 1].astype(int), i] = True
    return correct
