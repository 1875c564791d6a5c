"""Records calibrated BatchNorm running statistics for the synthetic weights (see
yolo_infer_pt_b200/synth.py): one training-mode forward with momentum 1.0 over a calibration batch,
so that running_mean/var equal the batch statistics of every layer.  Stored as fp16 so that every
machine loads bit-identical values.  Run once in the build container:

    python tests/golden/make_synth_bn.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from yolo_infer_pt_b200 import synth  # noqa: E402
from yolo_infer_pt_b200.nets import nn  # noqa: E402


def main(seed=0):
    for size in "ntsmlx":
        model = getattr(nn, f"yolo_v11_{size}")(80)
        model.load_state_dict(synth.synth_state_dict(model, seed, calibrated=False))
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.momentum = 1.0
        model.train()
        with torch.no_grad():
            model(synth.synth_images(4, 320, 320, seed=77))
        stats = {}
        for k, v in model.state_dict().items():
            if k.endswith("running_mean"):
                stats[k] = v.numpy().astype(np.float16)
            elif k.endswith("running_var"):
                stats[k] = np.maximum(v.numpy(), 1e-3).astype(np.float16)
        path = os.path.join(HERE, f"synth_bn_{size}_seed{seed}.npz")
        np.savez_compressed(path, **stats)
        print(size, len(stats), "tensors", f"{os.path.getsize(path) / 1024:.0f} KB")


if __name__ == "__main__":
    main()
