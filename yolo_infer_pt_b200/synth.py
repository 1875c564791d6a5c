"""Deterministic synthetic weights and inputs (no checkpoints exist offline: the reference's
weights_old/best.pt is listed in .MISSING_LARGE_BLOBS).  numpy's Mersenne Twister is used so the
same tensors are produced on any machine / torch version; SURVEY.md §8(d) `synth_weights`.

  conv weights      U(-b, b), b = sqrt(3/fan_in)        (variance preserving)
  BatchNorm         weight ~ U(0.8, 1.2), bias ~ N(0, 0.1); running_mean / running_var = calibrated
                    per-channel batch statistics (committed fixture) so that activations keep unit
                    scale and spatial structure at every depth — with arbitrary running statistics a
                    random network collapses to a constant and a parity test would see nothing
  head box tails    bias 1.0 (nets/nn.py:277)
  head cls tails    bias CLS_BIAS + N(0, 0.5) so that a few % of the scores exceed the 0.001
                    confidence threshold (the reference's bias init, nn.py:279, tops out at 1.6e-4
                    on random weights and NMS would see no candidates)
"""
import numpy as np
import torch

import os

CLS_BIAS = -8.5
_BN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _variant_name(model):
    width = tuple(model.net.p1[0].conv.weight.shape[:1]) + tuple(model.head.box[0][0].conv.weight.shape[1:2])
    depth = len(model.net.p2[1].res_m)
    table = {(16, 64, 1): "n", (24, 96, 1): "t", (32, 128, 1): "s", (64, 256, 1): "m", (64, 256, 2): "l",
             (96, 384, 2): "x"}
    return table.get((width[0], width[1], depth))


WIDEHEAD_TAIL = 64.0      # survey_widehead: factor on the weights of the six head tail convs
WIDEHEAD_CLS_BIAS = -5.0  # survey_widehead: centre of the class-tail biases


def synth_state_dict(model, seed=0, recipe="calibrated", calibrated=True, gain=None, bn_gain=1.0, tail=None):
    """Seeded weights for `model` (reference or product YOLO, same state_dict keys).

    recipe="survey"      SURVEY.md §8(d) `synth_weights`: torch-default conv init range
                         U(+-1/sqrt(fan_in)), arbitrary BatchNorm statistics.  Activations shrink
                         with depth, so round-off is barely amplified - but the head is nearly constant
                         (every score < 0.003): a parity gate on it cannot fail.
    recipe="survey_widehead"  the same backbone with the six head tail convs (head.box.*.2, head.cls.*.4)
                         scaled by WIDEHEAD_TAIL and the class-tail biases centred on WIDEHEAD_CLS_BIAS:
                         class scores spread over (0, 1) (> 5 % inside (0.1, 0.9)), DFL distributions
                         vary by side and level, and 16-bit round-off arriving at the head is amplified
                         64x - the recipe of the falsifiable tolerance gate (tests/test_gpu_parity.py).
    recipe="calibrated"  variance-preserving conv init and BatchNorm running statistics taken from
                         tests/golden/synth_bn_<size>_seed<seed>.npz (per-channel batch statistics
                         recorded by tests/golden/make_synth_bn.py, stored as fp16 so every machine
                         loads identical values).  Every layer keeps unit scale and spatial
                         structure — the recipe for layer-level parity (a tap-order or slice bug is
                         invisible on a collapsed network).  At bn_gain = 1 the network is chaotic (a
                         1-ulp input perturbation moves the fp32 oracle's boxes by hundreds of pixels).
    gain      conv init bound = gain / sqrt(fan_in) (default 1 for survey, sqrt(3) for calibrated)
    bn_gain   factor on every BatchNorm weight: < 1 damps the calibrated network below the edge of chaos
    tail      factor on the head tail weights (default WIDEHEAD_TAIL for survey_widehead, else 1)
    """
    wide = recipe == "survey_widehead"
    if wide:
        recipe = "survey"
    if recipe == "calibrated_damped":
        # calibrated network pulled below the edge of chaos (BatchNorm weights x0.8: a 1-ulp input perturbation
        # moves the fp32 oracle by < 0.5 px instead of hundreds) with the head tails x30: spatially structured
        # scores and boxes - the recipe of the detection-level end-to-end test
        recipe, bn_gain = "calibrated", 0.8 if bn_gain == 1.0 else bn_gain
        tail = 30.0 if tail is None else tail
    if tail is None:
        tail = WIDEHEAD_TAIL if wide else 1.0
    rng = np.random.RandomState(seed)
    out = {}
    bn = None
    size = _variant_name(model)
    survey = recipe == "survey"
    if not survey and calibrated and size is not None:
        path = os.path.join(_BN_DIR, f"synth_bn_{size}_seed{seed}.npz")
        if os.path.exists(path):
            bn = np.load(path)
    for key, ref in model.state_dict().items():
        shape = tuple(ref.shape)
        if key.endswith("num_batches_tracked"):
            val = np.zeros(shape, dtype=np.int64)
        elif key == "head.dfl.conv.weight":
            val = np.arange(16, dtype=np.float32).reshape(shape)
        elif key.endswith("running_mean"):
            val = rng.normal(0.0, 0.1, shape)
        elif key.endswith("running_var"):
            val = rng.uniform(0.5, 1.5, shape)
        elif key.endswith("norm.weight"):
            val = rng.uniform(0.8, 1.2, shape)
        elif key.endswith("norm.bias"):
            val = rng.normal(0.0, 0.1, shape)
        elif key.endswith(".weight"):  # conv weight OIHW, variance preserving: std = 1/sqrt(fan_in)
            fan_in = int(np.prod(shape[1:]))
            bound = (gain if gain is not None else (1.0 if survey else np.sqrt(3.0))) / np.sqrt(fan_in)
            val = rng.uniform(-bound, bound, shape)
        elif key.endswith(".bias"):
            if key.startswith("head.box."):
                val = np.full(shape, 1.0)
            elif key.startswith("head.cls."):
                val = (np.full(shape, WIDEHEAD_CLS_BIAS if wide else -9.0) + rng.normal(0.0, 1.0, shape)) if survey else \
                    (np.full(shape, CLS_BIAS) + rng.normal(0.0, 0.5, shape))
            else:
                val = rng.normal(0.0, 0.1, shape)
        else:
            raise KeyError(key)
        if bn is not None and key in bn.files:
            val = bn[key].astype(np.float32)
        if key.endswith("norm.weight"):
            val = np.asarray(val) * bn_gain
        if tail != 1.0 and key.endswith(".weight") and (
                (key.startswith("head.box.") and key.split(".")[3] == "2") or
                (key.startswith("head.cls.") and key.split(".")[3] == "4")):
            val = np.asarray(val) * tail
        out[key] = torch.from_numpy(np.asarray(val)).to(ref.dtype)
    return out


def load_synth(model, seed=0, recipe="calibrated", **kw):
    model.load_state_dict(synth_state_dict(model, seed, recipe, **kw))
    return model


def synth_images(batch, height, width, seed=0):
    """Structured fp32 NCHW images in [0,1]: random coloured rectangles over a low-amplitude noise
    floor (iid noise alone averages out after the first stride-2 stages and exercises nothing)."""
    rng = np.random.RandomState(1000 + seed)
    img = rng.random_sample((batch, 3, height, width)) * 0.25
    for b in range(batch):
        for _ in range(24):
            h = int(rng.randint(max(2, height // 16), max(3, height // 2)))
            w = int(rng.randint(max(2, width // 16), max(3, width // 2)))
            y0 = int(rng.randint(0, max(1, height - h)))
            x0 = int(rng.randint(0, max(1, width - w)))
            colour = rng.random_sample(3)
            alpha = rng.uniform(0.5, 1.0)
            patch = img[b, :, y0:y0 + h, x0:x0 + w]
            patch *= (1.0 - alpha)
            patch += alpha * colour[:, None, None]
    return torch.from_numpy(np.clip(img, 0.0, 1.0).astype(np.float32))


def synth_predictions(batch, nc, anchors, img=640, mode="sparse", seed=0):
    """Synthetic (B, 4+nc, A) fp32 prediction tensors for NMS tests:
    cx,cy ~ U(0,img), w,h ~ U(4,204); scores sparse = rand^8 * Bernoulli(0.02) or dense = a
    permutation ladder in (0,1).  Scores are tie-free among candidates (the reference's argsort is
    unstable on ties, utils/util.py:157)."""
    rng = np.random.RandomState(2000 + seed)
    pred = np.empty((batch, 4 + nc, anchors), dtype=np.float32)
    pred[:, 0:2] = rng.uniform(0, img, (batch, 2, anchors))
    pred[:, 2:4] = rng.uniform(4, 204, (batch, 2, anchors))
    n_el = nc * anchors
    if mode == "dense":
        s = np.stack([(rng.permutation(n_el).astype(np.float64) + 1.0) / (n_el + 1.0) for _ in range(batch)])
        s = s.reshape(batch, nc, anchors)
    elif mode == "sparse":
        s = rng.random_sample((batch, nc, anchors)) ** 8 * (rng.random_sample((batch, nc, anchors)) < 0.02)
    elif mode == "few":
        s = rng.random_sample((batch, nc, anchors)) * (rng.random_sample((batch, nc, anchors)) < 2e-4)
    elif mode == "empty":
        s = np.zeros((batch, nc, anchors))
    else:
        raise ValueError(mode)
    s = s.astype(np.float32)
    for b in range(batch):  # nudge exact fp32 duplicates among candidates to the next free float
        flat = s[b].reshape(-1)
        nz = np.flatnonzero(flat > 0)
        vals = flat[nz]
        order = np.argsort(vals, kind="stable")
        sv = vals[order]
        for _ in range(64):
            dup = sv[1:] <= sv[:-1]
            if not dup.any():
                break
            sv[1:][dup] = np.nextafter(sv[:-1][dup], np.float32(2.0)).astype(np.float32)
        vals[order] = sv
        flat[nz] = vals
    pred[:, 4:] = s
    return pred
