"""The falsifiable float parity gate (run with -m gpu on a B200; every call goes through the C ABI).

The SURVEY 8(d) `survey` recipe damps the network so much that every class score is < 0.003 and the DFL
logits are almost constant: a 1e-2 / 0.5 px gate on it passes for any output.  These tests use

  * `survey_widehead` - same backbone, head tails x64, class biases centred on -5: scores spread over (0, 1)
    (> 5 % inside (0.1, 0.9)), DFL expectations spread over > 2 bins, and whatever 16-bit round-off arrives
    at the head is amplified 64x.  The ranges are asserted, then the north-star tolerance is asserted for the
    default fp16 activation storage; bf16 storage is measured next to it (it does not meet 0.5 px here, which is
    why fp16 - the reference's own evaluation dtype, main.py:251,266 - is the default);
  * a gain sweep from the damped to the chaotic regime with the fp32 oracle's OWN sensitivity (output change
    under a 1-bf16-ulp relative perturbation of the input) measured next to the GPU-vs-oracle error: the
    tolerance is asserted wherever the oracle itself is stable, and a bounded multiple of the sensitivity where
    it is not ("chaotic" is a measured statement, not an excuse);
  * the configurations BASELINE.json names that had no parity case: YOLO11x at 640x640, 1280x1280 inputs
    (1600-token attention, 33 600 anchors), images of the B = 256 bench tensor, and detections after NMS against
    the reference's own output (tests/golden/e2e_cd_n_640_nms.npz, written by make_golden.py).
"""
import os

import numpy as np
import pytest
import torch

from oracle import nms_oracle, yolo_oracle
from yolo_infer_pt_b200 import synth
from yolo_infer_pt_b200.nets import nn
from yolo_infer_pt_b200.utils import util

pytestmark = pytest.mark.gpu

BOX_TOL_PX = 0.5     # north star: max abs error on box coordinates
SCORE_TOL = 1e-2     # north star: max abs error on class scores
# the oracle counts as stable when a 2^-9 relative input perturbation moves its own outputs by less than this
SENS_BOX_PX, SENS_SCORE = 0.05, 1e-3
SENS_FACTOR = 20.0   # ... and beyond that the GPU may differ from it by at most this multiple of the sensitivity


def _model(size, recipe, **kw):
    m = getattr(nn, f"yolo_v11_{size}")(80)
    synth.load_synth(m, 0, recipe, **kw)
    return m.fuse().eval()


def _oracle(model, x):
    sd = {k: v.float().cpu() for k, v in model.state_dict().items()}
    with torch.no_grad():
        return yolo_oracle.forward(sd, *model._arch, x)


def _gpu(model, x, dtype=None):
    model.set_activation_dtype(dtype)
    with torch.no_grad():
        y = model.to("cuda:0")(x.to("cuda:0")).cpu()
    model.set_activation_dtype(None)
    return y


def _errs(y, ref):
    return (y[:, :4] - ref[:, :4]).abs().max().item(), (y[:, 4:] - ref[:, 4:]).abs().max().item()


@pytest.mark.parametrize("size,hw,batch", [("n", 320, 2), ("n", 640, 1), ("s", 320, 1), ("m", 128, 1), ("x", 128, 2)])
def test_widehead_gate(size, hw, batch):
    model = _model(size, "survey_widehead")
    x = synth.synth_images(batch, hw, hw, seed=0)
    sd = {k: v.float().cpu() for k, v in model.state_dict().items()}
    with torch.no_grad():
        maps = yolo_oracle.forward_raw(sd, *model._arch, x)
        ref = yolo_oracle.decode(maps, 80)
    # the gate can fail: scores span (0, 1) and the DFL expectations spread over bins
    sc = ref[:, 4:]
    mid = ((sc > 0.1) & (sc < 0.9)).float().mean().item()
    dist = torch.cat([(m[:, :64].reshape(m.shape[0], 4, 16, -1).softmax(2) * torch.arange(16.0).view(1, 1, 16, 1))
                      .sum(2).flatten() for m in maps])
    assert mid >= 0.05, f"only {mid:.3f} of the scores lie in (0.1, 0.9)"
    assert dist.std().item() >= 2.0, f"DFL expectations have std {dist.std().item():.2f} bins"
    assert sc.max().item() > 0.5 and (sc > 0.001).float().mean().item() > 0.3
    e16 = _errs(_gpu(model, x, torch.float16), ref)
    ebf = _errs(_gpu(model, x, torch.bfloat16), ref)
    print(f"widehead {size}@{hw} B={batch}: scores in (0.1,0.9) {mid:.3f}, max {sc.max():.3f}, DFL std {dist.std():.2f} bins | "
          f"fp16 storage: box {e16[0]:.3f} px, score {e16[1]:.2e} | bf16 storage: box {ebf[0]:.3f} px, score {ebf[1]:.2e}")
    assert e16[0] <= BOX_TOL_PX and e16[1] <= SCORE_TOL, "fp16 activation storage misses the north-star tolerance"
    # bf16 storage: measured, bounded (a kernel bug is hundreds of pixels), not held to 0.5 px
    assert ebf[0] <= 8.0 and ebf[1] <= 5e-2


SWEEP = [("survey", dict(gain=1.0)), ("survey", dict(gain=1.25)), ("survey", dict(gain=1.5)),
         ("survey", dict(gain=3 ** 0.5)),
         ("calibrated", dict(bn_gain=0.7, tail=10.0)), ("calibrated", dict(bn_gain=0.8, tail=10.0)),
         ("calibrated", dict(bn_gain=0.9, tail=10.0)), ("calibrated", dict(bn_gain=1.0))]


@pytest.mark.parametrize("recipe,kw", SWEEP)
def test_gain_sweep_error_vs_oracle_sensitivity(recipe, kw):
    """GPU-vs-oracle error next to the oracle's own sensitivity, from the damped to the chaotic regime."""
    model = _model("n", recipe, **kw)
    x = synth.synth_images(1, 320, 320, seed=0)
    ref = _oracle(model, x)
    rng = np.random.RandomState(5)
    xp = x * (1 + torch.from_numpy(rng.choice([-1.0, 1.0], tuple(x.shape)).astype(np.float32)) * 2.0 ** -9)
    sens = _errs(_oracle(model, xp), ref)
    e16 = _errs(_gpu(model, x, torch.float16), ref)
    ebf = _errs(_gpu(model, x, torch.bfloat16), ref)
    stable = sens[0] <= SENS_BOX_PX and sens[1] <= SENS_SCORE
    print(f"sweep {recipe} {kw}: oracle sensitivity box {sens[0]:.4f} px score {sens[1]:.2e} ({'stable' if stable else 'unstable'}) | "
          f"fp16: box {e16[0]:.4f} px score {e16[1]:.2e} | bf16: box {ebf[0]:.4f} px score {ebf[1]:.2e}")
    if stable:
        assert e16[0] <= BOX_TOL_PX and e16[1] <= SCORE_TOL
    else:
        assert e16[0] <= max(BOX_TOL_PX, SENS_FACTOR * sens[0]) and e16[1] <= max(SCORE_TOL, SENS_FACTOR * sens[1])
    assert torch.isfinite(torch.tensor(ebf)).all()


def test_x_at_640():
    """BASELINE config 4's model at its real input size (was only ever checked at 64x64)."""
    model = _model("x", "survey_widehead")
    x = synth.synth_images(2, 640, 640, seed=3)
    ref = _oracle(model, x)
    e = _errs(_gpu(model, x), ref)
    print(f"x@640 B=2 widehead: box {e[0]:.3f} px, score {e[1]:.2e}")
    assert e[0] <= BOX_TOL_PX and e[1] <= SCORE_TOL


def test_n_at_1280():
    """BASELINE config 5's input size: 1600-token attention, 33 600 anchors, 160 x 160 head maps."""
    model = _model("n", "survey_widehead")
    x = synth.synth_images(1, 1280, 1280, seed=4)
    ref = _oracle(model, x)
    y = _gpu(model, x)
    assert tuple(y.shape) == (1, 84, 33600)
    e = _errs(y, ref)
    print(f"n@1280 B=1 widehead: box {e[0]:.3f} px, score {e[1]:.2e}")
    assert e[0] <= BOX_TOL_PX and e[1] <= SCORE_TOL
    det = util.non_max_suppression(y.to("cuda:0"), 0.25, 0.7)
    ora = nms_oracle.non_max_suppression(y.numpy(), 0.25, 0.7)
    assert np.array_equal(det[0].cpu().numpy(), ora[0])


@pytest.mark.parametrize("recipe", ["survey", "survey_widehead"])
def test_bench_batch_spot_check(recipe):
    """Two images of the B = 256 tensor bench.py times (uint8 input, same seeds; `survey` is the bench's own
    recipe, `survey_widehead` the one on which the tolerance can fail), against the oracle."""
    model = _model("n", recipe)
    base = (synth.synth_images(8, 640, 640, seed=0) * 255).round().to(torch.uint8)
    x = base.repeat(32, 1, 1, 1).contiguous()
    with torch.no_grad():
        y = model.to("cuda:0")(x.to("cuda:0"))
        pick = [3, 255]
        got = y[pick].cpu()
    ref = _oracle(model, x[pick].float() / 255)
    e = _errs(got, ref)
    print(f"bench tensor n@640 B=256 {recipe}, images {pick}: box {e[0]:.3f} px, score {e[1]:.2e}")
    assert e[0] <= BOX_TOL_PX and e[1] <= SCORE_TOL
    # images 3 and 11 are the same picture: every image of the batch goes through the same arithmetic
    assert torch.equal(y[3], y[11]) and torch.equal(y[255], y[7])


def _match_detections(got, ref, iou_min=0.98, score_tol=5e-3):
    """Every reference detection has a GPU detection of the same class with IoU >= iou_min and score within
    score_tol (and vice versa); returns the number of unmatched rows on either side."""
    def iou(a, b):
        lt = np.maximum(a[:, None, :2], b[None, :, :2])
        rb = np.minimum(a[:, None, 2:4], b[None, :, 2:4])
        inter = np.clip(rb - lt, 0, None).prod(-1)
        area = lambda z: (z[:, 2] - z[:, 0]) * (z[:, 3] - z[:, 1])  # noqa: E731
        return inter / (area(a)[:, None] + area(b)[None] - inter + 1e-9)
    if len(got) == 0 or len(ref) == 0:
        return len(got) + len(ref)
    m = iou(ref, got)
    ok = (m >= iou_min) & (ref[:, None, 5] == got[None, :, 5]) & (np.abs(ref[:, None, 4] - got[None, :, 4]) <= score_tol)
    return int((~ok.any(1)).sum() + (~ok.any(0)).sum())


def test_widehead_matches_golden_reference_output(golden_dir):
    """The same gate against outputs of the REFERENCE ITSELF (fixtures written by make_golden.py)."""
    for size, hw in (("n", 320), ("x", 128)):
        g = np.load(os.path.join(golden_dir, f"fwdwh_{size}_{hw}.npz"))
        model = _model(size, "survey_widehead")
        y = _gpu(model, synth.synth_images(2, hw, hw, seed=0))
        ref = torch.from_numpy(g["out_sub"])
        e = _errs(y[:, :, torch.from_numpy(g["idx"])], ref)
        print(f"widehead {size}@{hw} vs reference golden: box {e[0]:.3f} px, score {e[1]:.2e}")
        assert e[0] <= BOX_TOL_PX and e[1] <= SCORE_TOL


def test_detections_match_the_reference_end_to_end(golden_dir):
    """forward + non_max_suppression on the GPU against the REFERENCE's own detections for the same weights and
    image (golden written by tests/golden/make_golden.py from /root/reference).  Recipe `calibrated_damped`:
    spatially structured scores (on survey_widehead all 8400 anchors of a class tie to 1e-3, so the greedy order
    - and with it the kept set - is decided by round-off in ANY implementation).  The 300 kept boxes compete with
    ~14 000 candidates, so a few near-ties still swap: >= 85 % of the rows must match (same class, IoU >= 0.98,
    score within 5e-3); a wrong implementation matches none."""
    g = np.load(os.path.join(golden_dir, "e2e_cd_n_640_nms.npz"))
    model = _model("n", "calibrated_damped")
    x = synth.synth_images(1, 640, 640, seed=0)
    conf, iou = float(g["conf"]), float(g["iou"])
    with torch.no_grad():
        y = model.to("cuda:0")(x.to("cuda:0"))
    det = util.non_max_suppression(y, conf, iou)[0].cpu().numpy()
    ref = g["det"]
    assert len(ref) > 20
    unmatched = _match_detections(det, ref)
    print(f"e2e n@640 calibrated_damped conf {conf} iou {iou}: reference {len(ref)} detections, GPU {len(det)}, "
          f"unmatched rows {unmatched} of {len(ref) + len(det)}")
    assert abs(len(det) - len(ref)) <= max(2, len(ref) // 20)
    assert unmatched <= 0.15 * (len(ref) + len(det))
    # and bit-exact NMS on the GPU's own predictions (the oracle C restatement of util.py:123-169)
    ora = nms_oracle.non_max_suppression(y.cpu().numpy(), conf, iou)[0]
    assert np.array_equal(det, ora)
