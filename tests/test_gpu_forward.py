"""GPU parity tests of the forward path (run with -m gpu on a B200).  Every call goes through the
C ABI (libyolob200.so via ctypes); the oracle is only the checker."""
import os

import numpy as np
import pytest
import torch

from oracle import yolo_oracle
from oracle.plan_replay import PlanReplay
from yolo_infer_pt_b200 import _lib, synth
from yolo_infer_pt_b200.engine import Engine
from yolo_infer_pt_b200.nets import nn

pytestmark = pytest.mark.gpu

BOX_TOL_PX = 0.5     # north star: max abs error on box coordinates
SCORE_TOL = 1e-2     # north star: max abs error on class scores


def _model(size, recipe):
    m = getattr(nn, f"yolo_v11_{size}")(80)
    synth.load_synth(m, 0, recipe)
    return m.fuse().eval()


def _oracle(model, x):
    sd = {k: v.float().cpu() for k, v in model.state_dict().items()}
    with torch.no_grad():
        return yolo_oracle.forward(sd, *model._arch, x)


def _taps_from_engine(eng, desc):
    out = {}
    for op in desc["ops"]:
        if op["kind"] in (1, 2, 4):
            t = eng.debug_read(op["name"])
            r0 = op["dst_row_off"]
            out[op["name"]] = t[:, r0:r0 + op["Hout"] * op["Wout"]]
    return out


@pytest.mark.parametrize("size,hw", [("n", 64), ("n", 160), ("x", 64)])
def test_layers_tensor_core_vs_direct_vs_cpu_replay(size, hw, monkeypatch):
    """Layer-level parity on the calibrated recipe (every layer has unit-scale, spatially varying
    activations): tcgen05 kernel == scalar cross-check kernel == bf16-faithful CPU replay of the plan."""
    monkeypatch.setenv("YB_NO_REUSE", "1")
    model = _model(size, "calibrated")
    dev = torch.device("cuda:0")
    x = synth.synth_images(2, hw, hw, seed=1)
    eng = Engine(*model._arch, 2, hw, hw, dev)
    blob = eng.pack_from_model(model)
    desc = eng.describe()
    cpu_taps = {}
    with torch.no_grad():
        PlanReplay(desc, eng.convs, blob, emulate_bf16=True).run(x, taps=cpu_taps)
    xg = x.to(dev)
    eng.set_conv_impl(1)
    y_direct = eng.forward(xg).clone()
    direct = _taps_from_engine(eng, desc)
    eng.set_conv_impl(0)
    y_tc = eng.forward(xg).clone()
    tc = _taps_from_engine(eng, desc)
    torch.cuda.synchronize()
    worst = []
    for name, ref in cpu_taps.items():
        if name not in tc:
            continue
        scale = max(1.0, ref.abs().max().item())
        e_tc = (tc[name] - ref).abs().max().item() / scale
        e_dir = (direct[name] - ref).abs().max().item() / scale
        e_x = (tc[name] - direct[name]).abs().max().item() / scale
        worst.append((max(e_tc, e_dir, e_x), name, e_tc, e_dir, e_x))
    worst.sort(reverse=True)
    report = "\n".join(f"{n:42s} tc-cpu {a:.4f} direct-cpu {b:.4f} tc-direct {c:.4f}" for _, n, a, b, c in worst[:12])
    print(report)
    # bf16 stores differ by at most a few ulp (0.4 % each) between summation orders; errors compound
    # slowly with depth, a real bug shows up as O(1)
    assert worst[0][0] < 0.08, report
    assert (y_tc - y_direct).abs().max().item() < 0.1 * max(1.0, y_direct.abs().max().item())


@pytest.mark.parametrize("size,hw,batch", [("n", 64, 2), ("t", 64, 1), ("s", 64, 1), ("m", 64, 1), ("l", 64, 1),
                                           ("x", 64, 2), ("n", 640, 2), ("s", 320, 1)])
def test_forward_within_north_star_tolerance(size, hw, batch):
    """Pre-NMS head outputs vs the fp32 oracle on the SURVEY §8(d) synthetic recipe:
    boxes within 0.5 px, scores within 1e-2 (bf16 activations, fp32 accumulation and decode)."""
    model = _model(size, "survey")
    x = synth.synth_images(batch, hw, hw, seed=1 if hw == 64 else 0)
    ref = _oracle(model, x)
    before = _lib.lib().yb_launch_count()
    with torch.no_grad():
        y = model.to("cuda:0")(x.to("cuda:0"))
    torch.cuda.synchronize()
    assert _lib.lib().yb_launch_count() > before, "no kernel of libyolob200 ran"
    assert y.dtype == torch.float32 and tuple(y.shape) == tuple(ref.shape)
    box_err = (y.cpu()[:, :4] - ref[:, :4]).abs().max().item()
    cls_err = (y.cpu()[:, 4:] - ref[:, 4:]).abs().max().item()
    print(f"{size}@{hw}: box max-abs {box_err:.4f} px, score max-abs {cls_err:.2e}")
    assert box_err <= BOX_TOL_PX
    assert cls_err <= SCORE_TOL


@pytest.mark.parametrize("size", ["n", "x"])
def test_forward_matches_golden_reference_output(size, golden_dir):
    g = np.load(os.path.join(golden_dir, f"fwdsv_{size}_64.npz"))
    model = _model(size, "survey").to("cuda:0")
    with torch.no_grad():
        y = model(synth.synth_images(2, 64, 64, seed=1).to("cuda:0")).cpu()
    ref = torch.from_numpy(g["out"])
    assert (y[:, :4] - ref[:, :4]).abs().max() <= BOX_TOL_PX
    assert (y[:, 4:] - ref[:, 4:]).abs().max() <= SCORE_TOL


def test_forward_640_matches_golden_subsample(golden_dir):
    g = np.load(os.path.join(golden_dir, "fwd_n_640.npz"))
    model = _model("n", "survey").to("cuda:0")
    with torch.no_grad():
        y = model(synth.synth_images(1, 640, 640, seed=0).to("cuda:0")).cpu()
    sub = y[:, :, torch.from_numpy(g["idx"])]
    ref = torch.from_numpy(g["out_sub"])
    assert (sub[:, :4] - ref[:, :4]).abs().max() <= BOX_TOL_PX
    assert (sub[:, 4:] - ref[:, 4:]).abs().max() <= SCORE_TOL


def test_calibrated_recipe_round_off_report():
    """Round-off stress: unit-gain random weights amplify bf16 rounding ~10x more than the survey
    recipe.  Asserted against the bf16-faithful CPU replay (tight) and reported against fp32."""
    model = _model("n", "calibrated")
    x = synth.synth_images(1, 320, 320, seed=2)
    eng = Engine(*model._arch, 1, 320, 320, "cuda:0")
    blob = eng.pack_from_model(model)
    with torch.no_grad():
        rep = PlanReplay(eng.describe(), eng.convs, blob, emulate_bf16=True).run(x)
    y = eng.forward(x.to("cuda:0")).cpu()
    ref = _oracle(model, x)
    print(f"calibrated n@320 vs fp32 oracle: box {(y[:, :4] - ref[:, :4]).abs().max():.3f} px, "
          f"score {(y[:, 4:] - ref[:, 4:]).abs().max():.4f}; vs bf16 CPU replay: "
          f"box {(y[:, :4] - rep[:, :4]).abs().max():.3f} px, score {(y[:, 4:] - rep[:, 4:]).abs().max():.4f}")
    assert (y[:, 4:] - rep[:, 4:]).abs().max() < 0.1
    assert (y[:, :4] - rep[:, :4]).abs().median() < 1.0


def test_raw_logits_and_input_dtypes():
    model = _model("n", "survey").to("cuda:0")
    x = synth.synth_images(2, 96, 96, seed=4)
    sd = {k: v.float().cpu() for k, v in model.state_dict().items()}
    with torch.no_grad():
        maps = yolo_oracle.forward_raw(sd, *model._arch, x)
        raw = model.forward_raw(x.to("cuda:0")).cpu()
        y32 = model(x.to("cuda:0")).clone()
        y16 = model(x.to("cuda:0").half()).clone()
        ybf = model(x.to("cuda:0").bfloat16()).clone()
        yu8 = model((x * 255).round().to(torch.uint8).to("cuda:0")).clone()
    ref_rows = yolo_oracle.raw_to_rows(maps)
    assert raw.shape == ref_rows.shape
    assert (raw - ref_rows).abs().max() < 0.05 * max(1.0, ref_rows.abs().max().item())
    assert (y16[:, :4] - y32[:, :4]).abs().max() < 1.0 and (ybf[:, :4] - y32[:, :4]).abs().max() < 2.0
    assert (yu8[:, :4] - y32[:, :4]).abs().max() < 2.0
    assert (y16[:, 4:] - y32[:, 4:]).abs().max() < 1e-2


def test_cuda_graph_replay_equals_stream_launch():
    model = _model("n", "survey")
    x = synth.synth_images(1, 128, 128, seed=5).to("cuda:0")
    eng = Engine(*model._arch, 1, 128, 128, "cuda:0")
    eng.pack_from_model(model)
    a = eng.forward(x).clone()
    eng.use_graph(True)
    b = eng.forward(x).clone()   # capture + first replay
    c = eng.forward(x).clone()   # cached replay
    torch.cuda.synchronize()
    assert torch.equal(a, b) and torch.equal(b, c)


def test_batch_independence():
    """Images are independent (no cross-image op on the path): image i of a batch equals a batch of one."""
    model = _model("n", "survey")
    x = synth.synth_images(3, 64, 64, seed=6).to("cuda:0")
    e3 = Engine(*model._arch, 3, 64, 64, "cuda:0")
    e3.pack_from_model(model)
    e1 = Engine(*model._arch, 1, 64, 64, "cuda:0")
    e1.pack_from_model(model)
    y3 = e3.forward(x).clone()
    for i in range(3):
        assert torch.equal(e1.forward(x[i:i + 1].contiguous())[0], y3[i])
