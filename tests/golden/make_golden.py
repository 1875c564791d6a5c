"""Generates the golden vectors under tests/golden/ by running the REFERENCE ITSELF
(t0saki/YOLO-Infer-pt, imported read-only from /root/reference) on seeded synthetic weights and
inputs.  Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Weights/inputs are regenerated from seeds by yolo_infer_pt_b200/synth.py (numpy Mersenne Twister),
so only the reference's OUTPUTS are stored.  The script also cross-checks oracle/nms_oracle.c
against torchvision.ops.nms (the third-party kernel behind utils/util.py:162) with full_scan=1.
"""
import os
import sys

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REF)   # reference's `nets`, `utils`
sys.path.insert(1, ROOT)  # yolo_infer_pt_b200.synth, oracle

import warnings  # noqa: E402

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torchvision  # noqa: E402
from nets import nn as ref_nn  # noqa: E402
from utils import util as ref_util  # noqa: E402

from oracle import nms_oracle  # noqa: E402
from yolo_infer_pt_b200 import synth  # noqa: E402

warnings.filterwarnings("ignore")
assert ref_nn.__file__.startswith(REF) and ref_util.__file__.startswith(REF)
ref_util.time = lambda: 0.0  # neutralise the wall-clock bail-out, utils/util.py:133-134,166-167
torch.set_num_threads(8)


def ref_model(size, nc=80, seed=0, recipe="calibrated"):
    m = getattr(ref_nn, f"yolo_v11_{size}")(nc)
    m.load_state_dict(synth.synth_state_dict(m, seed, recipe))
    return m


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print(f"wrote {name}: " + ", ".join(f"{k}{tuple(v.shape)}" for k, v in arrays.items()),
          f"{os.path.getsize(path) / 1024:.0f} KB")


def main():
    # ---- forward, every size, 2 x 3 x 64 x 64 -------------------------------------------------
    for size in ("" if os.environ.get("YB_GOLDEN_ONLY_NEW") else "ntsmlx"):
        m = ref_model(size)
        x = synth.synth_images(2, 64, 64, seed=1)
        with torch.no_grad():
            unfused = m.eval()(x)
            m.head.training = True          # raw maps with eval-mode BatchNorm
            raw = [t.clone() for t in m(x)]
            m.head.training = False
            fused = m.fuse().eval()(x)
        save(f"fwd_{size}_64.npz", out=fused.numpy(), out_unfused=unfused.numpy(),
             raw0=raw[0].numpy(), raw1=raw[1].numpy(), raw2=raw[2].numpy())
    # ---- forward on the SURVEY §8(d) recipe (the tolerance gate's weights) -----------------------
    for size in ("" if os.environ.get("YB_GOLDEN_ONLY_NEW") else "nx"):
        m = ref_model(size, recipe="survey").fuse().eval()
        x = synth.synth_images(2, 64, 64, seed=1)
        with torch.no_grad():
            y = m(x)
        save(f"fwdsv_{size}_64.npz", out=y.numpy())
    # ---- forward, n and s at 640 (BASELINE config 1), sub-sampled anchors ----------------------
    for size in ("" if os.environ.get("YB_GOLDEN_ONLY_NEW") else "ns"):
        m = ref_model(size, recipe="survey").fuse().eval()
        x = synth.synth_images(1, 640, 640, seed=0)
        with torch.no_grad():
            y = m(x)
        idx = np.arange(0, 8400, 16)
        frac = float((y[0, 4:] > 0.001).float().mean())
        save(f"fwd_{size}_640.npz", out_sub=y[:, :, idx].numpy(), idx=idx,
             stats=np.array([frac, float(y[0, 4:].max()), float(y[0, :4].abs().mean())], dtype=np.float64))
        if size == "n":
            dets = ref_util.non_max_suppression(y, 0.001, 0.65)
            save("e2e_n_640_nms.npz", det=dets[0].numpy())
    # ---- the falsifiable gate's recipes (tests/test_gpu_parity.py): scores spread over (0, 1) ----------
    for size, hw in (("n", 320), ("x", 128)):
        m = ref_model(size, recipe="survey_widehead").fuse().eval()
        x = synth.synth_images(2, hw, hw, seed=0)
        with torch.no_grad():
            y = m(x)
        idx = np.arange(0, y.shape[2], 4)
        sc = y[:, 4:]
        save(f"fwdwh_{size}_{hw}.npz", out_sub=y[:, :, idx].numpy(), idx=idx,
             stats=np.array([float(((sc > 0.1) & (sc < 0.9)).float().mean()), float(sc.max())], dtype=np.float64))
    m = ref_model("n", recipe="calibrated_damped").fuse().eval()
    with torch.no_grad():
        y = m(synth.synth_images(1, 640, 640, seed=0))
    dets = ref_util.non_max_suppression(y, 0.25, 0.65)
    save("e2e_cd_n_640_nms.npz", det=dets[0].numpy(), conf=np.float64(0.25), iou=np.float64(0.65),
         ncand=np.array(int((y[0, 4:] > 0.25).sum())))
    if os.environ.get("YB_GOLDEN_ONLY_NEW"):
        return
    # ---- non_max_suppression -------------------------------------------------------------------
    cases = [
        ("sparse_640", dict(batch=2, nc=80, anchors=8400, img=640, mode="sparse", seed=0), 0.001, 0.65),
        ("sparse_640_iou07", dict(batch=2, nc=80, anchors=8400, img=640, mode="sparse", seed=1), 0.001, 0.7),
        ("dense_640", dict(batch=1, nc=80, anchors=8400, img=640, mode="dense", seed=2), 0.001, 0.7),
        ("sparse_1280", dict(batch=1, nc=80, anchors=33600, img=1280, mode="sparse", seed=3), 0.001, 0.7),
        ("nc1_320", dict(batch=2, nc=1, anchors=2100, img=320, mode="dense", seed=4), 0.25, 0.65),
        ("conf25_640", dict(batch=2, nc=80, anchors=8400, img=640, mode="sparse", seed=5), 0.25, 0.45),
        ("few_640", dict(batch=3, nc=80, anchors=8400, img=640, mode="few", seed=8), 0.001, 0.65),
        ("empty", dict(batch=2, nc=80, anchors=8400, img=640, mode="empty", seed=6), 0.001, 0.65),
    ]
    for name, kw, conf, iou in cases:
        pred = synth.synth_predictions(**kw)
        ref = ref_util.non_max_suppression(torch.from_numpy(pred), conf, iou)
        ora = nms_oracle.non_max_suppression(pred, conf, iou)
        for r, o in zip(ref, ora):
            assert np.array_equal(r.numpy(), o), f"oracle != reference on {name}"
        counts = np.array([len(r) for r in ref], dtype=np.int32)
        det = np.zeros((len(ref), 300, 6), dtype=np.float32)
        for i, r in enumerate(ref):
            det[i, :len(r)] = r.numpy()
        ncand = np.array([(pred[b, 4:] > np.float32(conf)).sum() for b in range(pred.shape[0])])
        save(f"nms_{name}.npz", det=det, counts=counts, conf=np.float64(conf), iou=np.float64(iou),
             ncand=ncand, **{f"kw_{k}": np.array(v) for k, v in kw.items()})
        print("   candidates per image", ncand.tolist(), "kept", counts.tolist())
    # ---- greedy step vs torchvision, full scan, incl. edge cases --------------------------------
    rng = np.random.RandomState(7)
    for trial in range(20):
        n = 2000
        cxy = rng.uniform(0, 200, (n, 2)).astype(np.float32)
        wh = rng.uniform(1, 80, (n, 2)).astype(np.float32)
        if trial % 4 == 0:
            wh[::7] = 0  # degenerate boxes -> NaN IoU -> kept
        sc = (rng.permutation(n).astype(np.float32) + 1) / (n + 1)
        pred = np.concatenate([cxy.T, wh.T, sc[None]], 0)[None]  # (1, 5, n), nc = 1
        thr = [0.65, 0.7, 0.5, 1.0 / 3.0][trial % 4]
        boxes = torch.from_numpy(ref_util.wh2xy(np.concatenate([cxy, wh], 1)))
        keep = torchvision.ops.nms(boxes, torch.from_numpy(sc), thr).numpy()
        ora = nms_oracle.non_max_suppression(pred, 0.0, thr, max_det=n, max_nms=n, full_scan=True)[0]
        exp = np.concatenate([boxes.numpy()[keep], sc[keep, None], np.zeros((len(keep), 1), np.float32)], 1)
        assert np.array_equal(ora, exp), f"oracle greedy != torchvision (trial {trial})"
    print("oracle greedy step == torchvision.ops.nms on 20 random trials (incl. degenerate boxes, thr=1/3)")


if __name__ == "__main__":
    main()
