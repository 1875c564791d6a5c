"""GPU bring-up diagnostic: per-layer error table of the CUDA plan against the bf16-faithful CPU
replay (oracle/plan_replay.py).  Never raises on a mismatch; prints everything it sees.

    python tools/gpu_diag.py --impl direct|tc --size n --hw 64 [--batch 2]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["YB_NO_REUSE"] = "1"

import torch  # noqa: E402

from oracle.plan_replay import PlanReplay  # noqa: E402
from yolo_infer_pt_b200 import synth  # noqa: E402
from yolo_infer_pt_b200.engine import Engine  # noqa: E402
from yolo_infer_pt_b200.nets import nn  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="tc")
    ap.add_argument("--size", default="n")
    ap.add_argument("--hw", type=int, default=64)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--recipe", default="calibrated")
    a = ap.parse_args()
    model = getattr(nn, f"yolo_v11_{a.size}")(80)
    synth.load_synth(model, 0, a.recipe)
    model = model.fuse().eval()
    x = synth.synth_images(a.batch, a.hw, a.hw, seed=1)
    eng = Engine(*model._arch, a.batch, a.hw, a.hw, "cuda:0")
    blob = eng.pack_from_model(model)
    desc = eng.describe()
    with torch.no_grad():
        rep = PlanReplay(desc, eng.convs, blob, emulate_bf16=True)
        ref = rep.run(x)
    taps = {op["name"]: (rep.final_slice(op) if op["kind"] != 3 else None) for op in desc["ops"] if op["kind"] != 5}
    eng.set_conv_impl(1 if a.impl == "direct" else 0)
    try:
        y = eng.forward(x.to("cuda:0"))
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print("FORWARD FAILED:", repr(e))
        return 1
    print(f"impl={a.impl} size={a.size} hw={a.hw}: {len(desc['ops'])} ops")
    bad = 0
    for op in desc["ops"]:
        if op["kind"] == 5:
            continue
        name = op["name"]
        try:
            t = eng.debug_read(name)
        except Exception as e:  # noqa: BLE001
            print(f"{name:42s} debug_read failed: {e}")
            continue
        r0 = op["dst_row_off"]
        if op["kind"] == 3:
            continue
        else:
            got = t[:, r0:r0 + op["Hout"] * op["Wout"]]
            want = taps[name]
        if got.shape != want.shape:
            print(f"{name:42s} shape {tuple(got.shape)} vs {tuple(want.shape)}")
            continue
        err = ((got - want).pow(2).mean().sqrt() / want.pow(2).mean().sqrt().clamp_min(1e-6)).item()
        nan = int(torch.isnan(got).sum())
        flag = "" if err < 0.05 and not nan else "   <<<<<<"
        bad += bool(flag)
        print(f"{name:42s} k{op['k']} s{op['stride']} tma{op['a_tma']} K{op['K_pad']:5d} N{op['N_pad']:4d} "
              f"M{a.batch * op['Hout'] * op['Wout']:7d} rel-rms {err:.4f} nan {nan}{flag}")
    yc = y.cpu()
    print("final: box max-abs", (yc[:, :4] - ref[:, :4]).abs().max().item(), "score max-abs",
          (yc[:, 4:] - ref[:, 4:]).abs().max().item(), "bad layers", bad)
    return 0


if __name__ == "__main__":
    sys.exit(main())
