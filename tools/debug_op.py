"""Teacher-forced run of one op by name; prints where the GPU output deviates from the CPU replay."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
os.environ["YB_NO_REUSE"] = "1"
import torch
from yolo_infer_pt_b200 import synth
from yolo_infer_pt_b200.engine import Engine
from oracle.plan_replay import PlanReplay
from test_gpu_forward import _model
size, hw, target = sys.argv[1], int(sys.argv[2]), sys.argv[3]
model = _model(size, "calibrated")
x = synth.synth_images(2, hw, hw, seed=1)
eng = Engine(*model._arch, 2, hw, hw, "cuda:0")
blob = eng.pack_from_model(model)
desc = eng.describe()
rep = PlanReplay(desc, eng.convs, blob, emulate_bf16=True)
xg = x.to("cuda:0")
with torch.no_grad():
    for i, op in enumerate(desc["ops"]):
        if op["kind"] == 5: break
        touched = {s["buf"] for s in op["src"]} | {op["dst"]["buf"]} | ({op["res"]["buf"]} if op["has_res"] else set())
        if op.get("dw_fused"): touched |= {s["buf"] for s in desc["ops"][i - 1]["src"]}
        before = {b: rep.buffer_bytes(b) for b in touched if b >= 0}
        rep.step(op, x)
        if op["name"] != target: continue
        want = rep.final_slice(op)
        for b, t in before.items(): eng.debug_write(b, t)
        eng.set_conv_impl(0)
        eng.run_op(i, xg)
        got = eng.debug_read(op["name"])
        H, W = op["Hout"], op["Wout"]
        print(op)
        print("shapes", got.shape, want.shape)
        err = (got - want).abs()
        e = err.reshape(2, H, W, -1)
        print("max err", err.max().item(), "want max", want.abs().max().item())
        print("err by row (img0):", [round(v, 2) for v in e[0].amax(dim=(1, 2)).tolist()])
        print("err by col (img0):", [round(v, 2) for v in e[0].amax(dim=(0, 2)).tolist()])
        print("err by channel:", [round(v, 2) for v in e.amax(dim=(0, 1, 2)).tolist()][:80])
        g = got.reshape(2, H, W, -1); w_ = want.reshape(2, H, W, -1)
        print("got[0,5,5,:8]", g[0, 5, 5, :8].tolist()); print("want[0,5,5,:8]", w_[0, 5, 5, :8].tolist())
        import itertools
        for img in (0, 1):
            print("img", img, "err by row:", [round(v, 2) for v in e[img].amax(dim=(1, 2)).tolist()])
        # does a wrong row equal some other (row, col-shift) of the expected output?
        for ry in (4, 5, 6, 7):
            best = None
            for sy, sx in itertools.product(range(H), range(-3, 4)):
                xs = slice(max(0, -sx), min(W, W - sx)); xd = slice(max(0, sx), min(W, W + sx))
                d = (g[0, ry, xs] - w_[0, sy, xd]).abs().max().item()
                if best is None or d < best[0]: best = (d, sy, sx)
            print("row", ry, "best match (err, src row, col shift):", best)
