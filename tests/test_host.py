"""Host-logic tests (CPU only): C-ABI surface, plan builder, buffer aliasing, weight packer, and the
Python drop-in boundary.  The plan is replayed on the CPU by oracle/plan_replay.py (test
infrastructure) and compared with the oracle forward; no CUDA kernel runs here."""
import copy
import io
import os
import pickle
import re

import numpy as np
import pytest
import torch

from oracle import yolo_oracle
from oracle.plan_replay import PlanReplay
from yolo_infer_pt_b200 import _lib, synth
from yolo_infer_pt_b200.engine import Engine
from yolo_infer_pt_b200.nets import nn
from yolo_infer_pt_b200.utils import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "yolob200.h")).read()
    declared = set(re.findall(r"\b(yb_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = _lib.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/yolob200.h but not exported"
    assert declared == set(_lib.SYMBOLS), "ctypes table and header disagree"
    assert lib.yb_version() >= 100


def test_no_cpu_fallback():
    m = nn.yolo_v11_n(80).eval()
    with pytest.raises(RuntimeError, match="no CPU"):
        m(torch.zeros(1, 3, 64, 64))
    with pytest.raises(RuntimeError, match="no CPU"):
        util.non_max_suppression(torch.zeros(1, 84, 84))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CUDA device|no CPU"):
            Engine(*nn.yolo_v11_n(80)._arch, 1, 64, 64, device="cuda:0")


def test_training_mode_and_state_dict_contract():
    m = nn.yolo_v11_n(80)
    assert m.training
    maps = m(torch.zeros(1, 3, 64, 64))
    assert [tuple(t.shape) for t in maps] == [(1, 144, 8, 8), (1, 144, 4, 4), (1, 144, 2, 2)]
    sd = m.state_dict()
    assert len(sd) == 499  # 499 keys unfused for n
    for k in ("net.p1.0.conv.weight", "net.p1.0.norm.running_mean", "head.box.0.2.bias", "head.dfl.conv.weight",
              "net.p5.3.res_m.0.conv1.qkv.conv.weight", "fpn.h6.res_m.0.res_m.1.conv2.conv.weight"):
        assert k in sd
    assert torch.equal(m.stride, torch.tensor([8.0, 16.0, 32.0]))
    assert m.head.nc == 80 and m.head.no == 144 and m.head.ch == 16 and m.head.nl == 3
    m.fuse()
    assert len(m.state_dict()) == 175  # 175 keys fused
    assert not hasattr(m.net.p1[0], "norm")
    # pickling / deepcopy must not drag engine handles along
    m2 = pickle.load(io.BytesIO(pickle.dumps(m)))
    m3 = copy.deepcopy(m)
    for a, b in zip(m.state_dict().values(), m2.state_dict().values()):
        assert torch.equal(a, b)
    assert len(m3.state_dict()) == 175


def test_fuse_matches_batchnorm():
    torch.manual_seed(0)
    c = nn.Conv(8, 16, torch.nn.SiLU(), k=3, p=1)
    c.norm.running_mean.normal_()
    c.norm.running_var.uniform_(0.5, 2.0)
    c.norm.weight.data.uniform_(0.5, 1.5)
    c.norm.bias.data.normal_()
    c.eval()
    x = torch.randn(2, 8, 12, 12)
    ref = c(x)
    fused = nn.fuse_conv(c.conv, c.norm)
    assert (torch.nn.functional.silu(fused(x)) - ref).abs().max() < 1e-5


def bf16_weight_state(model, dtype=torch.bfloat16):
    """Fused state_dict whose dense-conv weights are rounded to the plan's 16-bit storage type — what the
    packed blob holds (stem and depthwise weights stay fp32 in the blob)."""
    m = copy.deepcopy(model).fuse()
    sd = {k: v.float().clone() for k, v in m.state_dict().items()}
    for name, mod in m.named_modules():
        if isinstance(mod, torch.nn.Conv2d) and mod.groups == 1 and name not in ("net.p1.0.conv", "head.dfl.conv"):
            sd[name + ".weight"] = sd[name + ".weight"].to(dtype).float()
    return sd


@pytest.mark.parametrize("size,hw,act", [("n", 64, torch.float16), ("t", 64, torch.float16), ("s", 64, torch.float16),
                                         ("m", 64, torch.float16), ("x", 64, torch.bfloat16), ("n", 96, torch.bfloat16),
                                         ("deep", 64, torch.float16)])
def test_plan_replay_matches_oracle(size, hw, act):
    """Plan + aliasing + packed weights, replayed in fp32 on the CPU, reproduce the oracle forward
    run on the same 16-bit-rounded weights — layer by layer and end to end, for both storage types.
    ("deep": a C3k2 with two plain bottlenecks - only the last one's residual is folded into conv2's weights.)"""
    if size == "deep":
        model = nn.YOLO([3, 16, 32, 64, 128, 256], [2, 2, 1, 1, 1, 2], [False, True], 80)
    else:
        model = getattr(nn, f"yolo_v11_{size}")(80)
    synth.load_synth(model, 0)
    eng = Engine(*model._arch, 2, hw, hw, host_only=True, act_dtype=act)
    blob = eng.pack_from_model(model)
    desc = eng.describe()
    x = synth.synth_images(2, hw, hw, seed=1)
    taps_o, taps_r = {}, {}
    sd = bf16_weight_state(model, act)
    fused_sd = copy.deepcopy(model).fuse().state_dict()
    for o in desc["ops"]:   # folded residuals: the blob holds round(W_x + W_m) for x's columns - give the oracle the same sum
        if o["kind"] == 1 and o["wfold"][2]:
            d0, s0, c = o["wfold"]
            key = o["name"] + ".conv.weight"
            w = fused_sd[key].float()
            sd[key][:, d0:d0 + c] = (w[:, d0:d0 + c] + w[:, s0:s0 + c]).to(act).float() - sd[key][:, s0:s0 + c]
    with torch.no_grad():
        ref = yolo_oracle.forward(sd, *model._arch, x, taps=taps_o)
        rep = PlanReplay(desc, eng.convs, blob, emulate_bf16=False)
        out = rep.run(x, taps=taps_r)
    assert not torch.isnan(out).any(), "an op read channels nobody wrote"
    checked = 0
    # C3k2's last bottleneck stores f(x) alone; its `x +` lives in the consumer's packed weights (plan.cu): the
    # reference tensor to compare with is oracle - x
    folded = {}
    for o in desc["ops"]:
        if o["kind"] == 1 and o["wfold"][2]:
            base, c = o["name"][:-len(".conv2")], o["wfold"][2]
            n = o["wfold"][0] // c
            folded[f"{base}.res_m.{n - 1}.conv2"] = (f"{base}.conv1", c) if n == 1 else (f"{base}.res_m.{n - 2}.conv2", 0)
    assert folded or size not in ("n", "s", "deep")
    if size == "deep":
        assert folded["net.p2.1.res_m.1.conv2"] == ("net.p2.1.res_m.0.conv2", 0)
    for name in [o["name"] for o in desc["ops"] if o["kind"] != 5]:
        if name not in taps_o:
            continue
        a = taps_r[name]                                             # (B, rows, C)
        b = taps_o[name].permute(0, 2, 3, 1).reshape(a.shape[0], -1, taps_o[name].shape[1])
        if name in folded:
            src, off = folded[name]
            xin = taps_o[src].permute(0, 2, 3, 1).reshape(a.shape[0], -1, taps_o[src].shape[1])
            b = b - xin[..., off:off + b.shape[-1]]
        assert a.shape == b.shape, name
        err = (a - b).abs().max().item()
        assert err <= 2e-3 * max(1.0, b.abs().max().item()), f"{name}: max err {err}"
        checked += 1
    assert checked >= 80
    assert (out[:, :4] - ref[:, :4]).abs().max() < 0.2   # fp32 summation-order noise on ~700 px boxes
    assert (out[:, 4:] - ref[:, 4:]).abs().max() < 2e-3


def test_conv_names_resolve_to_modules_of_every_size():
    for size in "ntsmlx":
        model = getattr(nn, f"yolo_v11_{size}")(80)
        eng = Engine(*model._arch, 1, 64, 64, host_only=True)
        seen = set()
        for c in eng.convs:
            mod = model.get_submodule(c["name"])
            conv = mod.conv if c["wrapped"] else mod
            assert tuple(conv.weight.shape) == (c["cout"], c["cin"], c["ksize"], c["ksize"]), c["name"]
            assert conv.groups == c["groups"] and conv.stride[0] == c["stride"], c["name"]
            seen.add(c["name"])
        convs_in_model = {n for n, m in model.named_modules()
                          if isinstance(m, torch.nn.Conv2d) and not n.startswith("head.dfl")}
        convs_in_model = {n[:-5] if n.endswith(".conv") else n for n in convs_in_model}
        assert seen == convs_in_model, f"{size}: plan misses {convs_in_model - seen} / extra {seen - convs_in_model}"


@pytest.mark.parametrize("size,batch,hw", [("n", 256, 640), ("x", 64, 640), ("s", 1, 1280)])
def test_arena_reuse_never_overlaps_live_buffers(size, batch, hw):
    model_arch = getattr(nn, f"yolo_v11_{size}")(80)._arch
    eng = Engine(*model_arch, batch, hw, hw, host_only=True)
    d = eng.describe()
    # recompute liveness from the op list
    first, last = {}, {}
    for i, op in enumerate(d["ops"]):
        slices = list(op["src"]) + ([op["dst"]] if op["kind"] != 5 else []) + ([op["res"]] if op["has_res"] else [])
        if op["kind"] == 5:
            slices.append({"buf": d["logits_buf"]})
        for s in slices:
            if s["buf"] < 0:
                continue
            first[s["buf"]] = min(first.get(s["buf"], i), i)
            last[s["buf"]] = max(last.get(s["buf"], i), i)
    bufs = d["bufs"]
    total = 0
    for i, a in enumerate(bufs):
        assert a["first_def"] <= first[i] and a["last_use"] >= last[i], a["tag"]
        assert a["offset"] % 1024 == 0 and a["offset"] + a["bytes"] <= d["workspace_bytes"]
        total += a["bytes"]
        for j in range(i):
            b = bufs[j]
            mem_overlap = a["offset"] < b["offset"] + b["bytes"] and b["offset"] < a["offset"] + a["bytes"]
            live_overlap = not (a["last_use"] < b["first_def"] or b["last_use"] < a["first_def"])
            assert not (mem_overlap and live_overlap), f"{a['tag']} and {b['tag']} alias while both live"
    assert d["workspace_bytes"] < total  # reuse actually happens
    assert d["workspace_bytes"] < 170e9


def test_unsupported_configs_fail_loudly():
    with pytest.raises(RuntimeError, match="multiple of 32"):
        Engine(*nn.yolo_v11_n(80)._arch, 1, 100, 100, host_only=True)
    with pytest.raises(RuntimeError):
        Engine([3, 16, 32, 64, 128, 200], [1] * 6, [False, True], 80, 1, 64, 64, host_only=True)


def test_c_abi_rejects_bad_arguments_without_touching_the_gpu():
    """Error behaviour of the entry points added around the candidate sink: negative code + message, no crash."""
    import ctypes
    from yolo_infer_pt_b200 import _lib
    L = _lib.lib()
    e = Engine(*nn.yolo_v11_n(80)._arch, 2, 64, 64, host_only=True)
    null = ctypes.c_void_p(0)
    # no workspace / no plan
    assert L.yb_forward_nms(e.plan, null, 3, null, ctypes.c_float(0.001), 30000, null, 0, null) < 0
    assert L.yb_forward_nms(null, null, 3, null, ctypes.c_float(0.001), 30000, null, 0, null) < 0
    assert b"yb_forward_nms" in L.yb_last_error()
    assert L.yb_nms_prefiltered(null, 2, 80, 84, ctypes.c_float(0.001), 0.65, 300, 30000, ctypes.c_float(7680.0), null, null,
                                null, 0, null) < 0
    # workspace size is a pure function of the sizes (same formula the sink uses)
    assert L.yb_nms_workspace_bytes(2, 80, 84, 30000) >= 2 * 32 + 2 * 2048 * 4 + 2 * 32768 * 8


def test_wh2xy_and_make_anchors():
    x = torch.tensor([[10.0, 20.0, 4.0, 6.0]])
    assert torch.equal(util.wh2xy(x), torch.tensor([[8.0, 17.0, 12.0, 23.0]]))
    assert np.array_equal(util.wh2xy(x.numpy()), np.array([[8.0, 17.0, 12.0, 23.0]], dtype=np.float32))
    maps = [torch.zeros(1, 1, 2, 3), torch.zeros(1, 1, 1, 1)]
    pts, st = util.make_anchors(maps, [8, 16])
    assert pts.tolist() == [[0.5, 0.5], [1.5, 0.5], [2.5, 0.5], [0.5, 1.5], [1.5, 1.5], [2.5, 1.5], [0.5, 0.5]]
    assert st.view(-1).tolist() == [8.0] * 6 + [16.0]


def test_reference_checkpoint_imports(tmp_path):
    """a checkpoint written by the reference itself ({'model': module.half()}, what
    strip_optimizer leaves in weights/best.pt, util.py:332-337) loads (a) through `load_weight` and (b) by
    plain torch.load with our `nets.nn` standing in for the reference's - `_arch` is inferred from the
    module tree.  Needs /root/reference (present in the build container only)."""
    import subprocess, sys, textwrap
    if not os.path.isdir("/root/reference/nets"):
        pytest.skip("reference sources not present")
    ck = str(tmp_path / "best.pt")
    code = textwrap.dedent(f"""
        import sys, torch
        sys.path.insert(0, '/root/reference')
        from nets import nn
        torch.manual_seed(3)
        m = nn.yolo_v11_s(7)
        for p in m.parameters():
            p.data.normal_(0, 0.05)
        torch.save({{'model': m.half()}}, {ck!r})
        torch.save({{k: v.float() for k, v in m.state_dict().items()}}, {ck!r} + '.sd')
    """)
    subprocess.run([sys.executable, "-c", code], check=True)
    pkg = os.path.join(ROOT, "yolo_infer_pt_b200")
    code2 = textwrap.dedent(f"""
        import sys, torch
        sys.path.insert(0, {pkg!r}); sys.path.insert(0, {ROOT!r})
        from nets import nn
        from utils import util
        sd = torch.load({ck!r} + '.sd')
        m = util.load_weight(nn.yolo_v11_s(7), {ck!r})
        assert all(torch.equal(v, sd[k]) for k, v in m.state_dict().items()), 'load_weight'
        m2 = torch.load({ck!r}, map_location='cpu', weights_only=False)['model'].float().fuse()
        assert type(m2).__module__ == 'nets.nn' and m2._arch == nn.yolo_v11_s(7)._arch, m2._arch
        assert len(m2.state_dict()) == len(nn.yolo_v11_s(7).fuse().state_dict())
        print('ok')
    """)
    out = subprocess.run([sys.executable, "-c", code2], check=True, capture_output=True, text=True)
    assert "ok" in out.stdout


def test_weights_fingerprint_sees_every_kind_of_update():
    """YOLO._weights_version (host logic of the engine cache): in-place updates, re-assignment, submodule dtype
    moves and submodule load_state_dict(assign=True) all change it; reading it does not."""
    m = nn.yolo_v11_n(80).fuse().eval()
    v = m._weights_version()
    assert v == m._weights_version()
    p = m.head.box[0][2].bias
    with torch.no_grad():
        p.add_(1)
    v1 = m._weights_version()
    assert v1 != v
    p.data = p.data.clone()
    v2 = m._weights_version()
    assert v2 != v1
    m.net.half()
    v3 = m._weights_version()
    assert v3 != v2
    m.head.load_state_dict(m.head.state_dict(), assign=True)
    assert m._weights_version() != v3


def test_widehead_recipe_spreads_scores_and_boxes():
    """The falsifiable gate's recipe really has scores across (0, 1) and DFL expectations across bins - checked on
    the fp32 oracle and against the fixture written from the reference itself."""
    m = nn.yolo_v11_n(80)
    synth.load_synth(m, 0, "survey_widehead")
    m = m.fuse().eval()
    sd = {k: v.float() for k, v in m.state_dict().items()}
    x = synth.synth_images(2, 320, 320, seed=0)
    with torch.no_grad():
        maps = yolo_oracle.forward_raw(sd, *m._arch, x)
        y = yolo_oracle.decode(maps, 80)
    sc = y[:, 4:]
    assert ((sc > 0.1) & (sc < 0.9)).float().mean() >= 0.05 and sc.max() > 0.5
    dist = torch.cat([(t[:, :64].reshape(2, 4, 16, -1).softmax(2) * torch.arange(16.0).view(1, 1, 16, 1)).sum(2).flatten()
                      for t in maps])
    assert dist.std() >= 2.0
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "fwdwh_n_320.npz"))
    ref = torch.from_numpy(g["out_sub"])
    sub = y[:, :, torch.from_numpy(g["idx"])]
    assert (sub[:, :4] - ref[:, :4]).abs().max() < 2e-2 and (sub[:, 4:] - ref[:, 4:]).abs().max() < 1e-5


def _to_ultralytics_keys(model):
    """Inverse of util.ultralytics_key, written independently from the Ultralytics module layout (yolo11.yaml +
    ultralytics/nn/modules: Conv{conv,bn}, C3k2{cv1,cv2,m}, C3k{cv1,cv2,cv3,m}, Bottleneck{cv1,cv2}, SPPF{cv1,cv2},
    C2PSA{cv1,cv2,m[PSABlock{attn{qkv,proj,pe},ffn}]}, Detect{cv2,cv3,dfl})."""
    layer = {"net.p1.0": 0, "net.p2.0": 1, "net.p2.1": 2, "net.p3.0": 3, "net.p3.1": 4, "net.p4.0": 5, "net.p4.1": 6,
             "net.p5.0": 7, "net.p5.1": 8, "net.p5.2": 9, "net.p5.3": 10, "fpn.h1": 13, "fpn.h2": 16, "fpn.h3": 17,
             "fpn.h4": 19, "fpn.h5": 20, "fpn.h6": 22}
    out = {}
    for key in model.state_dict():
        p = key.split(".")
        if p[0] == "head":
            if p[1] == "dfl":
                u = "model.23." + ".".join(p[1:])
            elif p[1] == "box":
                u = f"model.23.cv2.{p[2]}.{p[3]}." + ".".join(p[4:]).replace("norm", "bn")
            else:
                j = int(p[3])
                u = (f"model.23.cv3.{p[2]}.2." + ".".join(p[4:])) if j == 4 else \
                    (f"model.23.cv3.{p[2]}.{j // 2}.{j % 2}." + ".".join(p[4:]).replace("norm", "bn"))
        else:
            pre = ".".join(p[:3]) if p[0] == "net" else ".".join(p[:2])
            rest = p[3:] if p[0] == "net" else p[2:]
            if pre == "net.p5.3" and rest[0] == "res_m":      # PSABlock
                blk, sub = rest[1], rest[2:]
                if sub[0] == "conv1":
                    name = {"qkv": "qkv", "conv2": "proj", "conv1": "pe"}[sub[1]]
                    r = ["m", blk, "attn", name] + sub[2:]
                else:
                    r = ["m", blk, "ffn"] + sub[1:]
            else:
                r = [{"conv1": "cv1", "conv2": "cv2", "conv3": "cv3", "res_m": "m"}.get(t, t) for t in rest]
            u = f"model.{layer[pre]}." + ".".join("bn" if t == "norm" else t for t in r)
        out[u] = key
    return out


@pytest.mark.parametrize("size", ["n", "s", "m", "x"])
def test_ultralytics_key_map_hits_every_key(size):
    """Corrected Ultralytics -> nets.nn key map (reference util.py:358-516 drops the head, the nested C3k blocks,
    C2PSA and most BatchNorm statistics): every tensor of the model is reached exactly once, for every size."""
    model = getattr(nn, f"yolo_v11_{size}")(80)
    inv = _to_ultralytics_keys(model)
    assert len(inv) == len(model.state_dict()) and (size != "n" or len(inv) == 499)
    mapped = {u: util.ultralytics_key(u) for u in inv}
    assert mapped == inv
    # names as they appear in a real yolo11n.pt, with the shapes its tensors have
    if size == "n":
        sd = model.state_dict()
        known = {"model.0.conv.weight": ("net.p1.0.conv.weight", (16, 3, 3, 3)),
                 "model.2.m.0.cv1.conv.weight": ("net.p2.1.res_m.0.conv1.conv.weight", (8, 16, 3, 3)),
                 "model.6.m.0.m.1.cv2.bn.running_var": ("net.p4.1.res_m.0.res_m.1.conv2.norm.running_var", (32,)),
                 "model.9.cv2.conv.weight": ("net.p5.2.conv2.conv.weight", (256, 512, 1, 1)),
                 "model.10.m.0.attn.qkv.conv.weight": ("net.p5.3.res_m.0.conv1.qkv.conv.weight", (256, 128, 1, 1)),
                 "model.10.m.0.attn.pe.conv.weight": ("net.p5.3.res_m.0.conv1.conv1.conv.weight", (128, 1, 3, 3)),
                 "model.10.m.0.attn.proj.bn.bias": ("net.p5.3.res_m.0.conv1.conv2.norm.bias", (128,)),
                 "model.10.m.0.ffn.1.conv.weight": ("net.p5.3.res_m.0.conv2.1.conv.weight", (128, 256, 1, 1)),
                 "model.22.m.0.cv3.conv.weight": ("fpn.h6.res_m.0.conv3.conv.weight", (128, 128, 1, 1)),
                 "model.23.cv2.0.2.weight": ("head.box.0.2.weight", (64, 64, 1, 1)),
                 "model.23.cv2.2.0.conv.weight": ("head.box.2.0.conv.weight", (64, 256, 3, 3)),
                 "model.23.cv3.0.0.0.conv.weight": ("head.cls.0.0.conv.weight", (64, 1, 3, 3)),
                 "model.23.cv3.1.1.1.bn.weight": ("head.cls.1.3.norm.weight", (80,)),
                 "model.23.cv3.2.2.bias": ("head.cls.2.4.bias", (80,)),
                 "model.23.dfl.conv.weight": ("head.dfl.conv.weight", (1, 16, 1, 1))}
        for u, (ours, shape) in known.items():
            assert util.ultralytics_key(u) == ours and tuple(sd[ours].shape) == shape, u
    # loading an Ultralytics-named state_dict reproduces the donor model exactly; a checkpoint that lacks a branch fails loudly
    donor = getattr(nn, f"yolo_v11_{size}")(80)
    synth.load_synth(donor, 0, "survey")
    usd = {u: donor.state_dict()[k].clone() for u, k in inv.items()}
    util.load_ultralytics_weight(model, {"model": usd})
    for k, v in donor.state_dict().items():
        assert torch.equal(model.state_dict()[k], v), k
    broken = {u: v for u, v in usd.items() if ".cv3." not in u}
    with pytest.raises(KeyError):
        util.load_ultralytics_weight(getattr(nn, f"yolo_v11_{size}")(80), {"model": broken})
    assert util.ultralytics_key("model.11.foo") is None and util.ultralytics_key("model.23.cv4.0.0") is None


@pytest.mark.parametrize("size,batch,hw", [("n", 256, 640), ("n", 2, 64), ("s", 3, 320), ("x", 8, 640), ("m", 1, 1280)])
def test_arena_never_overlaps_live_buffers(size, batch, hw):
    """Lifetime-based arena reuse: two buffers whose lifetimes intersect never share memory.  Lifetimes are
    recomputed here from the op list, INCLUDING the reads a fused kernel makes on behalf of an op that is no
    longer launched (a 1x1 conv with the depthwise conv in front fused in reads that conv's input: round 1
    placed the 1x1's own output on top of it, a cross-CTA race that showed as rare 1-ulp score differences
    between identical images of a batch)."""
    m = getattr(nn, f"yolo_v11_{size}")(80)
    d = Engine(*m._arch, batch, hw, hw, host_only=True).describe()
    ops, bufs = d["ops"], d["bufs"]
    live = {}
    for i, op in enumerate(ops):
        used = [s["buf"] for s in op["src"]] + [op["dst"]["buf"]] + ([op["res"]["buf"]] if op["has_res"] else [])
        if op.get("dw_fused"):
            assert ops[i - 1]["fused_away"]
            used += [s["buf"] for s in ops[i - 1]["src"]]
        for b in used:
            if b >= 0:
                lo, hi = live.get(b, (10 ** 9, -1))
                live[b] = (min(lo, i), max(hi, i))
    assert any(op.get("dw_fused") for op in ops) or hw < 160 or size == "x"   # (x: 384-channel class branch is not fused)
    ids = sorted(live)
    for a in ids:
        for b in ids:
            if a < b and not (live[a][1] < live[b][0] or live[b][1] < live[a][0]):
                A, Bb = bufs[a], bufs[b]
                assert not (A["offset"] < Bb["offset"] + Bb["bytes"] and Bb["offset"] < A["offset"] + A["bytes"]), \
                    f"{A['tag']} {live[a]} and {Bb['tag']} {live[b]} overlap in the arena"
        assert bufs[a]["offset"] + bufs[a]["bytes"] <= d["workspace_bytes"]



@pytest.mark.parametrize("size,batch,hw,lanes", [("n", 256, 640, 4), ("x", 8, 640, 3), ("s", 3, 320, 2), ("m", 1, 64, 4)])
def test_stream_lanes_order_every_shared_arena_region(size, batch, hw, lanes, monkeypatch):
    """YB_LANES > 1: ops run on several streams.  Recompute 'complete before' from the plan (stream order inside a
    lane + the cross-lane event waits, transitive) and check (1) every data dependency between two ops is ordered,
    (2) two buffers that share arena memory have ALL their accesses ordered, not just disjoint op-index intervals."""
    monkeypatch.setenv("YB_LANES", str(lanes))
    m = getattr(nn, f"yolo_v11_{size}")(80)
    d = Engine(*m._arch, batch, hw, hw, host_only=True).describe()
    ops, bufs = d["ops"], d["bufs"]
    assert d["num_lanes"] == lanes and len({op["lane"] for op in ops}) == lanes
    n = len(ops)
    reach = [set() for _ in range(n)]
    tail = {}
    for i, op in enumerate(ops):
        for j in [tail.get(op["lane"])] + list(op["xdeps"]):
            if j is not None:
                assert j < i and (j == tail.get(op["lane"]) or ops[j]["signal"])
                reach[i] |= reach[j] | {j}
        tail[op["lane"]] = i

    def accesses(i):
        op = ops[i]
        rd = [s for s in op["src"]] + ([op["res"]] if op["has_res"] else [])
        if op.get("dw_fused"):
            rd += ops[i - 1]["src"]
        if op["kind"] == 2 and op["dw"][3]:
            rd.append(op["dst"])
        if op["kind"] == 5:
            return [(d["logits_buf"], 0, 10 ** 6, -1, False)]
        out = [(s["buf"], s["c_off"], s["c_off"] + s["C"], -1, False) for s in rd if s["buf"] >= 0]
        w = op["dst"]
        out.append((w["buf"], w["c_off"], w["c_off"] + w["C"], op["dst_row_off"] if op["out_f32"] else -1, True))
        return out

    acc = [accesses(i) for i in range(n)]
    touched = {}
    for i in range(n):
        for a in acc[i]:
            touched.setdefault(a[0], []).append(i)
        for j in range(i):
            for a in acc[i]:
                for b in acc[j]:
                    if a[0] != b[0] or not (a[4] or b[4]) or a[2] <= b[1] or b[2] <= a[1]:
                        continue
                    if a[4] and b[4] and a[3] >= 0 and b[3] >= 0 and a[3] != b[3]:
                        continue
                    assert j in reach[i], f"{ops[j]['name']} -> {ops[i]['name']} is not ordered"
    ids = sorted(touched)
    for a in ids:
        for b in ids:
            A, Bb = bufs[a], bufs[b]
            if a < b and A["offset"] < Bb["offset"] + Bb["bytes"] and Bb["offset"] < A["offset"] + A["bytes"]:
                first, second = (a, b) if max(touched[a]) < min(touched[b]) else (b, a)
                for y in touched[second]:
                    for x in touched[first]:
                        assert x in reach[y], f"{bufs[first]['tag']} (op {x}) / {bufs[second]['tag']} (op {y}) share memory unordered"


def test_plan_decisions_for_the_bench_workload(monkeypatch):
    """The plan-level choices the measured numbers rest on (YOLO11n, B = 256, 640 x 640): six depthwise convs fused
    into their 1x1 consumers, five C3k2 residuals folded into the consumer's weights (and only where the block has
    plain bottlenecks), two space-to-depth sources, one stream, 83 launches per forward (79 of them the tcgen05 conv kernel)."""
    for v in ("YB_LANES", "YB_NO_S2D", "YB_NO_RES_FOLD", "YB_NO_DWFUSE", "YB_NO_PATCH"):
        monkeypatch.delenv(v, raising=False)
    eng = Engine(*nn.yolo_v11_n(80)._arch, 256, 640, 640, host_only=True)
    d = eng.describe()
    ops = d["ops"]
    assert sum(o["fused_away"] for o in ops) == 6 == sum(o["dw_fused"] for o in ops)
    folded = [o["name"] for o in ops if o["wfold"][2]]
    assert folded == ["net.p2.1.conv2", "net.p3.1.conv2", "fpn.h1.conv2", "fpn.h2.conv2", "fpn.h4.conv2"]
    for o in ops:   # a bottleneck whose add was folded has no residual operand; every other Residual keeps it
        if o["name"].endswith(".res_m.0.conv2") and o["name"].rsplit(".res_m.0.conv2", 1)[0] + ".conv2" in folded:
            assert not o["has_res"], o["name"]
        elif ".res_m." in o["name"] and o["name"].endswith(".conv2") and o["k"] == 3:
            assert o["has_res"], o["name"]
    assert [o["name"] for o in ops if o["s2d"]] == ["net.p2.0", "net.p3.0"]
    assert d["num_lanes"] == 1 and eng.num_launches == 83
    assert sum(1 for o in ops if o["kind"] == 1 and not o["fused_away"]) == 79
    assert d["workspace_bytes"] < 3.5e9
    # the x architecture uses C3k blocks everywhere: nothing to fold, nothing stored space-to-depth
    dx = Engine(*nn.yolo_v11_x(80)._arch, 8, 640, 640, host_only=True).describe()
    assert not any(o["wfold"][2] or o["s2d"] for o in dx["ops"])


def test_stream_lanes_default_follows_the_batch_size(monkeypatch):
    """Lanes are a latency tool: on by default up to 16 images of 640 x 640 (measured crossover), one stream above;
    YB_LANES overrides either way, and with lanes every head level is emitted behind the FPN tensor it reads."""
    monkeypatch.delenv("YB_LANES", raising=False)
    arch = nn.yolo_v11_n(80)._arch
    lanes = lambda b, hw: Engine(*arch, b, hw, hw, host_only=True).describe()["num_lanes"]
    assert lanes(1, 640) == 4 and lanes(16, 640) == 4 and lanes(64, 320) == 4
    assert lanes(32, 640) == 1 and lanes(256, 640) == 1 and lanes(8, 1280) == 1
    names = [o["name"] for o in Engine(*arch, 1, 640, 640, host_only=True).describe()["ops"]]
    assert names.index("head.box.0.0") < names.index("fpn.h3") < names.index("head.box.1.0") < names.index("fpn.h5")
    names = [o["name"] for o in Engine(*arch, 256, 640, 640, host_only=True).describe()["ops"]]
    assert names.index("fpn.h6.conv2") < names.index("head.box.0.0")
    monkeypatch.setenv("YB_LANES", "1")
    assert lanes(1, 640) == 1
    monkeypatch.setenv("YB_LANES", "3")
    assert lanes(256, 640) == 3


def test_space_to_depth_sources_are_planned_only_where_they_are_safe(monkeypatch):
    """Stride-2 3x3 convs read a space-to-depth copy of their source through the halo-patch path when that source
    has no other reader and its producer can store it that way (stem: 16 channels; 1x1 conv on 8 x 16 tiles: 64)."""
    def plan(size, hw, batch=2):
        m = getattr(nn, f"yolo_v11_{size}")(80)
        return Engine(*m._arch, batch, hw, hw, host_only=True).describe()
    d = plan("n", 640)
    ops = {o["name"]: o for o in d["ops"]}
    assert ops["net.p2.0"]["s2d"] == 1 and ops["net.p3.0"]["s2d"] == 2
    assert d["bufs"][ops["net.p2.0"]["src"][0]["buf"]]["s2d"] == 1 and d["bufs"][ops["net.p3.0"]["src"][0]["buf"]]["s2d"] == 1
    # P3 / P4 / N3 / N4 have other readers: their stride-2 consumers keep the gather
    assert all(ops[n]["s2d"] == 0 for n in ("net.p4.0", "net.p5.0", "fpn.h3", "fpn.h5"))
    assert sum(b["s2d"] for b in d["bufs"]) == 2
    for name in ("net.p2.0", "net.p3.0"):   # a space-to-depth buffer has exactly one writer and one reader
        buf = ops[name]["src"][0]["buf"]
        users = [o["name"] for o in d["ops"] if o["dst"]["buf"] == buf or any(s["buf"] == buf for s in o["src"])
                 or (o["has_res"] and o["res"]["buf"] == buf)]
        assert len(users) == 2 and users[1] == name
    # 96 x 96 input: the 24 x 24 map of net.p2 is no multiple of the 8 x 16 store tile, and everything is < 40 x 40
    assert sum(b["s2d"] for b in plan("n", 96)["bufs"]) == 0
    # other widths (s: 32 / 128 channels, x: 96 / 384) are not covered
    assert sum(b["s2d"] for b in plan("s", 640)["bufs"]) == 0 and sum(b["s2d"] for b in plan("x", 320)["bufs"]) == 0
    monkeypatch.setenv("YB_NO_S2D", "1")
    assert sum(b["s2d"] for b in plan("n", 640)["bufs"]) == 0


def test_engine_export_round_trip(tmp_path):
    """export_engine -> load_engine (host-only here): the artefact alone reproduces the packed plan, and its CPU
    replay equals the replay of the engine packed from the model."""
    from yolo_infer_pt_b200 import export
    model = nn.yolo_v11_n(80)
    synth.load_synth(model, 0, "survey_widehead")
    path = export.export_engine(model, str(tmp_path / "yolo11n_2x64.npz"), 2, 64, 64)
    eng = export.load_engine(path, host_only=True)
    ref = Engine(*model._arch, 2, 64, 64, host_only=True)
    blob = ref.pack_from_model(model)
    assert np.array_equal(eng.host_blob, blob) and eng.describe() == ref.describe()
    x = synth.synth_images(2, 64, 64, seed=1)
    with torch.no_grad():
        a = PlanReplay(eng.describe(), eng.convs, eng.host_blob).run(x)
        b = PlanReplay(ref.describe(), ref.convs, blob).run(x)
    assert torch.equal(a, b)
    # an artefact from another plan layout is rejected
    g = dict(np.load(path))
    g["weight_bytes"] = np.int64(int(g["weight_bytes"]) + 256)
    np.savez(str(tmp_path / "bad.npz"), **g)
    with pytest.raises(RuntimeError):
        export.load_engine(str(tmp_path / "bad.npz"), host_only=True)
