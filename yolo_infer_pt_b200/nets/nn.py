"""Drop-in for the reference's `nets.nn` (t0saki/YOLO-Infer-pt nets/nn.py) whose eval-mode forward
runs on libyolob200.so (hand-written sm_100a kernels) instead of torch/cuDNN ops.

What is kept identical to the reference, so that checkpoints, pickles and callers keep working:
  * module path `nets.nn`, class names and attribute names, hence every `state_dict` key
    (`net.p1.0.conv.weight`, `net.p1.0.norm.running_mean`, `head.box.0.2.bias`, `head.dfl.conv.weight`);
  * `yolo_v11_{n,t,s,m,l,x}(num_classes)` and `YOLO(width, depth, csp, num_classes)` (nn.py:282-347);
  * `model.fuse()` folding BatchNorm into the convs (nn.py:8-25, 299-305);
  * eval forward -> `(B, 4+nc, A)` rows [cx, cy, w, h, scores], anchors level 8/16/32 row-major
    (nn.py:262-270); training forward -> list of 3 raw maps `(B, 64+nc, H_i, W_i)` (nn.py:258-259).

What is different:
  * eval forward requires a CUDA (sm_100) input and the built extension; it raises otherwise —
    there is no CPU or eager-PyTorch fallback for inference;
  * the eval output is always fp32 (fp16 - or bf16 - activations, fp32 accumulation and decode);
  * the modules below describe *parameters and topology*; their torch `forward`s are only the
    differentiable training-mode graph (autograd for ComputeLoss, utils/util.py:863-930) and are
    never reached from eval mode.
"""
import math
import os
import sys

import torch


def _silu():
    return torch.nn.SiLU()


def _engine_class():
    """The Engine lives in the package root; `nets` may have been imported drop-in style
    (sys.path pointing inside yolo_infer_pt_b200/), so make the root importable on demand."""
    try:
        from yolo_infer_pt_b200.engine import Engine
    except ImportError:
        root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        sys.path.insert(0, root)
        from yolo_infer_pt_b200.engine import Engine
    return Engine


def fuse_conv(conv, norm):
    """BatchNorm folding, reference nn.py:8-25: W' = diag(g / sqrt(var + eps)) W, b' = g (b - mean) / sqrt(var + eps) + beta."""
    out = torch.nn.Conv2d(conv.in_channels, conv.out_channels, conv.kernel_size, conv.stride, conv.padding,
                          groups=conv.groups, bias=True).requires_grad_(False).to(conv.weight.device)
    with torch.no_grad():
        scale = norm.weight / torch.sqrt(norm.running_var + norm.eps)
        out.weight.copy_(conv.weight * scale.view(-1, 1, 1, 1))
        bias = conv.bias if conv.bias is not None else torch.zeros_like(norm.running_mean)
        out.bias.copy_(scale * (bias - norm.running_mean) + norm.bias)
    return out


class Conv(torch.nn.Module):
    """conv -> BatchNorm(eps 1e-3, momentum 0.03) -> activation (nn.py:28-39). After `fuse()` the
    `norm` attribute is gone and `conv` carries the folded weights plus a bias."""

    def __init__(self, in_ch, out_ch, activation, k=1, s=1, p=0, g=1):
        super().__init__()
        self.conv = torch.nn.Conv2d(in_ch, out_ch, k, s, p, groups=g, bias=False)
        self.norm = torch.nn.BatchNorm2d(out_ch, eps=0.001, momentum=0.03)
        self.relu = activation

    def forward(self, x):
        y = self.conv(x)
        norm = getattr(self, "norm", None)
        return self.relu(y if norm is None else norm(y))

    fuse_forward = forward  # the reference rebinds forward to this name after fusing (nn.py:303)


class Residual(torch.nn.Module):
    def __init__(self, ch, e=0.5):
        super().__init__()
        hidden = int(ch * e)
        self.conv1 = Conv(ch, hidden, _silu(), k=3, p=1)
        self.conv2 = Conv(hidden, ch, _silu(), k=3, p=1)

    def forward(self, x):
        return x + self.conv2(self.conv1(x))


class CSPModule(torch.nn.Module):
    def __init__(self, in_ch, out_ch):
        super().__init__()
        half = out_ch // 2
        self.conv1 = Conv(in_ch, half, _silu())
        self.conv2 = Conv(in_ch, half, _silu())
        self.conv3 = Conv(2 * half, out_ch, _silu())
        self.res_m = torch.nn.Sequential(Residual(half, e=1.0), Residual(half, e=1.0))

    def forward(self, x):
        return self.conv3(torch.cat((self.res_m(self.conv1(x)), self.conv2(x)), 1))


class CSP(torch.nn.Module):
    def __init__(self, in_ch, out_ch, n, csp, r):
        super().__init__()
        c = out_ch // r
        self.conv1 = Conv(in_ch, 2 * c, _silu())
        self.conv2 = Conv((2 + n) * c, out_ch, _silu())
        make = (lambda: CSPModule(c, c)) if csp else (lambda: Residual(c))
        self.res_m = torch.nn.ModuleList(make() for _ in range(n))

    def forward(self, x):
        parts = list(self.conv1(x).chunk(2, 1))
        for m in self.res_m:
            parts.append(m(parts[-1]))
        return self.conv2(torch.cat(parts, 1))


class SPP(torch.nn.Module):
    def __init__(self, in_ch, out_ch, k=5):
        super().__init__()
        self.conv1 = Conv(in_ch, in_ch // 2, _silu())
        self.conv2 = Conv(in_ch * 2, out_ch, _silu())
        self.res_m = torch.nn.MaxPool2d(k, stride=1, padding=k // 2)

    def forward(self, x):
        parts = [self.conv1(x)]
        for _ in range(3):
            parts.append(self.res_m(parts[-1]))
        return self.conv2(torch.cat(parts, 1))


class Attention(torch.nn.Module):
    def __init__(self, ch, num_head):
        super().__init__()
        self.num_head = num_head
        self.dim_head = ch // num_head
        self.dim_key = self.dim_head // 2
        self.scale = self.dim_key ** -0.5
        self.qkv = Conv(ch, ch + 2 * self.dim_key * num_head, torch.nn.Identity())
        self.conv1 = Conv(ch, ch, torch.nn.Identity(), k=3, p=1, g=ch)
        self.conv2 = Conv(ch, ch, torch.nn.Identity())

    def forward(self, x):
        b, c, h, w = x.shape
        qkv = self.qkv(x).view(b, self.num_head, 2 * self.dim_key + self.dim_head, h * w)
        q, k, v = qkv.split([self.dim_key, self.dim_key, self.dim_head], 2)
        weights = ((q.transpose(-2, -1) @ k) * self.scale).softmax(-1)
        y = (v @ weights.transpose(-2, -1)).view(b, c, h, w) + self.conv1(v.reshape(b, c, h, w))
        return self.conv2(y)


class PSABlock(torch.nn.Module):
    def __init__(self, ch, num_head):
        super().__init__()
        self.conv1 = Attention(ch, num_head)
        self.conv2 = torch.nn.Sequential(Conv(ch, 2 * ch, _silu()), Conv(2 * ch, ch, torch.nn.Identity()))

    def forward(self, x):
        x = x + self.conv1(x)
        return x + self.conv2(x)


class PSA(torch.nn.Module):
    def __init__(self, ch, n):
        super().__init__()
        self.conv1 = Conv(ch, 2 * (ch // 2), _silu())
        self.conv2 = Conv(2 * (ch // 2), ch, _silu())
        self.res_m = torch.nn.Sequential(*(PSABlock(ch // 2, ch // 128) for _ in range(n)))

    def forward(self, x):
        keep, y = self.conv1(x).chunk(2, 1)
        return self.conv2(torch.cat((keep, self.res_m(y)), 1))


class DarkNet(torch.nn.Module):
    def __init__(self, width, depth, csp):
        super().__init__()
        w, d = width, depth
        down = lambda i, o: Conv(i, o, _silu(), k=3, s=2, p=1)  # noqa: E731
        self.p1 = torch.nn.Sequential(down(w[0], w[1]))
        self.p2 = torch.nn.Sequential(down(w[1], w[2]), CSP(w[2], w[3], d[0], csp[0], r=4))
        self.p3 = torch.nn.Sequential(down(w[3], w[3]), CSP(w[3], w[4], d[1], csp[0], r=4))
        self.p4 = torch.nn.Sequential(down(w[4], w[4]), CSP(w[4], w[4], d[2], csp[1], r=2))
        self.p5 = torch.nn.Sequential(down(w[4], w[5]), CSP(w[5], w[5], d[3], csp[1], r=2),
                                      SPP(w[5], w[5]), PSA(w[5], d[4]))

    def forward(self, x):
        p3 = self.p3(self.p2(self.p1(x)))
        p4 = self.p4(p3)
        return p3, p4, self.p5(p4)


class DarkFPN(torch.nn.Module):
    def __init__(self, width, depth, csp):
        super().__init__()
        w, n = width, depth[5]
        self.up = torch.nn.Upsample(scale_factor=2)
        self.h1 = CSP(w[4] + w[5], w[4], n, csp[0], r=2)
        self.h2 = CSP(w[4] + w[4], w[3], n, csp[0], r=2)
        self.h3 = Conv(w[3], w[3], _silu(), k=3, s=2, p=1)
        self.h4 = CSP(w[3] + w[4], w[4], n, csp[0], r=2)
        self.h5 = Conv(w[4], w[4], _silu(), k=3, s=2, p=1)
        self.h6 = CSP(w[4] + w[5], w[5], n, csp[1], r=2)

    def forward(self, x):
        p3, p4, p5 = x
        t4 = self.h1(torch.cat((self.up(p5), p4), 1))
        n3 = self.h2(torch.cat((self.up(t4), p3), 1))
        n4 = self.h4(torch.cat((self.h3(n3), t4), 1))
        n5 = self.h6(torch.cat((self.h5(n4), p5), 1))
        return n3, n4, n5


class DFL(torch.nn.Module):
    """Distribution focal loss projection: softmax over 16 bins, expectation with weights 0..15."""

    def __init__(self, ch=16):
        super().__init__()
        self.ch = ch
        self.conv = torch.nn.Conv2d(ch, 1, 1, bias=False).requires_grad_(False)
        self.conv.weight.data[:] = torch.arange(ch, dtype=torch.float).view(1, ch, 1, 1)

    def forward(self, x):
        b, _, a = x.shape
        return self.conv(x.view(b, 4, self.ch, a).transpose(2, 1).softmax(1)).view(b, 4, a)


class Head(torch.nn.Module):
    anchors = torch.empty(0)
    strides = torch.empty(0)

    def __init__(self, nc=80, filters=()):
        super().__init__()
        self.ch = 16
        self.nc = nc
        self.nl = len(filters)
        self.no = nc + 4 * self.ch
        self.stride = torch.zeros(self.nl)
        box = max(64, filters[0] // 4)
        cls = max(80, filters[0], self.nc)
        self.dfl = DFL(self.ch)
        self.box = torch.nn.ModuleList(
            torch.nn.Sequential(Conv(f, box, _silu(), k=3, p=1), Conv(box, box, _silu(), k=3, p=1),
                                torch.nn.Conv2d(box, 4 * self.ch, 1)) for f in filters)
        self.cls = torch.nn.ModuleList(
            torch.nn.Sequential(Conv(f, f, _silu(), k=3, p=1, g=f), Conv(f, cls, _silu()),
                                Conv(cls, cls, _silu(), k=3, p=1, g=cls), Conv(cls, cls, _silu()),
                                torch.nn.Conv2d(cls, self.nc, 1)) for f in filters)

    def forward(self, x):
        """Training-mode graph only (eval mode is served by the CUDA engine in YOLO.forward)."""
        for i in range(self.nl):
            x[i] = torch.cat((self.box[i](x[i]), self.cls[i](x[i])), 1)
        return x

    def initialize_biases(self):
        for box, cls, s in zip(self.box, self.cls, self.stride):
            box[-1].bias.data[:] = 1.0
            cls[-1].bias.data[:self.nc] = math.log(5 / self.nc / (640 / s) ** 2)


class YOLO(torch.nn.Module):
    def __init__(self, width, depth, csp, num_classes):
        super().__init__()
        self._arch = (tuple(width), tuple(depth), tuple(bool(c) for c in csp), int(num_classes))
        self.net = DarkNet(width, depth, csp)
        self.fpn = DarkFPN(width, depth, csp)
        self.head = Head(num_classes, (width[3], width[4], width[5]))
        # the reference probes these with a 256x256 dummy forward (nn.py:288-290); the three
        # pyramid levels are /8, /16, /32 by construction
        self.head.stride = torch.tensor([8.0, 16.0, 32.0])
        self.stride = self.head.stride
        self.head.initialize_biases()

    # ---- reference API ---------------------------------------------------------------------
    def forward(self, x):
        if self.training:
            return self.head(list(self.fpn(self.net(x))))
        eng = self._engine_for(x)
        self._set_head_anchors(eng, x.device)
        # a fresh tensor per call, like the reference (two live predictions never alias); the Engine /
        # StreamingDetector API keeps static buffers for callers that want them
        if eng.graph:
            return eng.forward_static(x)
        return eng.forward(x, out=torch.empty_like(eng.out))

    def fuse(self):
        for m in self.modules():
            if type(m) is Conv and hasattr(m, "norm"):
                m.conv = fuse_conv(m.conv, m.norm)
                delattr(m, "norm")
        self.invalidate_engine()
        return self

    # ---- B200 engine management ------------------------------------------------------------
    MAX_ENGINES = 4          # cached plans (+ multi-GB workspaces) per model; least recently used is dropped

    def forward_raw(self, x):
        """Pre-decode head logits (B, A, 64+nc) fp32 computed by the CUDA engine."""
        return self._engine_for(x).forward_raw(x)

    def invalidate_engine(self):
        """Drop every cached engine (packed weights, plans, workspaces).  Called automatically by fuse(),
        .to()/.half()/.float() on the model and load_state_dict(); call it by hand after writing weights
        through `.data` IN PLACE (`p.data.copy_(w)`, `bias.data[:] = v`): such writes leave no trace
        (`.data` has its own version counter) and the engine would keep serving the packed copy."""
        self.__dict__["_yb_engines"] = {}
        self.__dict__.pop("_yb_slots", None)

    def _weights_version(self):
        """Fingerprint of the tensors the engine packed: (data_ptr, _version) of every parameter and buffer,
        plus the identity of the tensor each module slot currently holds.  Catches in-place updates of the
        parameters themselves (optimizer steps, `load_state_dict`), re-assignment (`p.data = w`,
        `load_state_dict(assign=True)` on a submodule) and dtype/device moves of submodules (`.half()` swaps
        the storage).  Not caught: in-place writes through `.data` and replacing whole modules by hand -
        call invalidate_engine() after those."""
        slots = self.__dict__.get("_yb_slots")
        if slots is None:
            slots = []
            for m in self.modules():
                slots += [(m._parameters, k, m._parameters[k]) for k in m._parameters if m._parameters[k] is not None]
                slots += [(m._buffers, k, m._buffers[k]) for k in m._buffers if m._buffers[k] is not None]
            self.__dict__["_yb_slots"] = slots
        h = 0
        for d, k, t in slots:
            if d[k] is not t:           # the slot holds another tensor now: rebuild the list, force a re-pack
                self.__dict__.pop("_yb_slots", None)
                self.__dict__["_yb_epoch"] = self.__dict__.get("_yb_epoch", 0) + 1
                return self._weights_version()
            h += t.data_ptr() + t._version
        return h + (self.__dict__.get("_yb_epoch", 0) << 56)

    def _engine_for(self, x):
        if not (isinstance(x, torch.Tensor) and x.is_cuda):
            raise RuntimeError(
                "yolo_infer_pt_b200: eval-mode forward needs a CUDA tensor on an sm_100 device; "
                "there is no CPU / eager-PyTorch fallback on the inference path")
        if x.dim() != 4 or x.shape[1] != 3:
            raise RuntimeError(f"expected a (B,3,H,W) image tensor, got {tuple(x.shape)}")
        Engine = _engine_class()
        engines = self.__dict__.setdefault("_yb_engines", {})
        key = (x.device.index, x.shape[0], x.shape[2], x.shape[3], self.__dict__.get("_yb_act_dtype"))
        version = self._weights_version()
        entry = engines.get(key)
        if entry is None or entry[1] != version:
            width, depth, csp, nc = self._arch
            eng = entry[0] if entry is not None else Engine(width, depth, csp, nc, x.shape[0], x.shape[2],
                                                           x.shape[3], x.device,
                                                           act_dtype=self.__dict__.get("_yb_act_dtype"))
            eng.pack_from_model(self)
            if entry is None and x.shape[0] <= 8:
                eng.use_graph(True)  # small batches are launch-bound: replay one CUDA graph
            engines.pop(key, None)
            engines[key] = entry = (eng, version)
            while len(engines) > self.MAX_ENGINES:   # variable shapes must not grow memory without bound
                engines.pop(next(iter(engines)))
        elif next(reversed(engines)) != key:
            engines[key] = engines.pop(key)          # most recently used last
        return entry[0]

    def set_activation_dtype(self, dtype):
        """Storage type of the engine's activations and packed weights: torch.float16 (default; what the
        reference's own evaluation runs in, main.py:251,266) or torch.bfloat16.  Accumulation, bias, SiLU,
        DFL decode and sigmoid are fp32 either way."""
        if dtype not in (torch.float16, torch.bfloat16, None):
            raise ValueError("activation dtype must be torch.float16 or torch.bfloat16")
        self.__dict__["_yb_act_dtype"] = dtype
        return self

    def _set_head_anchors(self, eng, device):
        """`head.anchors` / `head.strides` as the reference's eval forward leaves them (nn.py:261; (2, A) and
        (1, A)); the CUDA decode computes them from the anchor index, so they are built once per shape here."""
        key = (eng.height, eng.width, device)
        cache = self.__dict__.setdefault("_yb_anchor_cache", {})
        if key not in cache:
            if len(cache) > 8:
                cache.clear()
            pts, scl = [], []
            for s in (8, 16, 32):
                h, w = eng.height // s, eng.width // s
                ys = torch.arange(h, device=device, dtype=torch.float32) + 0.5
                xs = torch.arange(w, device=device, dtype=torch.float32) + 0.5
                gy, gx = torch.meshgrid(ys, xs, indexing="ij")
                pts.append(torch.stack((gx, gy), -1).view(-1, 2))
                scl.append(torch.full((h * w, 1), float(s), device=device))
            cache[key] = (torch.cat(pts).transpose(0, 1), torch.cat(scl).transpose(0, 1))
        self.head.anchors, self.head.strides = cache[key]

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate_engine()
        return out

    def _apply(self, fn, *args, **kwargs):
        self.invalidate_engine()
        return super()._apply(fn, *args, **kwargs)

    def __getstate__(self):
        state = self.__dict__.copy()
        for k in ("_yb_engines", "_yb_slots", "_yb_anchor_cache"):
            state.pop(k, None)
        return state

    def __setstate__(self, state):
        """Also accepts a module pickled by the reference itself (`torch.load('weights/best.pt')['model']`,
        main.py:247-249: with `yolo_infer_pt_b200/` on sys.path the pickle's `nets.nn.*` classes resolve
        to this file): such a state has no `_arch`, so it is read back from the module tree."""
        self.__dict__.update(state)
        if "_arch" not in self.__dict__:
            self.__dict__["_arch"] = self._infer_arch()
        for k in ("_yb_engines", "_yb_slots", "_yb_anchor_cache"):
            self.__dict__.pop(k, None)

    def _infer_arch(self):
        """(width, depth, csp, num_classes) of the constructor call (nn.py:308-347) from the modules."""
        net, fpn = self.net, self.fpn
        width = (net.p1[0].conv.in_channels, net.p1[0].conv.out_channels, net.p2[0].conv.out_channels,
                 net.p3[0].conv.out_channels, net.p4[0].conv.out_channels, net.p5[0].conv.out_channels)
        depth = (len(net.p2[1].res_m), len(net.p3[1].res_m), len(net.p4[1].res_m), len(net.p5[1].res_m),
                 len(net.p5[3].res_m), len(fpn.h1.res_m))
        is_csp = lambda m: type(m).__name__ == "CSPModule"  # noqa: E731
        csp = (is_csp(net.p2[1].res_m[0]), is_csp(net.p4[1].res_m[0]))
        return width, depth, csp, int(self.head.nc)

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k in ("_yb_engines", "_yb_slots", "_yb_anchor_cache"):
                continue
            new.__dict__[k] = copy.deepcopy(v, memo)
        return new


def _variant(width, depth, csp):
    def build(num_classes: int = 80):
        return YOLO(width, depth, csp, num_classes)
    return build


yolo_v11_n = _variant([3, 16, 32, 64, 128, 256], [1] * 6, [False, True])
yolo_v11_t = _variant([3, 24, 48, 96, 192, 384], [1] * 6, [False, True])
yolo_v11_s = _variant([3, 32, 64, 128, 256, 512], [1] * 6, [False, True])
yolo_v11_m = _variant([3, 64, 128, 256, 512, 512], [1] * 6, [True, True])
yolo_v11_l = _variant([3, 64, 128, 256, 512, 512], [2] * 6, [True, True])
yolo_v11_x = _variant([3, 96, 192, 384, 768, 768], [2] * 6, [True, True])
