"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.

CPU restatement, in plain fp32 PyTorch functional ops, of the reference's fused/unfused
`YOLO.forward` (t0saki/YOLO-Infer-pt nets/nn.py:28-297).  Only tests/, __graft_entry__.smoke()
and bench.py's CPU-baseline leg may import this file; nothing under yolo_infer_pt_b200/ does.

It is written as one function over a `state_dict` and the (width, depth, csp, nc) lists rather than
as nn.Module classes, so it shares no code with either the reference or the product.  Every conv
output is recorded under the op name the CUDA plan uses (e.g. "net.p2.1.res_m.0.conv2"), which is
what the layer-level parity tests compare.

Pinning: tests/golden/make_golden.py runs the *reference itself* (imported from /root/reference in
the build container) on seeded synthetic weights/inputs and commits its outputs under tests/golden/;
tests/test_oracle.py checks this restatement against those vectors.
"""
import torch
import torch.nn.functional as F

BN_EPS = 1e-3  # nets/nn.py:32


def fold_bn(weight, conv_bias, gamma, beta, mean, var, eps=BN_EPS):
    """fuse_conv, nets/nn.py:8-25, per output channel."""
    scale = gamma / torch.sqrt(var + eps)
    w = weight * scale.view(-1, 1, 1, 1)
    b0 = conv_bias if conv_bias is not None else torch.zeros_like(mean)
    b = scale * b0 + (beta - gamma * mean / torch.sqrt(var + eps))
    return w, b


class _Net:
    def __init__(self, sd, taps):
        self.sd = sd
        self.taps = taps

    def unit(self, name, x, k=1, s=1, act=True, groups=1):
        """`Conv` wrapper nets/nn.py:28-39 (conv, eval BatchNorm or folded bias, SiLU/Identity)."""
        sd = self.sd
        w = sd[name + ".conv.weight"].float()
        if name + ".norm.weight" in sd:
            y = F.conv2d(x, w, None, s, k // 2, 1, groups)
            y = F.batch_norm(y, sd[name + ".norm.running_mean"].float(), sd[name + ".norm.running_var"].float(),
                             sd[name + ".norm.weight"].float(), sd[name + ".norm.bias"].float(), False, 0.0, BN_EPS)
        else:
            y = F.conv2d(x, w, sd[name + ".conv.bias"].float(), s, k // 2, 1, groups)
        y = F.silu(y) if act else y
        return y

    def rec(self, name, y):
        if self.taps is not None:
            self.taps[name] = y
        return y

    def bottleneck(self, name, x, e):  # Residual nn.py:42-49
        t = self.rec(name + ".conv1", self.unit(name + ".conv1", x, 3))
        return self.rec(name + ".conv2", x + self.unit(name + ".conv2", t, 3))

    def c3k(self, name, x):  # CSPModule nn.py:52-63
        a = self.rec(name + ".conv1", self.unit(name + ".conv1", x))
        b = self.rec(name + ".conv2", self.unit(name + ".conv2", x))
        a = self.bottleneck(name + ".res_m.0", a, 1.0)
        a = self.bottleneck(name + ".res_m.1", a, 1.0)
        return self.rec(name + ".conv3", self.unit(name + ".conv3", torch.cat((a, b), 1)))

    def c3k2(self, name, x, n, use_c3k):  # CSP nn.py:66-80
        y = self.rec(name + ".conv1", self.unit(name + ".conv1", x))
        c = y.shape[1] // 2
        parts = [y[:, :c], y[:, c:]]
        for i in range(n):
            sub = f"{name}.res_m.{i}"
            parts.append(self.c3k(sub, parts[-1]) if use_c3k else self.bottleneck(sub, parts[-1], 0.5))
        return self.rec(name + ".conv2", self.unit(name + ".conv2", torch.cat(parts, 1)))

    def sppf(self, name, x):  # SPP nn.py:83-94
        y = self.rec(name + ".conv1", self.unit(name + ".conv1", x))
        pools = [y]
        for _ in range(3):
            pools.append(F.max_pool2d(pools[-1], 5, 1, 2))
        self.rec(name + ".res_m", torch.cat(pools[1:], 1))
        return self.rec(name + ".conv2", self.unit(name + ".conv2", torch.cat(pools, 1)))

    def attention(self, name, x, heads):  # Attention nn.py:97-123
        b, c, h, w = x.shape
        dh = c // heads
        dk = dh // 2
        qkv = self.rec(name + ".qkv", self.unit(name + ".qkv", x, act=False))
        q, k, v = qkv.view(b, heads, 2 * dk + dh, h * w).split([dk, dk, dh], 2)
        att = ((q.transpose(-2, -1) @ k) * dk ** -0.5).softmax(-1)
        o = (v @ att.transpose(-2, -1)).view(b, c, h, w)
        self.rec(name + ".attn", o)
        o = o + self.unit(name + ".conv1", v.reshape(b, c, h, w), 3, act=False, groups=c)
        self.rec(name + ".conv1", o)
        return self.unit(name + ".conv2", o, act=False)

    def psa_block(self, name, x, heads):  # PSABlock nn.py:126-136
        x = self.rec(name + ".conv1.conv2", x + self.attention(name + ".conv1", x, heads))
        f = self.rec(name + ".conv2.0", self.unit(name + ".conv2.0", x))
        return self.rec(name + ".conv2.1", x + self.unit(name + ".conv2.1", f, act=False))

    def c2psa(self, name, x, n):  # PSA nn.py:139-148
        y = self.rec(name + ".conv1", self.unit(name + ".conv1", x))
        c = y.shape[1] // 2
        keep, z = y[:, :c], y[:, c:]
        for i in range(n):
            z = self.psa_block(f"{name}.res_m.{i}", z, (2 * c) // 128)
        return self.rec(name + ".conv2", self.unit(name + ".conv2", torch.cat((keep, z), 1)))


def forward_raw(sd, width, depth, csp, nc, x, taps=None):
    """Returns the three pre-decode head maps [(B, 64+nc, H_i, W_i)] — the reference's training-mode
    output (nets/nn.py:256-259) computed with eval-mode BatchNorm."""
    net = _Net(sd, taps)
    d = depth
    x = x.float()
    # DarkNet nn.py:151-189
    p1 = net.rec("net.p1.0", net.unit("net.p1.0", x, 3, 2))
    t = net.rec("net.p2.0", net.unit("net.p2.0", p1, 3, 2))
    p2 = net.c3k2("net.p2.1", t, d[0], csp[0])
    t = net.rec("net.p3.0", net.unit("net.p3.0", p2, 3, 2))
    p3 = net.c3k2("net.p3.1", t, d[1], csp[0])
    t = net.rec("net.p4.0", net.unit("net.p4.0", p3, 3, 2))
    p4 = net.c3k2("net.p4.1", t, d[2], csp[1])
    t = net.rec("net.p5.0", net.unit("net.p5.0", p4, 3, 2))
    t = net.c3k2("net.p5.1", t, d[3], csp[1])
    t = net.sppf("net.p5.2", t)
    p5 = net.c2psa("net.p5.3", t, d[4])
    # DarkFPN nn.py:192-209
    up = lambda z: F.interpolate(z, scale_factor=2.0, mode="nearest")  # noqa: E731
    t4 = net.c3k2("fpn.h1", torch.cat((up(p5), p4), 1), d[5], csp[0])
    n3 = net.c3k2("fpn.h2", torch.cat((up(t4), p3), 1), d[5], csp[0])
    h3 = net.rec("fpn.h3", net.unit("fpn.h3", n3, 3, 2))
    n4 = net.c3k2("fpn.h4", torch.cat((h3, t4), 1), d[5], csp[0])
    h5 = net.rec("fpn.h5", net.unit("fpn.h5", n4, 3, 2))
    n5 = net.c3k2("fpn.h6", torch.cat((h5, p5), 1), d[5], csp[1])
    # Head nn.py:244-257
    outs = []
    for i, f in enumerate((n3, n4, n5)):
        bn, cn = f"head.box.{i}", f"head.cls.{i}"
        b = net.rec(bn + ".0", net.unit(bn + ".0", f, 3))
        b = net.rec(bn + ".1", net.unit(bn + ".1", b, 3))
        b = F.conv2d(b, sd[bn + ".2.weight"].float(), sd[bn + ".2.bias"].float())
        c = net.rec(cn + ".0", net.unit(cn + ".0", f, 3, groups=f.shape[1]))
        c = net.rec(cn + ".1", net.unit(cn + ".1", c))
        c = net.rec(cn + ".2", net.unit(cn + ".2", c, 3, groups=c.shape[1]))
        c = net.rec(cn + ".3", net.unit(cn + ".3", c))
        c = F.conv2d(c, sd[cn + ".4.weight"].float(), sd[cn + ".4.bias"].float())
        outs.append(torch.cat((b, c), 1))
    return outs


def decode(maps, nc, strides=(8.0, 16.0, 32.0)):
    """Head eval path nn.py:261-270 with make_anchors (utils/util.py:85-96) and DFL (nn.py:222-225)."""
    b = maps[0].shape[0]
    pts, scl = [], []
    for m, s in zip(maps, strides):
        h, w = m.shape[-2:]
        ys = torch.arange(h, dtype=torch.float32) + 0.5
        xs = torch.arange(w, dtype=torch.float32) + 0.5
        gy, gx = torch.meshgrid(ys, xs, indexing="ij")
        pts.append(torch.stack((gx, gy), -1).view(-1, 2))
        scl.append(torch.full((h * w, 1), s, dtype=torch.float32))
    anchors = torch.cat(pts).t().unsqueeze(0)   # (1, 2, A)
    strides_t = torch.cat(scl).t()             # (1, A)
    x = torch.cat([m.reshape(b, 64 + nc, -1) for m in maps], 2)
    box, cls = x[:, :64], x[:, 64:]
    a = box.shape[-1]
    prob = box.view(b, 4, 16, a).transpose(2, 1).softmax(1)            # (b, 16, 4, a)
    dist = (prob * torch.arange(16, dtype=torch.float32).view(1, 16, 1, 1)).sum(1)  # expectation
    lt, rb = dist[:, :2], dist[:, 2:]
    x1y1 = anchors - lt
    x2y2 = anchors + rb
    xywh = torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), 1)
    return torch.cat((xywh * strides_t, cls.sigmoid()), 1)


def forward(sd, width, depth, csp, nc, x, taps=None):
    """Eval-mode YOLO.forward (nets/nn.py:294-297): (B,3,H,W) -> (B, 4+nc, A) fp32."""
    return decode(forward_raw(sd, width, depth, csp, nc, x, taps), nc)


def raw_to_rows(maps):
    """[(B, no, H, W)] -> (B, A, no): the layout yb_forward_raw returns."""
    b, no = maps[0].shape[:2]
    return torch.cat([m.reshape(b, no, -1) for m in maps], 2).transpose(1, 2).contiguous()
