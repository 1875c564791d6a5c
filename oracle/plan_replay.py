"""ORACLE — TEST INFRASTRUCTURE ONLY.  Replays a libyolob200 execution plan on the CPU.

Takes the plan description (`Engine.describe()`: buffers, ops, channel slices, GEMM K layouts) and
the packed weight blob (`Engine.pack_from_model`) and executes every op with fp32 torch ops over
NHWC buffers, exactly as the CUDA kernels are specified to: same slices, same K ordering of the
packed weights, same padded channels, optional rounding to the plan's 16-bit storage type (fp16 or
bf16, `desc["act_f16"]`) at every activation store.  Buffers
start as NaN, so an op that reads a channel nobody wrote poisons the output.

This validates the *host logic* (plan builder, buffer aliasing, weight packer) without a GPU, and
gives the GPU tests a bf16-faithful expectation to separate precision from bugs.
"""
import numpy as np
import torch
import torch.nn.functional as F


def _bf16(t):
    return t.to(torch.bfloat16).to(torch.float32)


def _fp16(t):
    return t.to(torch.float16).to(torch.float32)


def _cpad8(c):
    return (c + 7) // 8 * 8


class PlanReplay:
    def __init__(self, desc, convs, blob, emulate_bf16=True):
        """emulate_bf16 (historic name): round every 16-bit activation store to the plan's storage type."""
        self.d = desc
        self.convs = convs
        self.blob = np.asarray(blob, dtype=np.uint8)
        self.bf16 = emulate_bf16
        self.f16 = bool(desc.get("act_f16", 0))
        self.rnd = _fp16 if self.f16 else _bf16
        self.tdtype = torch.float16 if self.f16 else torch.bfloat16
        self.B = desc["B"]
        self.bufs = {}

    def _buf(self, i):
        if i not in self.bufs:
            b = self.d["bufs"][i]
            self.bufs[i] = torch.full((self.B, b["rows_per_img"], b["C"]), float("nan"))
        return self.bufs[i]

    def _store(self, sl, rows, row_off=0):
        """rows: (B, n_rows, C_store) -> dst slice (bf16-rounded unless the buffer is fp32)."""
        buf = self._buf(sl["buf"])
        if self.bf16 and self.d["bufs"][sl["buf"]]["elem_bytes"] == 2:
            rows = self.rnd(rows)
        buf[:, row_off:row_off + rows.shape[1], sl["c_off"]:sl["c_off"] + rows.shape[2]] = rows

    def _load_nhwc(self, sl, channels=None):
        b = self.d["bufs"][sl["buf"]]
        c = sl["C"] if channels is None else channels
        t = self._buf(sl["buf"])[:, :, sl["c_off"]:sl["c_off"] + c].reshape(self.B, b["H"], b["W"], c)
        if sl["up"]:
            t = t.repeat_interleave(2, 1).repeat_interleave(2, 2)
        return t

    # ---- weights --------------------------------------------------------------------------
    def _dense_weights(self, op):
        cw = self.convs[op["conv_index"]]
        raw = self.blob[cw["blob_offset"]:cw["blob_offset"] + cw["blob_bytes"]]
        nw = op["N_pad"] * op["K_pad"]
        if self.f16:
            W = torch.from_numpy(raw[:nw * 2].view(np.float16).astype(np.float32).reshape(op["N_pad"], op["K_pad"]))
        else:
            w16 = raw[:nw * 2].view(np.uint16).astype(np.uint32) << 16
            W = torch.from_numpy(w16.view(np.float32).reshape(op["N_pad"], op["K_pad"]).copy())
        bias = torch.from_numpy(raw[nw * 2:nw * 2 + op["N_pad"] * 4].view(np.float32).copy())
        return W, bias

    def _f32_weights(self, op, taps):
        cw = self.convs[op["conv_index"]]
        cp = _cpad8(cw["cout"])
        raw = self.blob[cw["blob_offset"]:cw["blob_offset"] + (taps + 1) * cp * 4].view(np.float32)
        W = torch.from_numpy(raw[:taps * cp].reshape(taps, cp).copy())
        bias = torch.from_numpy(raw[taps * cp:(taps + 1) * cp].copy())
        return W, bias, cp

    # ---- ops ------------------------------------------------------------------------------
    def _conv(self, op):
        W, bias = self._dense_weights(op)
        k, s, pad = op["k"], op["stride"], op["k"] // 2
        cols = []
        for tap in range(k * k):
            dy, dx = (tap // 3, tap % 3) if k == 3 else (0, 0)
            for si, sl in enumerate(op["src"]):
                cp = _cpad8(sl["C"])
                x = self._load_nhwc(sl, cp)                      # (B, Hin, Win, cp)
                x = F.pad(x, (0, 0, pad, pad, pad, pad))         # zero padding like the kernel's OOB fill
                x = x[:, dy:dy + s * op["Hout"]:s, dx:dx + s * op["Wout"]:s, :][:, :op["Hout"], :op["Wout"]]
                x = x.reshape(self.B, op["Hout"] * op["Wout"], cp)
                # per-source K padding: 64-channel blocks for TMA-fed 1x1 layers and for tap-aligned 3x3
                # layers (seg_kpad == cp everywhere else)
                x = F.pad(x, (0, op["seg_kpad"][si] - cp))
                cols.append(x)
        A = torch.cat(cols, 2)
        A = F.pad(A, (0, op["K_pad"] - A.shape[2]))
        if self.bf16:
            A = self.rnd(A)  # activations are stored in 16 bits already; NaN stays NaN
        D = A @ W.t() + bias
        if op["act"]:
            D = F.silu(D)
        cout = op["dst"]["C"]
        if op["out_f32"]:
            store = (cout + 3) // 4 * 4
        else:
            store = _cpad8(cout)
            if op["has_res"]:
                r = self._buf(op["res"]["buf"])[:, :, op["res"]["c_off"]:op["res"]["c_off"] + store]
                D = D.clone()
                D[:, :, :store] += r
        self._store(op["dst"], D[:, :, :store], op["dst_row_off"])

    def _stem(self, op, x):
        W, bias, cp = self._f32_weights(op, 27)
        w = W.t().reshape(cp, 3, 3, 3)
        y = F.silu(F.conv2d(x.float(), w, bias, 2, 1))
        self._store(op["dst"], y.permute(0, 2, 3, 1).reshape(self.B, -1, cp))

    def _dw(self, op):
        W, bias, cp = self._f32_weights(op, 9)
        C = op["dst"]["C"]
        gsz, gstride, goff, add = op["dw"]
        sl = op["src"][0]
        b = self.d["bufs"][sl["buf"]]
        full = self._buf(sl["buf"]).reshape(self.B, b["H"], b["W"], b["C"])
        idx = torch.tensor([sl["c_off"] + (c // gsz) * gstride + goff + c % gsz for c in range(C)])
        x = full[..., idx].permute(0, 3, 1, 2)
        w = W[:, :C].t().reshape(C, 1, 3, 3)
        y = F.conv2d(x, w, bias[:C], 1, 1, 1, C)
        if op["act"]:
            y = F.silu(y)
        y = y.permute(0, 2, 3, 1).reshape(self.B, -1, C)
        if add:
            y = y + self._buf(op["dst"]["buf"])[:, :, op["dst"]["c_off"]:op["dst"]["c_off"] + C]
        self._store(op["dst"], y)

    def _pool(self, op):
        x = self._load_nhwc(op["src"][0]).permute(0, 3, 1, 2)
        outs = []
        for _ in range(3):
            x = F.max_pool2d(x, 5, 1, 2)
            outs.append(x)
        y = torch.cat(outs, 1).permute(0, 2, 3, 1).reshape(self.B, -1, 3 * op["src"][0]["C"])
        self._store(op["dst"], y)

    def _attn(self, op):
        sl = op["src"][0]
        heads = op["heads"]
        qkv = self._buf(sl["buf"])[:, :, sl["c_off"]:sl["c_off"] + sl["C"]]
        n = qkv.shape[1]
        qkv = qkv.reshape(self.B, n, heads, 128)
        q, k, v = qkv[..., :32], qkv[..., 32:64], qkv[..., 64:]
        att = torch.einsum("bihd,bjhd->bhij", q, k) * op["scale"]
        att = att.softmax(-1)
        o = torch.einsum("bhij,bjhd->bihd", att, v).reshape(self.B, n, heads * 64)
        self._store(op["dst"], o)

    def _decode(self):
        d = self.d
        nc = d["nc"]
        lg = self._buf(d["logits_buf"])[:, :, :64 + nc]           # (B, A, no)
        box, cls = lg[:, :, :64], lg[:, :, 64:]
        dist = (box.reshape(self.B, -1, 4, 16).softmax(-1) * torch.arange(16.0)).sum(-1)  # (B, A, 4)
        H, W = d["H"], d["W"]
        pts, scl = [], []
        for lvl in range(3):
            h, w, s = H >> (3 + lvl), W >> (3 + lvl), float(8 << lvl)
            gy, gx = torch.meshgrid(torch.arange(h) + 0.5, torch.arange(w) + 0.5, indexing="ij")
            pts.append(torch.stack((gx, gy), -1).reshape(-1, 2))
            scl.append(torch.full((h * w, 1), s))
        anc, st = torch.cat(pts), torch.cat(scl)
        x1y1 = anc - dist[:, :, :2]
        x2y2 = anc + dist[:, :, 2:]
        out = torch.cat(((x1y1 + x2y2) / 2 * st, (x2y2 - x1y1) * st, cls.sigmoid()), 2)
        return out.transpose(1, 2).contiguous()

    def step(self, op, x=None):
        """Execute one op on the current buffer state; returns the decode output for the last op."""
        kind = op["kind"]
        if kind == 0:
            self._stem(op, x)
        elif kind == 1:
            self._conv(op)
        elif kind == 2:
            self._dw(op)
        elif kind == 3:
            self._pool(op)
        elif kind == 4:
            self._attn(op)
        elif kind == 5:
            return self._decode()
        return None

    def buffer_bytes(self, i):
        """Buffer i in the GPU's storage format (fp16 / bf16 or fp32 NHWC); unwritten (NaN) entries as 0."""
        t = torch.nan_to_num(self._buf(i), nan=0.0)
        return t.to(self.tdtype) if self.d["bufs"][i]["elem_bytes"] == 2 else t.float()

    def run(self, x, taps=None):
        """x: (B,3,H,W) fp32 -> (B, 4+nc, A).  taps: optional dict filled with each op's dst slice
        (B, rows, C) after the op ran."""
        out = None
        for op in self.d["ops"]:
            kind = op["kind"]
            if kind == 0:
                self._stem(op, x)
            elif kind == 1:
                self._conv(op)
            elif kind == 2:
                self._dw(op)
            elif kind == 3:
                self._pool(op)
            elif kind == 4:
                self._attn(op)
            elif kind == 5:
                out = self._decode()
            if taps is not None and kind != 5:
                sl = op["dst"]
                b = self.d["bufs"][sl["buf"]]
                r0 = op["dst_row_off"]
                nrows = op["Hout"] * op["Wout"]
                taps[op["name"]] = self._buf(sl["buf"])[:, r0:r0 + nrows, sl["c_off"]:sl["c_off"] + sl["C"]].clone()
                del b
        return out

    def final_slice(self, op):
        """Content of op's dst slice at the END of the forward (in-place updates by later ops
        included) — what yb_plan_debug_read sees on the GPU."""
        sl = op["dst"]
        r0 = op["dst_row_off"]
        nrows = op["Hout"] * op["Wout"]
        return self._buf(sl["buf"])[:, r0:r0 + nrows, sl["c_off"]:sl["c_off"] + sl["C"]].clone()

    def raw_logits(self):
        return self._buf(self.d["logits_buf"])[:, :, :64 + self.d["nc"]].clone()
