"""Host-side driver of one libyolob200 execution plan.

An Engine owns, for one (architecture, batch, height, width, device):
  * the C-side plan (op list, buffer aliasing, TMA descriptors),
  * the packed weight blob (BN folded as the reference's fuse_conv does, nets/nn.py:8-25),
  * the activation workspace arena and the static (B, 4+nc, A) fp32 output tensor.
PyTorch is used for device memory and streams only; every kernel on the path lives in the .so.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib

_DTYPES = {torch.float32: _lib.YB_F32, torch.float16: _lib.YB_F16, torch.bfloat16: _lib.YB_BF16,
           torch.uint8: _lib.YB_U8}


def fold_conv_bn(conv, norm=None):
    """Fold an eval-mode BatchNorm into the preceding conv: the arithmetic of fuse_conv
    (reference nets/nn.py:8-25) written per output channel instead of as a diag matmul."""
    w = conv.weight.detach().float().cpu()
    b = conv.bias.detach().float().cpu() if conv.bias is not None else torch.zeros(w.shape[0])
    if norm is None:
        return w.contiguous(), b.contiguous()
    scale = norm.weight.detach().float().cpu() / torch.sqrt(norm.eps + norm.running_var.detach().float().cpu())
    w = w * scale.view(-1, 1, 1, 1)
    b = scale * b + (norm.bias.detach().float().cpu()
                     - norm.weight.detach().float().cpu() * norm.running_mean.detach().float().cpu()
                     / torch.sqrt(norm.running_var.detach().float().cpu() + norm.eps))
    return w.contiguous(), b.contiguous()


def default_act_dtype():
    """Storage type of activations and packed weights: fp16 unless YB_ACT_DTYPE=bf16.
    fp16 is what the reference's own evaluation runs in (main.py:251,266 `.half()`); with fp32 accumulation
    and fp32 decode it keeps the head outputs within the north-star tolerance (0.5 px / 1e-2) on networks
    whose head actually spreads scores over (0, 1), where bf16 storage does not (tests/test_gpu_parity.py)."""
    return torch.bfloat16 if os.environ.get("YB_ACT_DTYPE", "fp16").lower() in ("bf16", "bfloat16") else torch.float16


class Engine:
    def __init__(self, width, depth, csp, num_classes, batch, height, width_px, device=None,
                 host_only=False, act_dtype=None):
        L = _lib.lib()
        act_dtype = default_act_dtype() if act_dtype is None else act_dtype
        if act_dtype not in (torch.float16, torch.bfloat16):
            raise ValueError("act_dtype must be torch.float16 or torch.bfloat16")
        self.act_dtype = act_dtype
        self.L = L
        arch = _lib.ArchDesc()
        arch.width[:] = list(width)
        arch.depth[:] = list(depth)
        arch.csp[:] = [int(bool(c)) for c in csp]
        arch.num_classes = int(num_classes)
        self.arch = arch
        self.batch, self.height, self.width = int(batch), int(height), int(width_px)
        self.host_only = host_only
        if host_only:
            dev_index = -1
            self.device = None
        else:
            self.device = torch.device(device if device is not None else "cuda")
            if self.device.type != "cuda":
                raise RuntimeError("yolo_infer_pt_b200 runs on CUDA (sm_100a) only; there is no CPU path")
            dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
            self.device = torch.device("cuda", dev_index)
        handle = ctypes.c_void_p()
        _lib.check(L.yb_plan_create(ctypes.byref(arch), self.batch, self.height, self.width, _DTYPES[act_dtype],
                                    dev_index, ctypes.byref(handle)), "yb_plan_create")
        self.plan = handle
        self.num_anchors = L.yb_plan_num_anchors(self.plan)
        self.num_outputs = L.yb_plan_num_outputs(self.plan)
        self.num_launches = L.yb_plan_num_launches(self.plan)
        self.weight_bytes = L.yb_plan_weight_bytes(self.plan)
        self.workspace_bytes = L.yb_plan_workspace_bytes(self.plan)
        self.convs = []
        info = _lib.ConvInfo()
        for i in range(L.yb_plan_num_convs(self.plan)):
            _lib.check(L.yb_plan_conv_info(self.plan, i, ctypes.byref(info)), "yb_plan_conv_info")
            self.convs.append(dict(index=i, name=info.name.decode(), cout=info.cout, cin=info.cin,
                                   ksize=info.ksize, stride=info.stride, groups=info.groups, act=info.act,
                                   wrapped=info.wrapped, kind=info.kind, blob_offset=info.blob_offset,
                                   blob_bytes=info.blob_bytes))
        self.host_blob = None
        self.dev_weights = None
        self.workspace = None
        self.out = None
        self.raw = None
        self.graph = False

    def __del__(self):
        try:
            if getattr(self, "plan", None):
                self.L.yb_plan_destroy(self.plan)
                self.plan = None
        except Exception:
            pass

    # ---- weights -------------------------------------------------------------------------
    def pack_from_model(self, model):
        """Fold + pack every conv of `model` (a nets.nn.YOLO, fused or not) into the kernel layout."""
        blob = np.zeros(self.weight_bytes, dtype=np.uint8)
        for c in self.convs:
            mod = model.get_submodule(c["name"])
            if c["wrapped"]:
                w, b = fold_conv_bn(mod.conv, getattr(mod, "norm", None))
            else:
                w, b = fold_conv_bn(mod, None)
            exp = (c["cout"], c["cin"], c["ksize"], c["ksize"])
            if tuple(w.shape) != exp:
                raise RuntimeError(f"conv {c['name']}: weight shape {tuple(w.shape)} != plan {exp}")
            wn = np.ascontiguousarray(w.numpy(), dtype=np.float32)
            bn = np.ascontiguousarray(b.numpy(), dtype=np.float32)
            _lib.check(self.L.yb_plan_pack_conv(self.plan, c["index"], wn.ctypes.data, bn.ctypes.data,
                                                blob.ctypes.data), "yb_plan_pack_conv")
        self.host_blob = blob
        if not self.host_only:
            self._bind()
        return blob

    def _bind(self):
        dev = self.device
        self.dev_weights = torch.from_numpy(self.host_blob).to(dev)
        if self.workspace is None:
            self.workspace = torch.zeros(self.workspace_bytes, dtype=torch.uint8, device=dev)
            self.out = torch.empty(self.batch, self.num_outputs, self.num_anchors, dtype=torch.float32,
                                   device=dev)
        torch.cuda.synchronize(dev)
        _lib.check(self.L.yb_plan_bind(self.plan, self.dev_weights.data_ptr(), self.workspace.data_ptr()),
                   "yb_plan_bind")

    # ---- execution -----------------------------------------------------------------------
    def use_graph(self, enable=True):
        self.graph = bool(enable)
        _lib.check(self.L.yb_plan_use_graph(self.plan, int(self.graph)), "yb_plan_use_graph")

    def profile(self, enable=True):
        _lib.check(self.L.yb_plan_profile(self.plan, int(enable)), "yb_plan_profile")

    def profile_read(self):
        """(mean ms per op over the forwards recorded since the last read, number of forwards)."""
        n_ops = len(self.describe()["ops"])
        ms = np.zeros(n_ops, dtype=np.float32)   # one slot per op of the plan
        n = _lib.check(self.L.yb_plan_profile_read(self.plan, ms.ctypes.data, len(ms)), "yb_plan_profile_read")
        return ms, n

    def set_conv_impl(self, impl):
        _lib.check(self.L.yb_plan_set_conv_impl(self.plan, int(impl)), "yb_plan_set_conv_impl")

    def _check_input(self, x):
        if self.dev_weights is None:
            raise RuntimeError("Engine has no weights bound; call pack_from_model first")
        if not x.is_cuda or x.device != self.device:
            raise RuntimeError(f"input must live on {self.device}; the inference path has no CPU fallback")
        if tuple(x.shape) != (self.batch, 3, self.height, self.width):
            raise RuntimeError(f"input shape {tuple(x.shape)} != plan {(self.batch, 3, self.height, self.width)}")
        if x.dtype not in _DTYPES:
            raise RuntimeError(f"unsupported input dtype {x.dtype}")
        return x if x.is_contiguous() else x.contiguous()

    def forward(self, x, out=None, nms_sink=None):
        """(B,3,H,W) image tensor -> the engine's static (B, 4+nc, A) fp32 prediction tensor
        (layout of reference nets/nn.py:262-270). The returned tensor is overwritten by the next call;
        pass `out` (same shape, fp32, contiguous) to write somewhere else - pipelines that overlap the
        NMS of one batch with the forward of the next alternate between two such tensors.
        `nms_sink=(workspace, conf, max_nms)`: the class-score epilogues also append every score > conf to
        the NMS candidate lists of `workspace` (util.nms_workspace); follow with
        util.nms_padded(..., workspace=workspace, prefiltered=True)."""
        x = self._check_input(x)
        if out is None:
            out = self.out
        elif out.shape != self.out.shape or out.dtype != torch.float32 or not out.is_contiguous() or out.device != self.out.device:
            raise ValueError("Engine.forward: `out` must be a contiguous fp32 tensor shaped like the prediction tensor")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        if nms_sink is not None:
            ws, conf, max_nms = nms_sink
            conf32 = float(np.float32(conf))   # the NMS compares fp32 scores with an fp32 threshold
            if getattr(ws, "_yb_sink_pending", False):
                # a previous forward filled this workspace and no NMS consumed it: its counters would be
                # double-counted, so clear the headers first (stream-ordered memset)
                _lib.check(self.L.yb_nms_workspace_init(ws.data_ptr(), ws.numel(), ctypes.c_void_p(stream)),
                           "yb_nms_workspace_init")
            _lib.check(self.L.yb_forward_nms(self.plan, x.data_ptr(), _DTYPES[x.dtype], out.data_ptr(), conf32,
                                             int(max_nms), ws.data_ptr(), ws.numel(), ctypes.c_void_p(stream)),
                       "yb_forward_nms")
            ws._yb_sink_tag = (conf32, int(max_nms), out.data_ptr())   # checked by util.nms_padded(prefiltered=True)
            ws._yb_sink_pending = True
            return out
        _lib.check(self.L.yb_forward(self.plan, x.data_ptr(), _DTYPES[x.dtype], out.data_ptr(),
                                     ctypes.c_void_p(stream)), "yb_forward")
        return out

    def forward_static(self, x):
        """Graph-mode forward for callers that hand in a fresh tensor every call (the drop-in `model(x)`): the
        CUDA graph is keyed on its input / output pointers, so `x` is copied into a static input buffer, the
        graph replays into the static output and a fresh copy of it is returned (two small device copies
        instead of a re-capture per call)."""
        x = self._check_input(x)
        pool = self.__dict__.setdefault("_static_in", {})
        buf = pool.get(x.dtype)
        if buf is None:
            buf = pool[x.dtype] = torch.empty_like(x)
        buf.copy_(x)
        return self.forward(buf).clone()

    def forward_raw(self, x):
        """Pre-decode head logits (B, A, 64+nc) fp32 (reference training-mode output, nn.py:256-259)."""
        x = self._check_input(x)
        if self.raw is None:
            self.raw = torch.empty(self.batch, self.num_anchors, self.num_outputs + 60, dtype=torch.float32,
                                   device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.L.yb_forward_raw(self.plan, x.data_ptr(), _DTYPES[x.dtype], self.raw.data_ptr(),
                                         ctypes.c_void_p(stream)), "yb_forward_raw")
        return self.raw

    def describe(self):
        """The plan as a dict (buffers, ops, slices) — host-logic tests replay it on the CPU."""
        import json
        n = self.L.yb_plan_describe(self.plan, None, 0)
        buf = ctypes.create_string_buffer(int(n) + 16)
        _lib.check(int(self.L.yb_plan_describe(self.plan, buf, len(buf))), "yb_plan_describe")
        return json.loads(buf.value.decode())

    def debug_write(self, buf_index, tensor):
        """Overwrite activation buffer `buf_index` with `tensor` (already in the buffer's dtype/layout)."""
        t = tensor.contiguous()
        _lib.check(self.L.yb_plan_debug_write(self.plan, buf_index, t.data_ptr(), t.numel() * t.element_size()),
                   "yb_plan_debug_write")

    def run_op(self, op_index, x=None):
        """Run one op of the plan alone (teacher-forced per-op parity tests)."""
        stream = torch.cuda.current_stream(self.device).cuda_stream
        ptr, dt = (x.data_ptr(), _DTYPES[x.dtype]) if x is not None else (None, 0)
        _lib.check(self.L.yb_plan_run_op(self.plan, op_index, ptr, dt, self.out.data_ptr(),
                                         ctypes.c_void_p(stream)), "yb_plan_run_op")

    def debug_read(self, conv_name):
        """Activation written by op `conv_name` as an (B, H, W, C) fp32 CPU tensor (layer parity tests;
        meaningful only when the plan was created with YB_NO_REUSE=1)."""
        cap = self.batch * self.height * self.width * 64
        buf = np.empty(cap, dtype=np.float32)
        h, w, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        n = self.L.yb_plan_debug_read(self.plan, conv_name.encode(), buf.ctypes.data, cap, ctypes.byref(h),
                                      ctypes.byref(w), ctypes.byref(c))
        _lib.check(int(n), "yb_plan_debug_read")
        return torch.from_numpy(buf[:n].copy()).view(self.batch, -1, c.value)
