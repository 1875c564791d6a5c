#!/usr/bin/env python
"""Benchmark of the YOLO-Infer-pt inference hot path (YOLO.forward + non_max_suppression).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--model n|x|...]
                    [--batch B] [--size 640]

One step = forward + decode + NMS over one batch of synthetic images (random-init synthetic weights,
SURVEY.md §8(d) recipe).  Default workload (N=1): BASELINE.json configs[1] — YOLO11n, 16-bit
activations (fp16 default, bf16 in extra.bf16), batch 256 per GPU, 640x640.  Under torchrun every rank runs its own batch on its own
GPU (weak scaling, no data-path collective); times are CUDA-event times, max over ranks.

Prints ONE JSON line on rank 0 (see DESIGN.md §Measurement for every field).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "images/sec YOLO11n 640x640 forward+NMS"
REF_DIR = os.path.join(ROOT, "baseline", "_ref")   # unmodified copy of /root/reference/{nets,utils} (build())


def load_reference():
    """The UNMODIFIED reference (`nets.nn`, `utils.util` of t0saki/YOLO-Infer-pt) from baseline/_ref, where
    __graft_entry__.build() copied it (git-ignored; it travels to the GPU box with the snapshot).  Only the
    wall-clock bail-out of non_max_suppression is neutralised (utils/util.py:133-134,166-167: `time` -> a
    constant), as SURVEY.md 8c prescribes; nothing of this repo is on its path."""
    if not os.path.isfile(os.path.join(REF_DIR, "nets", "nn.py")):
        return None
    import importlib
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF_DIR)
    try:
        for name in ("nets", "nets.nn", "utils", "utils.util"):
            sys.modules.pop(name, None)
        ref_nn = importlib.import_module("nets.nn")
        ref_util = importlib.import_module("utils.util")
    finally:
        sys.path.remove(REF_DIR)
    assert ref_nn.__file__.startswith(REF_DIR) and ref_util.__file__.startswith(REF_DIR)
    ref_util.time = lambda: 0.0
    return ref_nn, ref_util


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=float(p["hbm_gbs"]), tf_burst=float(p["bf16_tflops"]),
                    tf_sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line):
    NVML from a thread every few ms (an `nvidia-smi -lms` child delivers its first line too late for a
    sub-second region); falls back to nvidia-smi when NVML is unavailable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.nvml = None
        self.handle = None
        self.samples = []
        self.running = False
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            try:
                uuid = str(torch.cuda.get_device_properties(gpu_index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            self.nvml, self.handle = pynvml, h
        except Exception:
            self.nvml = None

    def _poll(self):
        n, h = self.nvml, self.handle
        while self.running:
            try:
                sm = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
                try:
                    rs = n.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    rs = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((sm, rs))
            except Exception:
                pass
            time.sleep(0.004)

    def start(self):
        if self.nvml is not None:
            self.samples, self.running = [], True
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.running = False
            self.t.join(timeout=1)
            if not self.samples:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
            n = self.nvml
            try:
                mx = float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM))
            except Exception:
                mx = None
            reasons = set()
            for _, rs in self.samples:
                for bit, name in self.BITS.items():
                    if rs & bit:
                        reasons.add(name)
            return {"sm_mhz": float(np.median([s for s, _ in self.samples])), "sm_max_mhz": mx,
                    "reasons": sorted(reasons), "samples": len(self.samples), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi"}


# -------------------------------------------------------------------------------------------------
def conv_algorithmic_work(desc, batch):
    """Per op of the plan: algorithmic FLOPs (2*MAC) and bytes (input read once + output written once
    + weights once) for the whole batch."""
    work = []
    for op in desc["ops"]:
        flops = bytes_ = 0.0
        m = batch * op["Hout"] * op["Wout"]
        cout = op["dst"]["C"]
        if op["kind"] == 1:
            cin = sum(s["C"] for s in op["src"])
            flops = 2.0 * m * cout * cin * op["k"] ** 2
            in_elems = sum(batch * (op["Hin"] >> s["up"]) * (op["Win"] >> s["up"]) * s["C"] for s in op["src"])
            bytes_ = 2.0 * in_elems + (4.0 if op["out_f32"] else 2.0) * m * cout + 2.0 * cout * cin * op["k"] ** 2
            if op["has_res"]:
                bytes_ += 2.0 * m * cout
            # (a bottleneck whose `x +` was folded into this conv's weights: the reference's graph still reads x there -
            # the algorithmic figure stays SURVEY 8d's, the saved read shows up as time only)
            bytes_ += 2.0 * m * op.get("wfold", [0, 0, 0])[2]
        work.append((flops, bytes_))
    return work


def bind_to_gpu_numa_node(gpu_index):
    """One process per GPU: pin this process to the CPUs NVML reports as local to its GPU, so that the pinned
    host batches are allocated (first touch) on that NUMA node and the H2D copies do not cross sockets."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(gpu_index).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


def run_ours(args, rank, world, local_rank):
    from oracle import nms_oracle, yolo_oracle  # checker / CPU baseline only
    from yolo_infer_pt_b200 import _lib, synth
    from yolo_infer_pt_b200.nets import nn
    from yolo_infer_pt_b200.utils import util

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:   # (the single-GPU run keeps every core for its CPU-baseline leg)
        bind_to_gpu_numa_node(local_rank)   # before any pinned allocation: H2D then reads node-local host memory
    peaks = read_peaks()
    model = getattr(nn, f"yolo_v11_{args.model}")(80)
    synth.load_synth(model, 0, "survey")
    model = model.fuse().eval().to(dev)
    B, S = args.batch, args.size
    # synthetic uint8 images (what the reference's loader yields, main.py:265-267); /255 is fused in the stem
    base = (synth.synth_images(min(B, 8), S, S, seed=rank) * 255).round().to(torch.uint8)
    host = base.repeat((B + base.shape[0] - 1) // base.shape[0], 1, 1, 1)[:B].contiguous().pin_memory()
    x_dev = host.to(dev)
    L = _lib.lib()

    def step_resident():
        y = model(x_dev)
        return util.nms_padded(y, 0.001, 0.65)

    from yolo_infer_pt_b200.pipeline import StreamingDetector
    streamer = StreamingDetector(model, tuple(host.shape), torch.uint8, dev)
    resident = StreamingDetector(model, tuple(host.shape), torch.uint8, dev, resident=True)

    def run_pipeline(pipe, source, steps):
        n = 0
        for det, counts in pipe.run(source for _ in range(steps)):
            n += int(counts.shape[0])
        return n

    def timed_pipeline(pipe, source, steps):
        """K steps through the public streaming API between two CUDA events: the pipeline's streams fork from
        and join the current stream, so fill and drain of the pipeline are inside the timed region."""
        sampler = ClockSampler(local_rank)
        cur = torch.cuda.current_stream(dev)
        streams = (pipe.copy_stream, pipe.compute_stream, pipe.nms_stream, pipe.d2h_stream)
        barrier()
        torch.cuda.synchronize(dev)
        sampler.start()
        before = L.yb_launch_count()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(cur)
        for st in streams:
            st.wait_stream(cur)
        assert run_pipeline(pipe, source, steps) == B * steps
        for st in streams:
            cur.wait_stream(st)
        t1.record(cur)
        torch.cuda.synchronize(dev)
        barrier()
        clk = sampler.stop()
        t = t0.elapsed_time(t1)
        n_launch = L.yb_launch_count() - before
        if world > 1:
            tt = torch.tensor([t], device=dev)
            torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
            t = tt.item()
        return t, n_launch, clk

    def run_e2e(steps):
        """Public streaming API: pinned host uint8 batches in, host detections out; every step copies
        its own inputs H2D and its detections D2H (the next batch's copy overlaps this batch's kernels)."""
        n = 0
        for det, counts in streamer.run(host for _ in range(steps)):
            n += int(counts.shape[0])
        return n

    def barrier():
        if world > 1:
            torch.distributed.barrier()

    for _ in range(max(3, args.warmup)):
        step_resident()
    torch.cuda.synchronize(dev)
    eng = model._engine_for(x_dev)

    def timed(fn, steps, profile=False):
        sampler = ClockSampler(local_rank)
        was_graph = eng.graph
        if profile:
            if was_graph:   # (small batches replay a CUDA graph; the per-op events need stream launches)
                eng.use_graph(False)
            eng.profile(True)
        barrier()
        torch.cuda.synchronize(dev)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        before = L.yb_launch_count()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        barrier()
        clocks = sampler.stop()
        ms = e0.elapsed_time(e1)
        launches = L.yb_launch_count() - before
        op_ms = None
        if profile:
            op_ms, _ = eng.profile_read()
            eng.profile(False)
            if was_graph:
                eng.use_graph(True)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = t.item()
        return ms, launches, clocks, op_ms

    # headline: K clean steps (no per-op events in the stream, kernels overlap through programmatic
    # dependent launch exactly as in production); then K more steps with a CUDA event between every
    # op of the plan for the per-kernel roofline numbers (those boundaries serialise the kernels)
    # `value`: the streaming API on a batch that already lives in HBM (NMS of batch i-1 overlaps the forward of
    # batch i on its own stream); the plain back-to-back API calls on one stream are reported next to it
    ms_seq, _, _, _ = timed(step_resident, args.steps, profile=False)
    ms_prof, _, _, op_ms = timed(step_resident, args.steps, profile=True)
    run_pipeline(resident, x_dev, 2)
    ms, launches, clocks = timed_pipeline(resident, x_dev, args.steps)
    ms_per_step = ms / args.steps
    value = world * B * args.steps / (ms / 1e3)
    value_seq = world * B * args.steps / (ms_seq / 1e3)
    run_e2e(2)
    ms_e, _, _ = timed_pipeline(streamer, host, args.steps)
    e2e_value = world * B * args.steps / (ms_e / 1e3)

    # ---- roofline of the dominant kernel (the tcgen05 implicit-GEMM conv, all its launches) -------
    desc = eng.describe()
    work = conv_algorithmic_work(desc, B)
    conv_ms = sum(float(op_ms[i]) for i, op in enumerate(desc["ops"]) if op["kind"] == 1)
    conv_flops = sum(w[0] for w in work)
    conv_bytes = sum(w[1] for w in work)
    fwd_ms = float(op_ms.sum())
    t_flops = conv_flops / (peaks["tf_sustained"] * 1e12)
    t_bytes = conv_bytes / (peaks["hbm"] * 1e9)
    bound = "tensor" if t_flops > t_bytes else "hbm"
    if bound == "hbm":
        achieved = conv_bytes / (conv_ms / 1e3) / 1e9
        peak, unit = peaks["hbm"], "GB/s"
    else:
        achieved = conv_flops / (conv_ms / 1e3) / 1e12
        peak, unit = peaks["tf_sustained"], "TFLOP/s"
    n_conv = sum(1 for op in desc["ops"] if op["kind"] == 1)
    # per-layer roofline: every conv launch against max(flops / tensor peak, bytes / HBM peak)
    t_layers = sum(max(w[0] / (peaks["tf_sustained"] * 1e12), w[1] / (peaks["hbm"] * 1e9)) for w in work) * 1e3
    # DRAM bytes of the same launches from the committed ncu capture of this workload (profiles/)
    traffic, traffic_src = None, None
    try:
        import glob
        for tp in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")), reverse=True):
            tj = json.load(open(tp))
            if tj.get("model") == args.model and tj.get("batch") == B and tj.get("size") == S:
                traffic, traffic_src = round(tj["dram_bytes_per_step"] / 1e9, 3), os.path.relpath(tp, ROOT)
                break
    except Exception:
        pass
    roofline = {"bound": bound, "achieved": round(achieved, 2), "peak": peak, "unit": unit,
                "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_unit": "GB per step (all launches of the kernel)",
                "traffic_source": traffic_src, "kernel": "conv_gemm_tcgen05_kernel",
                "per_layer_roofline_ms": round(t_layers, 4), "frac_of_per_layer_roofline": round(t_layers / conv_ms, 4),
                "launches_per_step": n_conv, "kernel_ms_per_step": round(conv_ms, 4),
                "kernel_share_of_step": round(conv_ms / (ms_prof / args.steps), 4),
                "profiled_ms_per_step": round(ms_prof / args.steps, 4),
                "algorithmic_gbytes_per_step": round(conv_bytes / 1e9, 4),
                "algorithmic_tflop_per_step": round(conv_flops / 1e12, 4),
                "tensor_tflops_achieved": round(conv_flops / (conv_ms / 1e3) / 1e12, 2),
                "peak_source": peaks["source"] + (" (sustained)" if bound == "tensor" else "")}
    if rank == 0 and args.profile_json:
        rows = [dict(name=op["name"], kind=op["kind"], ms=float(op_ms[i]), gflop=work[i][0] / 1e9,
                     mbytes=work[i][1] / 1e6) for i, op in enumerate(desc["ops"])]
        json.dump(dict(model=args.model, batch=B, size=S, forward_ms=fwd_ms, step_ms=ms_per_step, ops=rows),
                  open(args.profile_json, "w"), indent=1)

    act_name = "bf16" if eng.act_dtype == torch.bfloat16 else "fp16"
    # ---- H2D ceiling: plain pinned cudaMemcpyAsync of the same batch, alone on the device (what e2e is bound by)
    h2d_gbs = None
    try:
        barrier()
        torch.cuda.synchronize(dev)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        x_dev.copy_(host, non_blocking=True)
        c0.record()
        for _ in range(5):
            x_dev.copy_(host, non_blocking=True)
        c1.record()
        torch.cuda.synchronize(dev)
        h2d_gbs = 5 * host.numel() / (c0.elapsed_time(c1) / 1e3) / 1e9
        if world > 1:   # all ranks copy at the same time: the per-rank share of the box's PCIe topology
            tt = torch.tensor([h2d_gbs], device=dev)
            torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MIN)
            h2d_gbs = tt.item()
    except Exception:
        h2d_gbs = None

    extra = {}
    # ---- the other activation storage type, same workload (BASELINE configs[1] says bf16; the default is fp16)
    if not args.no_extras:
        try:
            other = torch.bfloat16 if eng.act_dtype == torch.float16 else torch.float16
            model.set_activation_dtype(other)
            pipe2 = StreamingDetector(model, tuple(host.shape), torch.uint8, dev, resident=True)
            run_pipeline(pipe2, x_dev, 3)
            ms2, _, _ = timed_pipeline(pipe2, x_dev, max(3, args.steps // 3))
            extra["bf16" if other == torch.bfloat16 else "fp16"] = {
                "value": round(world * B * max(3, args.steps // 3) / (ms2 / 1e3), 2), "unit": "images/sec",
                "note": "same workload with the other 16-bit activation storage type"}
            del pipe2
            model.set_activation_dtype(None)
            model.invalidate_engine()
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            extra["other_dtype_error"] = f"{type(e).__name__}: {e}"[:200]
            model.set_activation_dtype(None)

    # ---- e2e from raw camera frames: (B, 480, 640, 3) uint8 BGR on the host, letterboxed to 640 x 640 on the device
    # (utils/dataset.py:86-103,292-313 moved behind the PCIe copy: 25 % fewer bytes per image cross the bus)
    if not args.no_extras and S == 640:
        try:
            fh, fw = 480, 640
            rng = np.random.RandomState(7 + rank)
            frames = torch.from_numpy(rng.randint(0, 256, (B, fh, fw, 3), dtype=np.uint8)).pin_memory()
            rawp = StreamingDetector(model, tuple(host.shape), torch.uint8, dev, raw_frames=(fh, fw))
            run_pipeline(rawp, frames, 2)
            ms_r, _, _ = timed_pipeline(rawp, frames, args.steps)
            extra["e2e_raw_frames"] = {
                "value": round(world * B * args.steps / (ms_r / 1e3), 2), "unit": "images/sec",
                "h2d_bytes_per_step": int(frames.numel()), "frame": f"{fh}x{fw}x3 uint8 BGR (HWC)",
                "h2d_achieved_gbs": round(frames.numel() * args.steps / (ms_r / 1e3) / 1e9, 2),
                "note": "host frames as cv2.imread yields them; resize + letterbox + BGR->RGB + HWC->CHW on the device "
                        "(yb_letterbox, bit-exact with cv2), then the same forward + NMS"}
            del rawp, frames
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            extra["e2e_raw_frames"] = {"error": f"{type(e).__name__}: {e}"[:200]}

    # ---- CPU baseline: the unmodified reference on the host cores, bounded sample (all threads, and 1 thread =
    # what the reference's setup_multi_processes would impose, utils/util.py:39-44) ----------------------------
    cpu = cpu1 = eager = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args.model, S, budget_s=12.0)
        cpu1 = cpu_baseline(args.model, S, budget_s=5.0, batch=1, threads=1)
        torch.set_num_threads(os.cpu_count() or 1)
    if rank == 0 and world == 1 and not args.no_extras:
        streamer = resident = None
        model.invalidate_engine()
        torch.cuda.empty_cache()
        eager = gpu_eager_baseline(args.model, B, S, dev, x_dev)
    if not args.no_extras:
        del x_dev
        model.invalidate_engine()
        torch.cuda.empty_cache()
        extra.update(run_extras(args, rank, world, dev, barrier, peaks))
    out = {
        "metric": METRIC.replace("YOLO11n", f"YOLO11{args.model}"), "value": round(value, 2), "unit": "images/sec",
        "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": round(ms_per_step, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": act_name, "data": "synthetic",
        "config": {"workload": f"YOLO11{args.model} fused, {act_name} activations + weights / fp32 accumulate + fp32 decode, "
                               f"batch {B} per GPU, {S}x{S}, forward + DFL decode + NMS (conf 0.001, IoU 0.65, max_det 300)",
                   "dtype_note": "north_star allows bf16/fp16 activations; fp16 is the reference's own eval dtype "
                                 "(main.py:251,266) and the one that meets the 0.5 px / 1e-2 gate on the falsifiable "
                                 "recipe (tests/test_gpu_parity.py); the bf16 number is in extra.bf16",
                   "weights": "random-init synthetic (SURVEY.md 8d recipe, seed 0)",
                   "input": "uint8 NCHW resident in HBM, /255 fused into the stem kernel; streaming API "
                            "(forward and NMS of consecutive batches overlap on two streams)",
                   "l2": f"inputs larger than L2 ({host.numel() / 1e6:.0f} MB images + "
                         f"{eng.workspace_bytes / 1e9:.1f} GB activation arena per step)",
                   "parallelism": f"image-sharded x{world}, no collective on the data path"},
        "roofline": roofline,
        "e2e": {"value": round(e2e_value, 2), "unit": "images/sec", "h2d_bytes_per_step": int(host.numel()),
                "d2h_bytes_per_step": int(B * (300 * 6 * 4 + 4)), "ms_per_step": round(ms_e / args.steps, 4),
                "h2d_ceiling_gbs": None if h2d_gbs is None else round(h2d_gbs, 2),
                "h2d_achieved_gbs": round(host.numel() * args.steps / (ms_e / 1e3) / 1e9, 2),
                "frac_of_h2d_ceiling": None if not h2d_gbs else round(host.numel() * args.steps / (ms_e / 1e3) / 1e9 / h2d_gbs, 4),
                "note": "pinned uint8 batch copied H2D and detections D2H every step; the ceiling is a plain pinned "
                        "cudaMemcpyAsync of the same batch with the GPU otherwise idle (min over ranks)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "forward_ms_per_step": round(fwd_ms, 4),
        "sequential_api": {"value": round(value_seq, 2), "ms_per_step": round(ms_seq / args.steps, 4),
                           "note": "model(x); non_max_suppression(y) back to back on one stream"},
    }
    if cpu is not None:
        out["cpu_baseline"] = cpu
        out["cpu_baseline_1thread"] = cpu1
    if eager is not None:
        out["gpu_eager_baseline"] = eager
        best = max((v.get("images_per_sec", 0.0) for v in eager.values() if isinstance(v, dict)), default=0.0)
        if best:
            out["vs_gpu_eager"] = {"value_ratio": round(value / best, 2), "e2e_ratio": round(e2e_value / best, 2),
                                   "against": "the faster of the reference's fp16 / fp32 eager runs (forward + its own NMS)"}
    if extra:
        out["extra"] = extra
    return out


def run_extras(args, rank, world, dev, barrier, peaks):
    """The other BASELINE.json configs, in the same JSON line: configs[3] YOLO11x, global batch 512 sharded by
    image over the ranks (strong scaling), and configs[2] batch-1 p50 latency of n / s / m with CUDA-graph replay."""
    from yolo_infer_pt_b200 import _lib, synth
    from yolo_infer_pt_b200.engine import Engine
    from yolo_infer_pt_b200.nets import nn
    from yolo_infer_pt_b200.utils import util
    extra = {}
    # ---- configs[3]: YOLO11x, B = 512 / world per GPU --------------------------------------------------------
    try:
        Bx = max(1, args.x_batch // world)
        mx = nn.yolo_v11_x(80)
        synth.load_synth(mx, 0, "survey")
        mx = mx.fuse().eval()
        eng = Engine(*mx._arch, Bx, args.size, args.size, dev)
        eng.pack_from_model(mx)
        base = (synth.synth_images(4, args.size, args.size, seed=rank) * 255).round().to(torch.uint8)
        xx = base.repeat((Bx + 3) // 4, 1, 1, 1)[:Bx].contiguous().to(dev)

        def step():
            return util.nms_padded(eng.forward(xx), 0.001, 0.65)

        for _ in range(3):
            step()
        steps = args.x_steps
        barrier()
        torch.cuda.synchronize(dev)
        sampler = ClockSampler(dev.index)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize(dev)
        barrier()
        clk = sampler.stop()
        ms = e0.elapsed_time(e1)
        eng.profile(True)
        for _ in range(2):
            step()
        op_ms, _ = eng.profile_read()
        eng.profile(False)
        if world > 1:
            tt = torch.tensor([ms], device=dev)
            torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
            ms = tt.item()
        desc = eng.describe()
        work = conv_algorithmic_work(desc, Bx)
        conv_ms = sum(float(op_ms[i]) for i, op in enumerate(desc["ops"]) if op["kind"] == 1)
        flops = sum(w[0] for w in work)
        t_layers = sum(max(w[0] / (peaks["tf_sustained"] * 1e12), w[1] / (peaks["hbm"] * 1e9)) for w in work) * 1e3
        tf = flops / (conv_ms / 1e3) / 1e12
        extra["x"] = {
            "metric": "images/sec YOLO11x 640x640 forward+NMS", "value": round(world * Bx * steps / (ms / 1e3), 2),
            "unit": "images/sec", "global_batch": Bx * world, "batch_per_gpu": Bx, "scaling": "strong", "steps": steps,
            "ms_per_step": round(ms / steps, 3), "dtype": "bf16" if eng.act_dtype == torch.bfloat16 else "fp16",
            "roofline": {"bound": "tensor", "achieved": round(tf, 2), "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                         "frac": round(tf / peaks["tf_sustained"], 4), "kernel": "conv_gemm_tcgen05_kernel",
                         "kernel_ms_per_step": round(conv_ms, 3), "per_layer_roofline_ms": round(t_layers, 3),
                         "frac_of_per_layer_roofline": round(t_layers / conv_ms, 4),
                         "peak_source": peaks["source"] + " (sustained)"},
            "clocks": clk}
        del eng, xx, mx
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001
        extra["x"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    # ---- configs[2]: batch-1 latency, CUDA-graph replay, wall-clock p50 (rank 0 of a single-GPU run) -----------
    if rank == 0 and world == 1:
        lat = {}
        for size in ("n", "s", "m"):
            try:
                m = getattr(nn, f"yolo_v11_{size}")(80)
                synth.load_synth(m, 0, "survey")
                m = m.fuse().eval().to(dev)
                x1 = (synth.synth_images(1, args.size, args.size) * 255).round().to(torch.uint8).to(dev)
                eng = m._engine_for(x1)        # B <= 8: the engine replays one CUDA graph per (input, output) pair
                for _ in range(30):
                    util.nms_padded(eng.forward(x1))
                    util.nms_padded(m(x1))
                torch.cuda.synchronize(dev)
                res = {}
                for name, fn in (("engine_forward_nms", lambda: util.nms_padded(eng.forward(x1))),
                                 ("engine_forward", lambda: eng.forward(x1)),
                                 ("model_call_nms", lambda: util.nms_padded(m(x1)))):
                    ws = []
                    for _ in range(args.latency_iters):
                        t0 = time.perf_counter()
                        fn()
                        torch.cuda.synchronize(dev)
                        ws.append((time.perf_counter() - t0) * 1e3)
                    res[name] = {"p50_ms": round(float(np.percentile(ws, 50)), 4), "p90_ms": round(float(np.percentile(ws, 90)), 4)}
                lat[size] = res
                del eng, m
                torch.cuda.empty_cache()
            except Exception as e:  # noqa: BLE001
                lat[size] = {"error": f"{type(e).__name__}: {e}"[:300]}
        lat["what"] = (f"batch 1, {args.size}x{args.size} uint8 image resident in HBM, wall clock around call + "
                       f"stream sync, {args.latency_iters} iterations; engine_*: Engine API (static buffers, one "
                       "CUDA graph for the forward, NMS kernels behind it); model_call_nms: the drop-in model(x) + "
                       "non_max_suppression-equivalent nms_padded (adds the weight fingerprint and two small copies)")
        extra["latency_p50_ms"] = lat
    return extra


def reference_model(ref_nn, model_size, device="cpu", dtype=torch.float32):
    from yolo_infer_pt_b200 import synth
    m = getattr(ref_nn, f"yolo_v11_{model_size}")(80)
    m.load_state_dict(synth.synth_state_dict(m, 0, "survey"))
    return m.fuse().eval().to(device=device, dtype=dtype)


def cpu_baseline(model_size, size, budget_s=12.0, batch=8, threads=None):
    """The reference's CPU path timed on the host cores: forward + non_max_suppression on batches of `batch`
    synthetic images until ~budget_s seconds of work were done.  kind "reference": the unmodified reference from
    baseline/_ref; kind "port" (only when baseline/_ref is absent): the oracle restatement of it."""
    from yolo_infer_pt_b200 import synth
    cores = threads or (os.cpu_count() or 1)
    torch.set_num_threads(cores)
    x = synth.synth_images(batch, size, size, seed=0)
    ref = load_reference()
    if ref is not None:
        ref_nn, ref_util = ref
        m = reference_model(ref_nn, model_size)
        kind, what = "reference", "unmodified reference from baseline/_ref (torch CPU ops + torchvision nms, NMS timer patched)"

        def step(xx):
            return ref_util.non_max_suppression(m(xx), 0.001, 0.65)
    else:
        from oracle import nms_oracle, yolo_oracle
        from yolo_infer_pt_b200.nets import nn
        mm = getattr(nn, f"yolo_v11_{model_size}")(80)
        synth.load_synth(mm, 0, "survey")
        mm.fuse()
        sd = {k: v.float() for k, v in mm.state_dict().items()}
        kind, what = "port", "fp32 oracle port (torch CPU ops + C NMS); baseline/_ref missing"

        def step(xx):
            return nms_oracle.non_max_suppression(yolo_oracle.forward(sd, *mm._arch, xx).numpy(), 0.001, 0.65)
    done, t_total = 0, 0.0
    with torch.no_grad():
        step(x[:1])  # warm-up
        while t_total < budget_s:
            t0 = time.perf_counter()
            step(x)
            t_total += time.perf_counter() - t0
            done += batch
    return {"value": round(done / t_total, 3), "unit": "images/sec", "cores": torch.get_num_threads(),
            "kind": kind, "sample": f"{done} images in batches of {batch}, {size}x{size}, fp32, {what}, {t_total:.1f} s"}


def gpu_eager_baseline(model_size, batch, size, dev, x_u8):
    """The real bar: the UNMODIFIED reference on this B200 through PyTorch eager (cuDNN / ATen /
    torchvision kernels), fp16 as its own test() runs it (main.py:251,266-267) and fp32; same weights, same
    uint8 batch.  Forward: median of 5 after 2 warm-ups.  non_max_suppression is the reference's per-image Python
    loop (several host syncs per image), timed on the first 16 images and scaled to the batch."""
    ref = load_reference()
    if ref is None:
        return {"unavailable": "baseline/_ref missing (run __graft_entry__.build() where /root/reference exists)"}
    ref_nn, ref_util = ref
    out = {}
    for name, dt in (("fp16", torch.float16), ("fp32", torch.float32)):
        try:
            m = reference_model(ref_nn, model_size, dev, dt)
            ts = []
            with torch.no_grad():
                for it in range(7):
                    torch.cuda.synchronize(dev)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    y = m(x_u8.to(dt) / 255.)
                    e1.record()
                    torch.cuda.synchronize(dev)
                    if it >= 2:
                        ts.append(e0.elapsed_time(e1))
                nb = min(16, batch)
                ref_util.non_max_suppression(y[:2], 0.001, 0.65)
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                dets = ref_util.non_max_suppression(y[:nb], 0.001, 0.65)
                torch.cuda.synchronize(dev)
                nms_ms_img = (time.perf_counter() - t0) * 1e3 / nb
            fwd = float(np.median(ts))
            out[name] = {"forward_ms": round(fwd, 3), "nms_ms_per_image": round(nms_ms_img, 3),
                         "images_per_sec": round(batch / ((fwd + nms_ms_img * batch) / 1e3), 1),
                         "forward_only_images_per_sec": round(batch / (fwd / 1e3), 1),
                         "kept_first_image": int(dets[0].shape[0])}
            del m, y
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001  (an OOM here must not lose the bench line)
            out[name] = {"error": f"{type(e).__name__}: {e}"[:200]}
    out["what"] = (f"unmodified reference (baseline/_ref) on this GPU, PyTorch eager, batch {batch}, {size}x{size}, "
                   "uint8 batch resident in HBM -> .to(dtype)/255 -> model -> non_max_suppression (timer patched)")
    return out


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (unmodified, from baseline/_ref; the
    oracle port only if that copy is missing), all host threads, same metric and workload, bounded samples."""
    if rank != 0:
        return None
    batch = 8
    per_step_budget = max(2.0, 60.0 / max(1, args.steps + args.warmup))
    cb = None
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_baseline(args.model, args.size, budget_s=0.5, batch=batch)
    vals = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cb = cpu_baseline(args.model, args.size, budget_s=per_step_budget, batch=batch)
        vals.append(cb["value"])
    v = float(np.mean(vals))
    cb["value"] = round(v, 3)
    one = cpu_baseline(args.model, args.size, budget_s=6.0, batch=1, threads=1)
    return {
        "impl": "reference", "metric": METRIC.replace("YOLO11n", f"YOLO11{args.model}"), "value": round(v, 3),
        "unit": "images/sec", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round((time.perf_counter() - t0) * 1e3 / max(1, args.steps), 2), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"YOLO11{args.model} fused fp32 on host CPU, {args.size}x{args.size}, forward + NMS; "
                               f"each step = a bounded sample (~{per_step_budget:.0f} s) in batches of {batch}"},
        "cpu_baseline": cb,
        "cpu_baseline_1thread": one,
        "e2e": {"value": round(v, 3), "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="n")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--size", type=int, default=640)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip extra.x / extra.latency / eager baseline / other dtype")
    ap.add_argument("--x-batch", type=int, default=512, help="global batch of the YOLO11x leg (configs[3])")
    ap.add_argument("--x-steps", type=int, default=3)
    ap.add_argument("--latency-iters", type=int, default=1000)
    ap.add_argument("--profile-json", default="")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        out = run_reference(args, rank, world)
        if out is not None:
            print(json.dumps(out), flush=True)
        return
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    out = run_ours(args, rank, world, local_rank)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
