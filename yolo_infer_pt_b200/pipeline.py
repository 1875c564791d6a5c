"""Host-to-device streaming around the hot path: the loop the reference's `test()` runs
(main.py:264-273: `samples.to(device)`, `model(samples)`, `non_max_suppression(outputs)`), software
pipelined over four CUDA streams:

    copy    H2D of batches i+1, i+2   (pinned host memory -> one of three device input buffers)
    compute forward of batch i        (-> one of two prediction tensors)
    nms     NMS of batch i-1          (reads the other prediction tensor)
    d2h     detections of batch i-1   (-> pinned host landing buffers)

`StreamingDetector` owns the buffers and events.  Pinned host batches (or, with `resident=True`, batches that
already live on the device) go in, padded detections `(det (B, 300, 6) fp32, counts (B,) int32)` come out on
the host.  PyTorch provides memory, streams and events only; every kernel is libyolob200's.

`raw_frames=(h, w)`: the host batches are the camera / decoder frames themselves - (B, h, w, 3) uint8 HWC BGR, as
cv2.imread yields them - and the reference's resize + letterbox + BGR->RGB + HWC->CHW (utils/dataset.py:86-103,
292-313) runs on the device (`yb_letterbox`, bit-exact with cv2) between the copy and the forward: only the raw
pixels cross PCIe (a 480 x 640 frame is 25 % smaller than its 640 x 640 letterbox), which is what bounds the
end-to-end rate from one GPU on.
"""
import ctypes

import torch

from . import _lib
from .utils import util


class StreamingDetector:
    def __init__(self, model, batch_shape, dtype=torch.uint8, device="cuda", conf=0.001, iou=0.65, resident=False,
                 fuse_filter=True, input_buffers=3, raw_frames=None):
        self.model = model
        self.raw_frames = raw_frames
        if raw_frames is not None:
            if resident or dtype != torch.uint8:
                raise ValueError("raw_frames needs uint8 host batches")
            self.input_size = int(batch_shape[-1])
            if tuple(batch_shape[1:]) != (3, self.input_size, self.input_size):
                raise ValueError("batch_shape is the model input (B, 3, S, S); frames are (B, h, w, 3)")
        self.device = torch.device(device)
        self.conf, self.iou = conf, iou
        self.resident = resident
        self.copy_stream = torch.cuda.Stream(self.device)
        self.compute_stream = torch.cuda.Stream(self.device)
        self.nms_stream = torch.cuda.Stream(self.device)
        self.d2h_stream = torch.cuda.Stream(self.device)
        # input ring: with three buffers the copy engine always has a free target, so a copy that takes about
        # as long as a forward (PCIe-bound uint8 batches) never waits for the kernels, nor they for it
        self.n_in = 2 if resident else max(2, int(input_buffers))
        in_shape = batch_shape if raw_frames is None else (batch_shape[0], int(raw_frames[0]), int(raw_frames[1]), 3)
        self.inputs = [None] * self.n_in if resident else [torch.empty(in_shape, dtype=dtype, device=self.device)
                                                           for _ in range(self.n_in)]
        if raw_frames is not None:
            # letterboxed model inputs (one per input slot) and the per-image descriptors of the frames in each slot
            # (device pointer, height, width): the slots are static, so the descriptors are built once
            self.boxed = [torch.empty(batch_shape, dtype=torch.uint8, device=self.device) for _ in range(self.n_in)]
            frame_bytes = int(raw_frames[0]) * int(raw_frames[1]) * 3
            self.frame_desc = [torch.tensor([[buf.data_ptr() + i * frame_bytes, int(raw_frames[0]), int(raw_frames[1])]
                                             for i in range(batch_shape[0])], dtype=torch.int64).to(self.device)
                               for buf in self.inputs]
            self.frame_meta = [torch.empty((batch_shape[0], 3), dtype=torch.float64, device=self.device)
                               for _ in range(self.n_in)]
        self.copied = [torch.cuda.Event() for _ in range(self.n_in)]
        self.consumed = [torch.cuda.Event() for _ in range(self.n_in)]
        self.fwd_done = [torch.cuda.Event() for _ in range(2)]
        self.nms_done = [torch.cuda.Event() for _ in range(2)]
        for e in self.consumed + self.nms_done:
            e.record(self.compute_stream)
        self.preds = None                       # two prediction tensors, allocated at the first batch
        # fuse_filter: the forward's class-score epilogues apply the NMS confidence filter (one candidate
        # workspace per slot); the NMS then never re-reads the (B, nc, A) scores
        self.fuse_filter = fuse_filter
        self.nms_ws = None
        self.h2d_bytes = 0 if resident else self.inputs[0].numel() * self.inputs[0].element_size()
        b = batch_shape[0]
        # pinned host landing buffers for the detections: three slots, because the D2H copy of batch i+2 is
        # enqueued before result i+1 is handed out - a yielded result therefore stays valid until the consumer
        # has asked for the NEXT one (hold two results at most; clone to keep more, e.g. list(pipe.run(...)))
        self.det_host = [torch.empty(b, util.MAX_DET, 6, dtype=torch.float32).pin_memory() for _ in range(3)]
        self.cnt_host = [torch.empty(b, dtype=torch.int32).pin_memory() for _ in range(3)]

    def _upload(self, index, batch):
        islot = index % self.n_in
        if self.resident:
            self.inputs[islot] = batch          # already on the device, owned by the caller
            return
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[islot])     # the kernels that read this buffer are done
            self.inputs[islot].copy_(batch, non_blocking=True)
            self.copied[islot].record(self.copy_stream)

    def _detect(self, index):
        islot, slot = index % self.n_in, index & 1      # input ring slot; prediction / NMS workspace / landing slot
        x = self.inputs[islot]
        with torch.cuda.stream(self.compute_stream):
            if not self.resident:
                self.compute_stream.wait_event(self.copied[islot])
            self.compute_stream.wait_event(self.nms_done[slot])   # the NMS that read this prediction tensor two batches ago
            if self.raw_frames is not None:
                # resize + letterbox + BGR->RGB + HWC->CHW of the whole batch, one kernel (bit-exact with cv2)
                with torch.cuda.device(self.device):
                    _lib.check(_lib.lib().yb_letterbox(ctypes.c_void_p(self.frame_desc[islot].data_ptr()), x.shape[0],
                                                       self.input_size, ctypes.c_void_p(self.boxed[islot].data_ptr()),
                                                       ctypes.c_void_p(self.frame_meta[islot].data_ptr()),
                                                       ctypes.c_void_p(self.compute_stream.cuda_stream)), "yb_letterbox")
                x = self.boxed[islot]
            eng = self.model._engine_for(x)
            if self.preds is None:
                self.preds = [torch.empty_like(eng.out) for _ in range(2)]
                if self.fuse_filter:
                    self.nms_ws = [util.nms_workspace(eng.batch, eng.num_outputs - 4, eng.num_anchors, self.device)
                                   for _ in range(2)]
            sink = (self.nms_ws[slot], self.conf, util.MAX_NMS) if self.fuse_filter else None
            y = eng.forward(x, out=self.preds[slot], nms_sink=sink)
            self.consumed[islot].record(self.compute_stream)
            self.fwd_done[slot].record(self.compute_stream)
        with torch.cuda.stream(self.nms_stream):                  # overlaps the next batch's forward
            self.nms_stream.wait_event(self.fwd_done[slot])
            if self.fuse_filter:
                det, counts = util.nms_padded(y, self.conf, self.iou, workspace=self.nms_ws[slot], prefiltered=True)
            else:
                det, counts = util.nms_padded(y, self.conf, self.iou)
            self.nms_done[slot].record(self.nms_stream)
        # detections go back on their own stream: no kernel queues behind the copy
        det_h, cnt_h = self.det_host[index % 3], self.cnt_host[index % 3]
        with torch.cuda.stream(self.d2h_stream):
            self.d2h_stream.wait_event(self.nms_done[slot])
            det.record_stream(self.d2h_stream)
            counts.record_stream(self.d2h_stream)
            det_h.copy_(det, non_blocking=True)
            cnt_h.copy_(counts, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.d2h_stream)
        return det_h, cnt_h, done

    def run(self, batches):
        """Generator over batches -> (det, counts) pinned host tensors, in order.  The H2D copy of the next batch
        and the NMS of the previous one are in flight while the current batch's forward runs.  A yielded pair is
        a view of a landing buffer that is reused three batches later: it is valid while the consumer holds at
        most the current and the previous result; clone what must live longer."""
        it = iter(batches)
        ahead = []                 # indices uploaded (or uploading) and not yet run
        state = {"next": 0, "more": True}

        def fill():
            while state["more"] and len(ahead) < self.n_in - 1:
                try:
                    b = next(it)
                except StopIteration:
                    state["more"] = False
                    return
                self._upload(state["next"], b)
                ahead.append(state["next"])
                state["next"] += 1

        fill()
        pending = None
        while ahead:
            index = ahead.pop(0)
            fill()                 # the ring minus the batch about to run is in flight
            result = self._detect(index)
            if pending is not None:
                pending[2].synchronize()
                yield pending[0], pending[1]
            pending = result
        if pending is not None:
            pending[2].synchronize()
            yield pending[0], pending[1]
