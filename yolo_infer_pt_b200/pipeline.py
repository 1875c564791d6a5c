"""Host-to-device streaming around the hot path: the loop the reference's `test()` runs
(main.py:264-273: `samples.to(device)`, `model(samples)`, `non_max_suppression(outputs)`), with the
H2D copy of batch i+1 overlapped with the kernels of batch i.

`StreamingDetector` owns two device input buffers, a copy stream and a compute stream.  Pinned host
batches go in, padded detections `(det (B, 300, 6) fp32, counts (B,) int32)` come out on the host.
PyTorch provides memory, streams and events only; every kernel is libyolob200's.
"""
import torch

from .utils import util


class StreamingDetector:
    def __init__(self, model, batch_shape, dtype=torch.uint8, device="cuda", conf=0.001, iou=0.65):
        self.model = model
        self.device = torch.device(device)
        self.conf, self.iou = conf, iou
        self.copy_stream = torch.cuda.Stream(self.device)
        self.compute_stream = torch.cuda.Stream(self.device)
        self.d2h_stream = torch.cuda.Stream(self.device)
        self.inputs = [torch.empty(batch_shape, dtype=dtype, device=self.device) for _ in range(2)]
        self.copied = [torch.cuda.Event() for _ in range(2)]
        self.consumed = [torch.cuda.Event() for _ in range(2)]
        for e in self.consumed:
            e.record(self.compute_stream)
        self.h2d_bytes = self.inputs[0].numel() * self.inputs[0].element_size()
        b = batch_shape[0]
        # pinned host landing buffers for the detections (two slots; a result stays valid for two steps)
        self.det_host = [torch.empty(b, util.MAX_DET, 6, dtype=torch.float32).pin_memory() for _ in range(2)]
        self.cnt_host = [torch.empty(b, dtype=torch.int32).pin_memory() for _ in range(2)]

    def _upload(self, slot, host_batch):
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[slot])      # the kernels that read this buffer are done
            self.inputs[slot].copy_(host_batch, non_blocking=True)
            self.copied[slot].record(self.copy_stream)

    def _detect(self, slot):
        with torch.cuda.stream(self.compute_stream):
            self.compute_stream.wait_event(self.copied[slot])
            y = self.model(self.inputs[slot])
            det, counts = util.nms_padded(y, self.conf, self.iou)
            self.consumed[slot].record(self.compute_stream)
            computed = torch.cuda.Event()
            computed.record(self.compute_stream)
        # detections go back on their own stream: the next batch's kernels do not queue behind the copy
        det_h, cnt_h = self.det_host[slot], self.cnt_host[slot]
        with torch.cuda.stream(self.d2h_stream):
            self.d2h_stream.wait_event(computed)
            det.record_stream(self.d2h_stream)
            counts.record_stream(self.d2h_stream)
            det_h.copy_(det, non_blocking=True)
            cnt_h.copy_(counts, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.d2h_stream)
        return det_h, cnt_h, done

    def run(self, host_batches):
        """Generator over pinned host batches -> (det, counts) host tensors, in order.  The H2D copy
        of the next batch is in flight while the current batch computes."""
        it = iter(host_batches)
        try:
            nxt = next(it)
        except StopIteration:
            return
        self._upload(0, nxt)
        slot, pending = 0, None
        while nxt is not None:
            try:
                following = next(it)
            except StopIteration:
                following = None
            if following is not None:
                self._upload(slot ^ 1, following)
            result = self._detect(slot)
            if pending is not None:
                pending[2].synchronize()
                yield pending[0], pending[1]
            pending = result
            nxt, slot = following, slot ^ 1
        pending[2].synchronize()
        yield pending[0], pending[1]
