"""The forward's class-score epilogues can apply non_max_suppression's confidence filter themselves
(yb_forward_nms + yb_nms_prefiltered).  Results must be identical, bit for bit, to yb_forward + yb_nms:
same prediction tensor, same detections, for sparse and overflowing candidate lists, for M tiles that
straddle two images, across reuse of the workspace, and through the streaming pipeline."""
import numpy as np
import pytest
import torch

from oracle import nms_oracle
from yolo_infer_pt_b200 import synth
from yolo_infer_pt_b200.nets import nn
from yolo_infer_pt_b200.pipeline import StreamingDetector
from yolo_infer_pt_b200.utils import util

pytestmark = pytest.mark.gpu


def _model(size="n"):
    m = getattr(nn, f"yolo_v11_{size}")(80)
    synth.load_synth(m, 0, "survey")
    return m.fuse().eval().to("cuda:0")


@pytest.mark.parametrize("hw,batch,conf", [(320, 3, 0.001), (320, 3, 0.25), (320, 2, -1.0), (64, 5, 0.001), (640, 2, 0.001)])
def test_fused_filter_equals_separate_nms(hw, batch, conf):
    model = _model()
    x = (synth.synth_images(batch, hw, hw, seed=3) * 255).round().to(torch.uint8).to("cuda:0")
    with torch.no_grad():
        eng = model._engine_for(x)
        y_ref = eng.forward(x).clone()
        det_ref, cnt_ref = util.nms_padded(y_ref, conf, 0.65)
        ws = util.nms_workspace(eng.batch, eng.num_outputs - 4, eng.num_anchors, x.device)
        out2 = torch.empty_like(y_ref)
        for rep in range(3):   # the NMS leaves the workspace ready for the next forward
            y = eng.forward(x, out=out2, nms_sink=(ws, conf, util.MAX_NMS))
            det, cnt = util.nms_padded(y, conf, 0.65, workspace=ws, prefiltered=True)
            torch.cuda.synchronize()
            assert torch.equal(y, y_ref), "the sink changed the prediction tensor"
            assert torch.equal(cnt, cnt_ref), f"rep {rep}: counts differ"
            assert torch.equal(det, det_ref), f"rep {rep}: detections differ"
    # and the detections are the oracle's (a few images)
    want = nms_oracle.non_max_suppression(y_ref[:2].cpu().numpy(), float(np.float32(conf)), 0.65)
    for b, w in enumerate(want):
        assert int(cnt_ref[b]) == len(w)
        assert np.array_equal(det_ref[b, :len(w)].cpu().numpy(), w)


def test_sink_and_plain_forward_alternate():
    """Graph replay keyed on the sink: alternating calls with / without a sink stay correct."""
    model = _model()
    x = (synth.synth_images(2, 320, 320, seed=5) * 255).round().to(torch.uint8).to("cuda:0")
    with torch.no_grad():
        eng = model._engine_for(x)
        ws = [util.nms_workspace(eng.batch, eng.num_outputs - 4, eng.num_anchors, x.device) for _ in range(2)]
        y0 = eng.forward(x).clone()
        d0, c0 = util.nms_padded(y0, 0.001, 0.65)
        for i in range(4):
            y = eng.forward(x, nms_sink=(ws[i & 1], 0.001, util.MAX_NMS))
            d, c = util.nms_padded(y, 0.001, 0.65, workspace=ws[i & 1], prefiltered=True)
            assert torch.equal(d, d0) and torch.equal(c, c0)
            y = eng.forward(x)
            d, c = util.nms_padded(y, 0.001, 0.65)
            assert torch.equal(d, d0) and torch.equal(c, c0)


def test_streaming_detector_fused_equals_unfused():
    model = _model()
    batches = [(synth.synth_images(4, 320, 320, seed=s) * 255).round().to(torch.uint8).pin_memory() for s in range(5)]
    res = {}
    for fuse in (False, True):
        pipe = StreamingDetector(model, tuple(batches[0].shape), fuse_filter=fuse)
        res[fuse] = [(d.clone(), c.clone()) for d, c in pipe.run(batches)]
    assert len(res[True]) == len(batches)
    for (d0, c0), (d1, c1) in zip(res[False], res[True]):
        assert torch.equal(c0, c1)
        for b in range(c0.numel()):
            assert torch.equal(d0[b, :int(c0[b])], d1[b, :int(c1[b])])


def test_prefiltered_nms_rejects_mismatched_threshold_or_tensor():
    model = _model()
    x = (synth.synth_images(2, 64, 64, seed=1) * 255).round().to(torch.uint8).to("cuda:0")
    with torch.no_grad():
        eng = model._engine_for(x)
        ws = util.nms_workspace(eng.batch, eng.num_outputs - 4, eng.num_anchors, x.device)
        with pytest.raises(ValueError):
            util.nms_padded(eng.forward(x), 0.001, 0.65, workspace=ws, prefiltered=True)     # no sink was attached
        y = eng.forward(x, nms_sink=(ws, 0.001, util.MAX_NMS))
        with pytest.raises(ValueError):
            util.nms_padded(y, 0.25, 0.65, workspace=ws, prefiltered=True)                   # other threshold
        with pytest.raises(ValueError):
            util.nms_padded(y.clone(), 0.001, 0.65, workspace=ws, prefiltered=True)          # other tensor
        util.nms_padded(y, 0.001, 0.65, workspace=ws, prefiltered=True)                      # consumes (and re-zeroes) the lists


def test_raw_frame_pipeline_letterboxes_on_the_device():
    """StreamingDetector(raw_frames=(h, w)): HWC BGR frames in, the device-side letterbox in front of the forward;
    detections equal those of the same frames letterboxed by utils.dataset.letterbox_batch and fed as NCHW batches."""
    from yolo_infer_pt_b200.pipeline import StreamingDetector
    from yolo_infer_pt_b200.utils import dataset
    model = nn.yolo_v11_n(80)
    synth.load_synth(model, 0, "survey")
    model = model.fuse().eval().to("cuda:0")
    rng = np.random.RandomState(3)
    B, h, w, S = 3, 96, 128, 128
    frames = [torch.from_numpy(rng.randint(0, 256, (B, h, w, 3)).astype(np.uint8)).pin_memory() for _ in range(4)]
    raw = StreamingDetector(model, (B, 3, S, S), torch.uint8, "cuda:0", raw_frames=(h, w))
    got = [(d.clone(), c.clone()) for d, c in raw.run(frames)]
    boxed = []
    for f in frames:
        out, _ = dataset.letterbox_batch([im for im in f.to("cuda:0")], S)
        boxed.append(out.cpu().pin_memory())
    plain = StreamingDetector(model, (B, 3, S, S), torch.uint8, "cuda:0")
    want = [(d.clone(), c.clone()) for d, c in plain.run(boxed)]
    assert len(got) == len(want) == 4
    for (d0, c0), (d1, c1) in zip(got, want):
        assert torch.equal(c0, c1)
        for i in range(B):
            assert torch.equal(d0[i, :c0[i]], d1[i, :c1[i]])
