"""GPU parity of the device-side letterbox (yb_letterbox) against the oracle, bit-exact, through the C ABI."""
import os

import numpy as np
import pytest
import torch

from oracle import letterbox_oracle as lo
from yolo_infer_pt_b200.utils import dataset

pytestmark = pytest.mark.gpu


def _run(images, S):
    dev = torch.device("cuda:0")
    out, meta = dataset.letterbox_batch([torch.from_numpy(im).to(dev) for im in images], S)
    torch.cuda.synchronize()
    return out.cpu().numpy(), meta.cpu().numpy()


def test_letterbox_bit_exact_vs_reference_fixtures(golden_dir):
    g = np.load(os.path.join(golden_dir, "letterbox_cases.npz"))
    S, n = int(g["input_size"]), int(g["n"])
    out, meta = _run([g[f"img{i}"] for i in range(n)], S)
    for i in range(n):
        assert np.array_equal(out[i], g[f"out{i}"]), f"case {i}"
        assert np.allclose(meta[i], g[f"meta{i}"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("S", [640, 320, 1280])
def test_letterbox_bit_exact_vs_oracle_random_shapes(S):
    rng = np.random.default_rng(S)
    shapes = [(480, 640), (640, 480), (S, S), (S, S // 2), (1, 1), (3, 5), (2, 700), (1080, 1920), (427, 640), (333, 500),
              (S - 1, S), (S + 1, S - 3)]
    shapes += [(int(rng.integers(1, 1500)), int(rng.integers(1, 1500))) for _ in range(12)]
    # images whose short side collapses to 0 pixels make the reference's cv2.resize raise
    shapes = [(h, w) for h, w in shapes if min(int(h * (S / max(h, w))), int(w * (S / max(h, w)))) >= 1]
    images = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in shapes]
    out, meta = _run(images, S)
    for i, im in enumerate(images):
        ref, m = lo.letterbox(im, S)
        assert np.array_equal(out[i], ref), f"shape {im.shape[:2]}"
        assert np.allclose(meta[i], np.array(m), rtol=0, atol=1e-12)


def test_letterbox_feeds_the_forward():
    """The uint8 batch is what YOLO.forward takes: same output as the float path on the same pixels."""
    from yolo_infer_pt_b200 import synth
    from yolo_infer_pt_b200.nets import nn
    rng = np.random.default_rng(3)
    images = [rng.integers(0, 256, (48, 64, 3), dtype=np.uint8), rng.integers(0, 256, (64, 40, 3), dtype=np.uint8)]
    x8, _ = dataset.letterbox_batch([torch.from_numpy(im).cuda() for im in images], 64)
    m = nn.yolo_v11_n(80)
    synth.load_synth(m, 0, "survey")
    m = m.fuse().eval().cuda()
    y8 = m(x8)
    yf = m(x8.float() / 255.0)
    assert torch.isfinite(y8).all()
    assert (y8[:, :4] - yf[:, :4]).abs().max().item() <= 0.5
    assert (y8[:, 4:] - yf[:, 4:]).abs().max().item() <= 1e-2


def test_letterbox_rejects_cpu_tensors_and_degenerate_images():
    with pytest.raises(RuntimeError):
        dataset.letterbox_batch([torch.zeros((4, 4, 3), dtype=torch.uint8)], 64)
    with pytest.raises(ValueError):   # 2 x 700 at 320: the resized height is 0 (cv2.resize raises in the reference)
        dataset.letterbox_batch([torch.zeros((2, 700, 3), dtype=torch.uint8, device="cuda:0")], 320)
