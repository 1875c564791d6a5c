"""Two-rank GPU test of the sharded path: one process per GPU, images sharded contiguously, forward +
NMS on each rank through libyolob200, then the optional NCCL gather of the padded detections over NVLink
(parallel.gather_detections).  Rank 0's gathered result must equal a single-GPU run over the whole batch.
Needs two GPUs (`gpurun --gpus 2`); skipped on a single-GPU box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _model(dev):
    from yolo_infer_pt_b200 import synth
    from yolo_infer_pt_b200.nets import nn
    m = nn.yolo_v11_n(80)
    synth.load_synth(m, 0, "calibrated_damped")
    return m.fuse().eval().to(dev)


def _worker(rank, world, port, batch, out_dir):
    from yolo_infer_pt_b200 import parallel, synth
    from yolo_infer_pt_b200.utils import util
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    x = synth.synth_images(batch, 160, 160, seed=9)
    lo, hi = parallel.shard_bounds(batch, world, rank)
    with torch.no_grad():
        y = _model(dev)(x[lo:hi].to(dev))
    det, cnt = util.nms_padded(y, 0.25, 0.65)
    det_all, cnt_all = parallel.gather_detections(det, cnt)      # NCCL all_gather over NVLink
    slowest = parallel.max_over_ranks(1.0 + rank, dev)
    torch.cuda.synchronize(dev)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), det=det_all.cpu().numpy(), cnt=cnt_all.cpu().numpy(),
             slowest=slowest, backend=dist.get_backend())
    dist.destroy_process_group()


def test_two_gpu_sharded_forward_nms_and_nccl_gather(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from yolo_infer_pt_b200 import synth
    from yolo_infer_pt_b200.utils import util
    batch, world = 5, 2
    mp.spawn(_worker, args=(world, _free_port(), batch, str(tmp_path)), nprocs=world, join=True)
    dev = torch.device("cuda:0")
    with torch.no_grad():
        y = _model(dev)(synth.synth_images(batch, 160, 160, seed=9).to(dev))
    det, cnt = util.nms_padded(y, 0.25, 0.65)
    det, cnt = det.cpu().numpy(), cnt.cpu().numpy()
    assert cnt.sum() > 0
    for rank in range(world):
        g = np.load(tmp_path / f"rank{rank}.npz")
        assert str(g["backend"]) == "nccl" and float(g["slowest"]) == 2.0
        assert np.array_equal(g["cnt"], cnt)
        for b in range(batch):
            assert np.array_equal(g["det"][b, :cnt[b]], det[b, :cnt[b]])
