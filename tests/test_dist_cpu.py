"""world_size-2 gloo test of the multi-GPU host logic: contiguous image shards, no
data-path collective, optional gather of the padded detections, max-over-ranks timing."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import nms_oracle
from yolo_infer_pt_b200 import parallel, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, batch, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pred = synth.synth_predictions(batch, 80, 2100, img=320, mode="sparse", seed=50)
    lo, hi = parallel.shard_bounds(batch, world, rank)
    # each rank post-processes only its own images (CPU oracle stands in for the device kernels here)
    dets = nms_oracle.non_max_suppression(pred[lo:hi], 0.001, 0.65)
    det = torch.zeros(hi - lo, 300, 6)
    cnt = torch.zeros(hi - lo, dtype=torch.int32)
    for i, d in enumerate(dets):
        det[i, :len(d)] = torch.from_numpy(d)
        cnt[i] = len(d)
    det_all, cnt_all = parallel.gather_detections(det, cnt)
    slowest = parallel.max_over_ranks(10.0 + rank, torch.device("cpu"))
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), det=det_all.numpy(), cnt=cnt_all.numpy(), slowest=slowest,
             lo=lo, hi=hi)
    dist.destroy_process_group()


def test_shard_bounds_cover_the_batch():
    for n in (1, 5, 8, 256, 511):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def test_two_rank_sharded_nms_equals_single_process(tmp_path):
    batch, world = 5, 2          # uneven shards: 3 + 2 images
    mp.spawn(_worker, args=(world, _free_port(), batch, str(tmp_path)), nprocs=world, join=True)
    pred = synth.synth_predictions(batch, 80, 2100, img=320, mode="sparse", seed=50)
    ref = nms_oracle.non_max_suppression(pred, 0.001, 0.65)
    for rank in range(world):
        g = np.load(tmp_path / f"rank{rank}.npz")
        assert g["det"].shape == (batch, 300, 6)
        assert float(g["slowest"]) == 11.0
        for b in range(batch):
            assert int(g["cnt"][b]) == len(ref[b])
            assert np.array_equal(g["det"][b, :len(ref[b])], ref[b])
