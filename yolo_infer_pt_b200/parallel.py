"""Multi-GPU plumbing for the inference path.

Images are independent — eval BatchNorm is folded, attention and NMS are per image — so the batch is
partitioned into contiguous slices, one process per GPU, weights replicated, and there is NO
collective on the data path.  The only optional exchange is a gather of the fixed-size detections
(<= 300 x 6 fp32 + a count per image, 7.2 KB) onto every rank over NCCL / NVLink; with gloo the same
code runs on CPU tensors (tests/test_dist_cpu.py).
"""
import torch
import torch.distributed as dist


def shard_bounds(n_items, world, rank):
    """Contiguous slice [lo, hi) of rank `rank`: the first n % world ranks take one extra item."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_detections(det, counts, group=None):
    """det (b_local, max_det, 6) fp32, counts (b_local,) int32 -> the same for the whole batch, in
    global image order, on every rank.  Shards may differ in size by one image (padded for the
    collective, trimmed afterwards)."""
    world = dist.get_world_size(group)
    if world == 1:
        return det, counts
    n_local = torch.tensor([det.shape[0]], device=det.device, dtype=torch.int64)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    pad = max(sizes)
    det_p = torch.zeros((pad,) + tuple(det.shape[1:]), dtype=det.dtype, device=det.device)
    cnt_p = torch.zeros(pad, dtype=counts.dtype, device=counts.device)
    det_p[:det.shape[0]] = det
    cnt_p[:counts.shape[0]] = counts
    det_all = [torch.empty_like(det_p) for _ in range(world)]
    cnt_all = [torch.empty_like(cnt_p) for _ in range(world)]
    dist.all_gather(det_all, det_p, group=group)
    dist.all_gather(cnt_all, cnt_p, group=group)
    return (torch.cat([d[:n] for d, n in zip(det_all, sizes)]),
            torch.cat([c[:n] for c, n in zip(cnt_all, sizes)]))


def max_over_ranks(value, device, group=None):
    """Max of a python float over ranks (multi-GPU timings are the slowest rank's)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
