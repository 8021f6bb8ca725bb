"""Multi-GPU plumbing: one process per GPU (torchrun), reads sharded by contiguous ranges, the Bloom
filter replicated once, no steady-state collectives.

The reference is single-process (orpheum/translate.py:367-379 iterates the reads serially); reads are
independent, so the only cross-rank steps are (1) the one-off replication of the filter and (2) bringing
the per-read records back in global read order.  torch.distributed is only the transport.
"""
import numpy as np


def shard_bounds(n_items, world_size):
    """Contiguous, balanced [lo, hi) ranges: rank r owns items bounds[r]..bounds[r+1]."""
    base, extra = divmod(int(n_items), int(world_size))
    bounds = [0]
    for r in range(world_size):
        bounds.append(bounds[-1] + base + (1 if r < extra else 0))
    return bounds


def shard_reads(offsets, rank, world_size):
    """This rank's slice of a batch: (read lo, read hi, offsets view).  The offsets stay absolute -- the C ABI
    accepts offsets[0] != 0 -- so no bytes are copied or re-based."""
    n = len(offsets) - 1
    bounds = shard_bounds(n, world_size)
    lo, hi = bounds[rank], bounds[rank + 1]
    return lo, hi, offsets[lo:hi + 1]


class _CudaView:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def broadcast_filter(graph, src=0, group=None):
    """Replicate the filter from `src` to every rank: ONE NCCL broadcast of the single device allocation
    (NVLink / NVSwitch).  Every rank must hold a filter of the same geometry."""
    import torch
    import torch.distributed as dist

    ptr, nbytes, _ = graph.device_buffer()
    view = torch.as_tensor(_CudaView(ptr, nbytes), device=torch.device("cuda", graph.ctx.device))
    dist.broadcast(view, src=src, group=group)
    torch.cuda.synchronize(graph.ctx.device)


def gather_records(local, dst=0, group=None, device="cpu"):
    """Concatenate every rank's record array on `dst`, in rank order (= global read order for contiguous
    shards).  `local` is a 1-D numpy structured array; returns the full array on dst and None elsewhere."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = torch.device(device)
    raw = np.ascontiguousarray(local).view(np.uint8).reshape(-1)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([raw.size], dtype=torch.int64, device=dev), group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    mine = torch.zeros(cap, dtype=torch.uint8, device=dev)
    if raw.size:
        mine[: raw.size] = torch.from_numpy(raw.copy()).to(dev)
    bufs = [torch.zeros(cap, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == dst else None
    dist.gather(mine, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    parts = [bufs[r][: sizes[r]].cpu().numpy() for r in range(world)]
    return np.concatenate(parts).view(local.dtype)
