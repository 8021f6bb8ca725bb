"""The N>1 host logic on CPU: world_size-2 gloo processes shard a batch by contiguous ranges and bring the
per-read records back in global order (no GPU; the records are synthetic)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from orpheum_b200 import parallel
from orpheum_b200._lib import READ_RESULT_DTYPE


def test_shard_bounds_properties():
    for n in (0, 1, 2, 7, 100, 10_000_001):
        for world in (1, 2, 3, 4, 8):
            b = parallel.shard_bounds(n, world)
            assert b[0] == 0 and b[-1] == n and len(b) == world + 1
            sizes = np.diff(b)
            assert (sizes >= 0).all() and sizes.max() - sizes.min() <= 1
    offsets = np.arange(0, 11 * 150, 150, dtype=np.uint64)  # 10 reads
    lo, hi, view = parallel.shard_reads(offsets, 1, 4)
    assert (lo, hi) == (3, 6) and view.tolist() == [450, 600, 750, 900]


def _fake_records(lo, hi):
    rec = np.zeros(hi - lo, dtype=READ_RESULT_DTYPE)
    idx = np.arange(lo, hi)
    rec["n_kmers"][:, 0] = idx % 65521
    rec["n_hits"][:, 5] = (idx * 7) % 65521
    rec["category"][:, 2] = idx % 6
    rec["best_frame"] = (idx % 7) - 3
    return rec


def _worker(rank, world, port, n_reads, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        offsets = np.arange(n_reads + 1, dtype=np.uint64) * 150
        lo, hi, view = parallel.shard_reads(offsets, rank, world)
        assert len(view) == hi - lo + 1
        full = parallel.gather_records(_fake_records(lo, hi), dst=0)
        if rank == 0:
            np.save(out_path, full)
        else:
            assert full is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_reads", [1001, 1])
def test_two_rank_gather_preserves_global_order(tmp_path, n_reads):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "full.npy")
    mp.spawn(_worker, args=(2, port, n_reads, out), nprocs=2, join=True)
    full = np.load(out)
    assert full.tobytes() == _fake_records(0, n_reads).tobytes()
