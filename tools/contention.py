#!/usr/bin/env python
"""Does PCIe traffic slow the scoring kernels (and vice versa)?  Kernel time alone vs under concurrent H2D / D2H copies,
for an L2-resident 50 MB filter and for a tiny one."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from orpheum_b200 import _lib, gpu, sequence_encodings as enc, synth

n = 2_000_000
residues, poffs = synth.make_proteome(n_proteins=20000, seed=1001)
flat, offsets = synth.concat_fixed(synth.make_reads_fixed(n, residues, poffs, length=150, seed=2002))
lut = enc.encode_lut("protein")
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
ctx = gpu.Context(0, stream=stream.cuda_stream)
packed, _ = gpu.pack_reads(flat, offsets)
d_reads = torch.from_numpy(packed.view(np.int64)).to(dev)
d_off = torch.from_numpy(offsets.view(np.int64)).to(dev)
d_out = torch.zeros(n * 32, dtype=torch.uint8, device=dev)
h_buf = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
d_buf = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
copy_stream = torch.cuda.Stream(device=dev)


def kernels(g, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(reps):
            gpu.score_reads_device(g, d_reads.data_ptr(), d_off.data_ptr(), n, 150, lut, 0.5, d_out.data_ptr(), ctx=ctx)
        e1.record(stream)
    return e0, e1


for tablesize in (1e8, 1e6, 1e9):
    g = gpu.GpuNodegraph(7, int(tablesize), 4, ctx=ctx)
    g.add_peptides(residues, poffs, lut, count_unique=False)
    e0, e1 = kernels(g, 5)
    torch.cuda.synchronize()
    e0, e1 = kernels(g, 20)
    torch.cuda.synchronize()
    alone = e0.elapsed_time(e1) / 20
    for direction in ("h2d", "d2h", "both"):
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_copies = 40
        with torch.cuda.stream(copy_stream):
            c0.record(copy_stream)
            for _ in range(n_copies):
                if direction in ("h2d", "both"):
                    d_buf.copy_(h_buf, non_blocking=True)
                if direction in ("d2h", "both"):
                    h_buf.copy_(d_buf, non_blocking=True)
            c1.record(copy_stream)
        e0, e1 = kernels(g, 20)
        torch.cuda.synchronize()
        copy_ms = c0.elapsed_time(c1)
        nbytes = n_copies * (64 << 20) * (2 if direction == "both" else 1)
        print(f"tablesize {tablesize:.0e}: kernels alone {alone:.3f} ms / 2M reads; with {direction} copies {e0.elapsed_time(e1) / 20:.3f} ms "
              f"(copies: {nbytes / copy_ms / 1e6:.1f} GB/s over {copy_ms:.1f} ms)", flush=True)
    # copies alone
    for direction in ("h2d", "d2h"):
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(copy_stream):
            c0.record(copy_stream)
            for _ in range(20):
                (d_buf.copy_(h_buf, non_blocking=True) if direction == "h2d" else h_buf.copy_(d_buf, non_blocking=True))
            c1.record(copy_stream)
        torch.cuda.synchronize()
        print(f"   {direction} alone: {20 * (64 << 20) / c0.elapsed_time(c1) / 1e6:.1f} GB/s", flush=True)
    del g
