#!/usr/bin/env python
"""Per-chunk timeline of orph_score_reads (ORPHEUM_B200_TRACE=1) for the three host-input routes of bench.py."""
import os
import sys

os.environ["ORPHEUM_B200_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from orpheum_b200 import _lib, gpu, sequence_encodings as enc, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
residues, poffs = synth.make_proteome(n_proteins=20000, seed=1001)
flat, offsets = synth.concat_fixed(synth.make_reads_fixed(n, residues, poffs, length=150, seed=2002))
lut = enc.encode_lut("protein")
ctx = gpu.default_context()
g = gpu.GpuNodegraph(7, int(1e8), 4, ctx=ctx)
g.add_peptides(residues, poffs, lut, count_unique=False)
packed, _ = gpu.pack_reads(flat, offsets)
pin = lambda a: (lambda t: (t.numpy().__setitem__(slice(None), a), t)[1])(torch.empty(a.size, dtype=torch.from_numpy(a[:1]).dtype).pin_memory())
h_reads, h_off, h_packed = pin(flat), pin(offsets.view(np.int64)), pin(packed.view(np.int64))
h_out = torch.empty(n * 32, dtype=torch.uint8).pin_memory()
lib = _lib.load()


def call(buf, fmt):
    _lib.check(lib.orph_score_reads(ctx._h, g._h, buf.data_ptr(), fmt, h_off.data_ptr(), n, lut.ctypes.data, 0.5, h_out.data_ptr()))


for name, threads, buf, fmt in (("ascii over PCIe only", 0, h_reads, _lib.READS_ASCII), ("hybrid host packing", -1, h_reads, _lib.READS_ASCII),
                                ("packed host input", 0, h_packed, _lib.READS_PACKED2)):
    ctx.set_host_threads(threads)
    for rep in range(3):
        print(f"==== {name}, repetition {rep}", file=sys.stderr, flush=True)
        call(buf, fmt)
