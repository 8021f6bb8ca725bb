#!/usr/bin/env python
"""File-to-records throughput of orph_stream_* on a synthetic four-line FASTQ on tmpfs; host-only ingest rate next to it.
    python tools/stream_bench.py [n_reads] [threads ...]"""
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from orpheum_b200 import _lib, gpu, sequence_encodings as enc, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
thread_list = [int(x) for x in sys.argv[2:]] or [0]
residues, poffs = synth.make_proteome(n_proteins=20000, seed=1001)
lut = enc.encode_lut("protein")
g = gpu.GpuNodegraph(7, int(1e8), 4)
g.add_peptides(residues, poffs, lut, count_unique=False)
reads = synth.make_reads_fixed(n, residues, poffs, length=150, seed=2002)
fq = "/dev/shm/orph_stream_bench.fq"
qual = b"I" * 150
with open(fq, "wb") as fh:
    for lo in range(0, n, 200_000):
        fh.write(b"".join(b"@read%d 1:N:0\n%s\n+\n%s\n" % (i, reads[i].tobytes(), qual) for i in range(lo, min(n, lo + 200_000))))
size = os.path.getsize(fq)
flat, offsets = synth.concat_fixed(reads)
want = gpu.score_reads(g, flat, offsets, lut, 0.5)
lib = _lib.load()
for threads in thread_list:
    best = 1e9
    for rep in range(3):
        h = ctypes.c_void_p()
        assert lib.orph_ingest_open(fq.encode(), threads, ctypes.byref(h)) == 0
        b = _lib.IngestBatch()
        t0 = time.perf_counter()
        while True:
            assert lib.orph_ingest_next(h, 1 << 20, ctypes.byref(b)) == 0
            if b.n_reads == 0:
                break
        best = min(best, time.perf_counter() - t0)
        lib.orph_ingest_close(h)
    print(f"ingest only, {threads} threads: {n / best / 1e6:.1f} M reads/s, {size / best / 1e9:.2f} GB/s of text", flush=True)
    for batch_reads in (1 << 19, 1 << 20):
        best = 1e9
        got = np.zeros(n, dtype=want.dtype)  # the caller's result array, pages touched before the clock starts
        for rep in range(3):
            t0 = time.perf_counter()
            at = 0
            with gpu.ScoringStream(g, fq, lut, 0.5, batch_reads=batch_reads, n_threads=threads) as st:
                for batch in st:
                    at = batch.copy_results_into(got, at)
            best = min(best, time.perf_counter() - t0)
        same = at == n and got.tobytes() == want.tobytes()
        print(f"stream, {threads} threads, batches of {batch_reads}: {n / best / 1e6:.1f} M reads/s, {size / best / 1e9:.2f} GB/s of text, identical {same}", flush=True)
os.remove(fq)
