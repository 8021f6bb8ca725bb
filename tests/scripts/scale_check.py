#!/usr/bin/env python
"""cfg4 at scale on one GPU: Bloom build from a ~1e8-residue synthetic proteome (tablesize 1e8 and 1e9),
bits checked against the CPU oracle, then scoring of synthetic reads against the HBM-resident 1e9 filter
with a parity check on a prefix.  Prints one JSON line per stage.  Usage: python tests/scripts/scale_check.py [n_reads]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle as orc  # noqa: E402
from orpheum_b200 import gpu, sequence_encodings as enc, synth  # noqa: E402


def main():
    n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    t = time.perf_counter()
    residues, poffs = synth.make_proteome(n_proteins=187000, seed=1004)
    print(json.dumps({"stage": "proteome", "residues": int(poffs[-1]), "seconds": time.perf_counter() - t}), flush=True)
    lut = enc.encode_lut("protein")
    filters = {}
    for tablesize in (int(1e8), int(1e9)):
        g = gpu.GpuNodegraph(7, tablesize, 4)
        g.ctx.synchronize()
        t = time.perf_counter()
        g.add_peptides(residues, poffs, lut)
        dt = time.perf_counter() - t
        filters[tablesize] = g
        g_fast = gpu.GpuNodegraph(7, tablesize, 4)
        g_fast.ctx.synchronize()
        t = time.perf_counter()
        g_fast.add_peptides(residues, poffs, lut, count_unique=False)
        dt_fast = time.perf_counter() - t
        assert [g_fast.n_occupied(i) for i in range(4)] == [g.n_occupied(i) for i in range(4)]
        del g_fast
        # oracle on a prefix proteome (same geometry): every bit the prefix sets must be set in the full filter,
        # and a filter built from the prefix alone must match the oracle bit for bit
        n_prefix = int(np.searchsorted(poffs, 1_000_000))
        g_small = gpu.GpuNodegraph(7, tablesize, 4)
        g_small.add_peptides(residues, poffs[: n_prefix + 1], lut)
        o_small = orc.Nodegraph(7, tablesize, 4)
        o_small.add_peptides(residues, poffs[: n_prefix + 1], orc.encode_lut("protein"))
        pops_gpu = [g_small.n_occupied(i) for i in range(4)]
        pops_cpu = [o_small.n_occupied(i) for i in range(4)]
        tabs, _ = g_small.tables()
        identical = all((tabs[i] == o_small.table(i)).all() for i in range(4))
        print(json.dumps({"stage": "index_build", "tablesize": tablesize, "residues": int(poffs[-1]), "seconds_incl_h2d": dt, "seconds_without_unique_count": dt_fast,
                          "residues_per_s": float(poffs[-1]) / dt, "fill_table0": g.n_occupied(0) / g.hashsizes()[0],
                          "n_unique_kmers": g.n_unique_kmers(), "prefix_bits_identical_to_oracle": bool(identical),
                          "prefix_popcounts_equal": pops_gpu == pops_cpu}), flush=True)
        assert identical and pops_gpu == pops_cpu
    # scoring against the 500 MB (HBM-resident) filter
    reads = synth.make_reads_fixed(n_reads, residues, poffs, length=150, seed=3003)
    flat, offsets = synth.concat_fixed(reads)
    for tablesize, g in filters.items():
        g.ctx.set_kernel_timing(True)
        g.ctx.kernel_times()
        for _ in range(3):
            res = gpu.score_reads(g, flat, offsets, lut, 0.5)
        times = g.ctx.kernel_times()
        g.ctx.set_kernel_timing(False)
        n_chk = 50_000
        o = orc.Nodegraph(7, tablesize, 4)
        o.add_peptides(residues, poffs, orc.encode_lut("protein"))
        ora, _ = orc.score_reads(o, flat, offsets[: n_chk + 1], orc.encode_lut("protein"), 0.5, n_threads=0)
        scored = ora["n_kmers"] >= 0
        ok = bool((res["category"][:n_chk].astype(np.int64) == ora["category"]).all()
                  and (res["n_kmers"][:n_chk][scored] == ora["n_kmers"][scored]).all()
                  and (res["n_hits"][:n_chk][scored] == ora["n_hits"][scored]).all())
        kernel_ms = sum(v[0] for v in times.values()) / 3
        print(json.dumps({"stage": "score", "tablesize": tablesize, "reads": n_reads, "kernel_ms_per_pass": kernel_ms,
                          "kernel_reads_per_s": n_reads / (kernel_ms * 1e-3),
                          "kernel_ms": {k: v[0] / 3 for k, v in times.items()},
                          "probes_min_per_read": float(ora["probes_min"].sum()) / n_chk, "parity_prefix_ok": ok}), flush=True)
        assert ok


if __name__ == "__main__":
    main()
