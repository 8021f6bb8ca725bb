#!/usr/bin/env python
"""Randomised differential test: GPU path against the oracle over random filter geometries, alphabets, k-mer sizes,
read-length mixes (all kernel tiers), N / lower-case density, input formats and sub-batches.

    python tests/scripts/fuzz_parity.py [seconds] [seed]
Prints one line per round and a JSON summary; exits 1 on the first difference (after dumping the case).
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle as orc  # noqa: E402
from orpheum_b200 import _lib, gpu, sequence_encodings as enc, synth  # noqa: E402
from orpheum_b200.sequence_encodings import BEST_KSIZES  # noqa: E402

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def random_reads(rng, residues, poffs, n, ksize):
    """Mixed batch: coding / random / repeats / homopolymers of random lengths; a random fraction gets N, IUPAC or
    lower-case bytes."""
    tier = rng.choice(5, p=[0.35, 0.3, 0.2, 0.1, 0.05])
    longest_protein = int(np.diff(poffs.astype(np.int64)).max())
    max_len = min([150, 320, 700, 1600, 700][tier], 3 * (longest_protein - 2))
    flat, offsets = synth.make_reads_mixed(n, residues, poffs, ksize=ksize, seed=int(rng.integers(1 << 30)), min_len=1 + int(rng.integers(0, 60)),
                                           max_len=max_len)
    flat = flat.copy()
    if tier == 4:  # long reads (block-per-read kernel beyond 6144 nt): runs of 16 consecutive reads joined into one
        offsets = np.ascontiguousarray(np.concatenate([offsets[:-1:16], offsets[-1:]]))
        n = len(offsets) - 1
    dirty_frac = float(rng.choice([0.0, 0.01, 0.1, 0.5]))
    if dirty_frac:
        for r in np.nonzero(rng.random(n) < dirty_frac)[0]:
            a, b = int(offsets[r]), int(offsets[r + 1])
            if b <= a:
                continue
            kind = rng.integers(0, 4)
            if kind == 0:
                flat[rng.integers(a, b, rng.integers(1, 6))] = ord("N")
            elif kind == 1:
                flat[a:b] = np.frombuffer(flat[a:b].tobytes().lower(), dtype=np.uint8)
            elif kind == 2:
                flat[rng.integers(a, b, rng.integers(1, 4))] = np.frombuffer(b"RYKMSWBDHVnacgt*-", dtype=np.uint8)[rng.integers(0, 17)]
            else:
                c = int(rng.integers(a, b))
                flat[c:min(b, c + int(rng.integers(1, 30)))] = ord("N")
    return flat, offsets


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    ctx = gpu.default_context()
    alphabets = list(BEST_KSIZES)
    t_end = time.time() + seconds
    rounds = reads_total = 0
    while time.time() < t_end:
        alphabet = str(rng.choice(alphabets))
        ksize = int(rng.choice([BEST_KSIZES[alphabet], BEST_KSIZES[alphabet], int(rng.integers(1, 41))]))
        n_tables = int(rng.choice([1, 2, 3, 4, 4, 4, 7]))
        tablesize = int(rng.choice([3e3, 1e5, 1e6, 3e7]))
        thr = float(rng.choice([0.5, 0.8, 0.0, 1.0, 0.05, rng.random()]))
        residues, poffs = synth.make_proteome(n_proteins=int(rng.integers(20, 400)), seed=int(rng.integers(1 << 30)))
        lut, olut = enc.encode_lut(alphabet), orc.encode_lut(alphabet)
        gg = gpu.GpuNodegraph(ksize, tablesize, n_tables, ctx=ctx)
        gg.add_peptides(residues, poffs, lut)
        go = orc.Nodegraph(ksize, tablesize, n_tables)
        go.add_peptides(residues, poffs, olut)
        assert [gg.n_occupied(i) for i in range(n_tables)] == [go.n_occupied(i) for i in range(n_tables)], "filter bits differ"
        if tablesize <= 1e6:  # the tables byte for byte
            tabs, _ = gg.tables()
            assert all((tabs[i] == go.table(i)).all() for i in range(n_tables)), "filter tables differ"
        n = int(rng.choice([1, 33, 1000, 20_000]))
        flat, offsets = random_reads(rng, residues, poffs, n, ksize)
        n = len(offsets) - 1
        lengths = np.diff(offsets.astype(np.int64))
        keep = lengths > 0  # an empty read is an error in both implementations (tested elsewhere)
        if not keep.all():
            seqs = [flat[int(offsets[i]):int(offsets[i + 1])].tobytes() for i in np.nonzero(keep)[0]]
            flat, offsets = gpu.concat_reads(seqs)
            n = len(seqs)
        if n == 0:
            continue
        ctx.set_force_exact_dedup(bool(rng.random() < 0.1))
        ctx.set_thread_path(bool(rng.random() < 0.8))
        ctx.set_host_threads(int(rng.choice([-1, 0, 2])))
        try:
            lo = int(rng.integers(0, n)) if rng.random() < 0.3 else 0
            res = gpu.score_reads(gg, flat, offsets[lo:], lut, thr)
            packed, needs = gpu.pack_reads(flat, offsets)
            res_p = None
            if not needs.any():
                res_p = gpu.score_reads(gg, packed, offsets, lut, thr, reads_format=_lib.READS_PACKED2)
        finally:
            ctx.set_force_exact_dedup(False)
            ctx.set_thread_path(True)
            ctx.set_host_threads(-1)
        ora, _ = orc.score_reads(go, flat, offsets, olut, thr, n_threads=0)
        scored = ora["n_kmers"] >= 0
        for name, got, o in (("ascii", res, ora[lo:]), ("packed", res_p, ora)):
            if got is None:
                continue
            sc = o["n_kmers"] >= 0
            ok = ((got["category"].astype(np.int64) == o["category"]).all() and (got["n_kmers"][sc] == o["n_kmers"][sc]).all()
                  and (got["n_hits"][sc] == o["n_hits"][sc]).all())
            if not ok:
                bad = np.nonzero((got["category"].astype(np.int64) != o["category"]).any(axis=1) | (got["n_kmers"] * sc != o["n_kmers"] * sc).any(axis=1)
                                 | (got["n_hits"] * sc != o["n_hits"] * sc).any(axis=1))[0]
                i = int(bad[0]) + (lo if name == "ascii" else 0)
                print(json.dumps({"FAIL": name, "round": rounds, "alphabet": alphabet, "ksize": ksize, "n_tables": n_tables, "tablesize": tablesize,
                                  "thr": thr, "read_index": i, "read": flat[int(offsets[i]):int(offsets[i + 1])].tobytes().decode("latin-1"),
                                  "gpu": {k: got[int(bad[0])][k].tolist() for k in ("category", "n_kmers", "n_hits")},
                                  "oracle": {k: o[int(bad[0])][k].tolist() for k in ("category", "n_kmers", "n_hits")}}))
                sys.exit(1)
        # the peptides the reference would emit (coding / low-complexity frames of a few reads), translated on the GPU
        if lo == 0:
            cat = res["category"]
            ri, fi = np.nonzero((cat[:200] == _lib.CAT_CODING) | (cat[:200] == _lib.CAT_LOW_COMPLEXITY_PEPTIDE))
            if len(ri):
                keys = np.asarray(_lib.FRAMES, dtype=np.int8)[fi]
                peps = gpu.translate_frames(ctx, flat, offsets, ri, keys)
                for r, k, pep in zip(ri.tolist(), keys.tolist(), peps):
                    want = orc.six_frame_translation(flat[int(offsets[r]):int(offsets[r + 1])].tobytes())[k]
                    assert pep == want, ("peptide differs", r, k)
        rounds += 1
        reads_total += n
        print(f"round {rounds}: {alphabet} k={ksize} tables={n_tables}x{tablesize:.0e} thr={thr:.3f} reads={n} max_len={int(lengths.max())} "
              f"scored_frames={int(scored.sum())} byte_path={int((res['flags'] & 0x80 != 0).sum())} ok", flush=True)
        del gg, go
    print(json.dumps({"fuzz_rounds": rounds, "reads": reads_total, "seconds": seconds, "seed": seed, "differences": 0}))


if __name__ == "__main__":
    main()
