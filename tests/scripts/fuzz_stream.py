#!/usr/bin/env python
"""Randomised differential test of the file-to-records path: random reads (every length tier, N / IUPAC / lower-case bytes)
written as FASTQ in a random layout (plain, CRLF, no final newline, one-member gzip, BGZF with random block sizes, wrapped,
FASTA), streamed through orph_stream_* with a random batch size and 1-3 lanes, against (a) the ORACLE on the same reads,
(b) the batch call orph_score_reads, (c) for the emitted FASTA text, the oracle's six-frame translation of the frames the
oracle calls coding.

    python tests/scripts/fuzz_stream.py [seconds] [seed]
Prints one line per round and a JSON summary; exits 1 on the first difference.
"""
import gzip
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from fuzz_parity import random_reads  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from orpheum_b200 import _lib, gpu, sequence_encodings as enc, synth  # noqa: E402
from orpheum_b200.sequence_encodings import BEST_KSIZES  # noqa: E402


def write_file(path, layout, names, seqs, rng):
    eol = b"\r\n" if layout == "crlf" else b"\n"
    parts = []
    if layout == "fasta":
        for name, s in zip(names, seqs):
            w = int(rng.integers(20, 200))
            parts.append(b">" + name + eol + eol.join(s[j:j + w] for j in range(0, len(s), w)) + eol)
    else:
        wrap = int(rng.integers(30, 120)) if layout in ("wrapped", "bgzf_wrapped") else 0
        for name, s in zip(names, seqs):
            q = rng.integers(33, 74, len(s)).astype(np.uint8).tobytes()
            if wrap:
                s_txt = eol.join(s[j:j + wrap] for j in range(0, len(s), wrap))
                q_txt = eol.join(q[j:j + wrap] for j in range(0, len(q), wrap))
            else:
                s_txt, q_txt = s, q
            parts.append(b"@" + name + eol + s_txt + eol + b"+" + eol + q_txt + eol)
    blob = b"".join(parts)
    if layout == "no_final_newline":
        blob = blob[: -len(eol)]
    if layout == "gz":
        with gzip.open(path, "wb", compresslevel=1) as f:
            f.write(blob)
    elif layout in ("bgzf", "bgzf_wrapped"):
        synth.write_bgzf(path, blob, block=int(rng.choice([997, 8192, 65280])), eof_marker=bool(rng.random() < 0.7))
    else:
        with open(path, "wb") as f:
            f.write(blob)


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    ctx = gpu.default_context()
    tmp = tempfile.mkdtemp(prefix="orph_fuzz_stream_")
    t_end = time.time() + seconds
    rounds = reads_total = 0
    layouts = ["plain", "plain", "crlf", "no_final_newline", "gz", "bgzf", "bgzf", "wrapped", "fasta", "bgzf_wrapped"]
    while time.time() < t_end:
        alphabet = str(rng.choice(["protein", "dayhoff", "hp", "botvinnik"]))
        ksize = int(rng.choice([BEST_KSIZES[alphabet], int(rng.integers(2, 20))]))
        n_tables, tablesize = int(rng.choice([1, 2, 4])), int(rng.choice([1e5, 1e6, 3e7]))
        thr = float(rng.choice([0.5, 0.8, 0.05]))
        residues, poffs = synth.make_proteome(n_proteins=int(rng.integers(20, 400)), seed=int(rng.integers(1 << 30)))
        lut, olut = enc.encode_lut(alphabet), orc.encode_lut(alphabet)
        gg = gpu.GpuNodegraph(ksize, tablesize, n_tables, ctx=ctx)
        gg.add_peptides(residues, poffs, lut)
        go = orc.Nodegraph(ksize, tablesize, n_tables)
        go.add_peptides(residues, poffs, olut)
        n = int(rng.choice([1, 40, 3000, 60_000]))
        flat, offsets = random_reads(rng, residues, poffs, n, ksize)
        lengths = np.diff(offsets.astype(np.int64))
        seqs = [flat[int(offsets[i]):int(offsets[i + 1])].tobytes() for i in np.nonzero(lengths > 0)[0]]
        # bytes a FASTQ / FASTA line cannot hold are not reads (the parsers strip white space)
        seqs = [s for s in seqs if not any(c in s for c in b" \t\r\n>@+")] or [b"ACGTACGTAC"]
        n = len(seqs)
        names = [b"r%d fuzz/%d" % (i, i % 3) for i in range(n)]
        flat, offsets = gpu.concat_reads(seqs)
        layout = str(rng.choice(layouts))
        path = os.path.join(tmp, "reads.fq.gz" if layout in ("gz", "bgzf", "bgzf_wrapped") else "reads.fa" if layout == "fasta" else "reads.fq")
        write_file(path, layout, names, seqs, rng)
        n_lanes = int(rng.choice([1, 1, 2, 3]))
        lanes = gpu.replica_lanes(gg, n_lanes, devices=[0] * (n_lanes - 1)) if n_lanes > 1 else None
        batch_reads = int(rng.choice([4096, 5000, 65536, 1 << 20]))
        kinds = int(rng.choice([1, 15, 9, 7]))
        got_parts, got_names, emitted = [], [], [b"", b"", b"", b""]
        with gpu.ScoringStream(gg, path, lut, thr, batch_reads=batch_reads, n_threads=int(rng.choice([0, 1, 3])), replicas=lanes) as st:
            for batch in st:
                assert batch.status == 0, batch.status
                got_parts.append(batch.results.copy())
                got_names.extend(batch.names().tolist())
                blobs = st.emit(ctx, kinds)
                emitted = [a + b for a, b in zip(emitted, blobs)]
        got = np.concatenate(got_parts) if got_parts else np.zeros(0, dtype=_lib.READ_RESULT_DTYPE)
        assert len(got) == n, ("record count", len(got), n, layout)
        assert got_names == [nm.decode() for nm in names], ("names", layout)
        ora, _ = orc.score_reads(go, flat, offsets, olut, thr, n_threads=0)
        sc = ora["n_kmers"] >= 0
        ok = ((got["category"].astype(np.int64) == ora["category"]).all() and (got["n_kmers"][sc] == ora["n_kmers"][sc]).all()
              and (got["n_hits"][sc] == ora["n_hits"][sc]).all())
        if not ok:
            print(json.dumps({"FAIL": "stream vs oracle", "round": rounds, "layout": layout, "alphabet": alphabet, "ksize": ksize, "lanes": n_lanes,
                              "batch_reads": batch_reads}))
            sys.exit(1)
        batch_call = gpu.score_reads(gg, flat, offsets, lut, thr)
        assert got.tobytes() == batch_call.tobytes(), ("stream vs batch call", layout)
        # the stdout stream (kind 0): one record per frame the ORACLE calls coding, with the oracle's translation
        want = []
        for i in range(n):
            frames = None
            for f in range(6):
                if ora["category"][i, f] == orc.CAT_CODING:
                    if frames is None:
                        frames = orc.six_frame_translation(seqs[i])
                    key = _lib.FRAMES[f]
                    jac = repr(int(ora["n_hits"][i, f]) / int(ora["n_kmers"][i, f]))
                    want.append(b">" + names[i].replace(b" ", b"__") + b"__translation_frame:%d__jaccard:%s\n" % (key, jac.encode())
                                + frames[key].encode("latin-1") + b"\n")
        assert emitted[0] == b"".join(want), ("emitted coding peptides", layout, kinds)
        rounds += 1
        reads_total += n
        print(f"round {rounds}: {layout} {alphabet} k={ksize} reads={n} lanes={n_lanes} batch={batch_reads} kinds={kinds} "
              f"coding_frames={len(want)} byte_path={int((got['flags'] & 0x80 != 0).sum())} ok", flush=True)
        del lanes, gg, go
    print(json.dumps({"fuzz_rounds": rounds, "reads": reads_total, "seconds": seconds, "seed": seed, "differences": 0, "mode": "stream"}))


if __name__ == "__main__":
    main()
