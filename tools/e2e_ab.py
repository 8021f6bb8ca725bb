#!/usr/bin/env python
"""A/B harness for the host-buffer pipeline of orph_score_reads: one subprocess per environment setting (the library reads
its development knobs once), same cached inputs, median ms per 10 M-read call for the packed and the ASCII route.
    python tools/e2e_ab.py [n_reads]"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CACHE = "/dev/shm/orph_ab"

CONFIGS = [
    ("default", {}),
    ("no_l2_window", {"ORPHEUM_B200_NO_L2_WINDOW": "1"}),
    ("host_threads_4", {"ORPH_AB_HOST_THREADS": "4"}),
]


def gen(n):
    import numpy as np

    from orpheum_b200 import gpu, synth

    os.makedirs(CACHE, exist_ok=True)
    residues, poffs = synth.make_proteome(n_proteins=20000, seed=1001)
    flat, offsets = synth.concat_fixed(synth.make_reads_fixed(n, residues, poffs, length=150, seed=2002))
    packed, _ = gpu.pack_reads(flat, offsets)
    for name, a in (("residues", residues), ("poffs", poffs), ("flat", flat), ("offsets", offsets), ("packed", packed)):
        np.save(os.path.join(CACHE, name + ".npy"), a)


def run(routes):
    import numpy as np
    import torch

    from orpheum_b200 import _lib, gpu, sequence_encodings as enc

    ld = lambda name: np.load(os.path.join(CACHE, name + ".npy"))
    residues, poffs, offsets = ld("residues"), ld("poffs"), ld("offsets")
    n = len(offsets) - 1
    lut = enc.encode_lut("protein")
    ctx = gpu.default_context()
    g = gpu.GpuNodegraph(7, int(1e8), 4, ctx=ctx)
    g.add_peptides(residues, poffs, lut, count_unique=False)

    def pin(a):
        t = torch.empty(a.size, dtype=torch.from_numpy(a[:1]).dtype).pin_memory()
        t.numpy()[:] = a
        return t

    h_off = pin(offsets.view(np.int64))
    h_out = torch.empty(n * 32, dtype=torch.uint8).pin_memory()
    lib = _lib.load()
    if os.environ.get("ORPH_AB_HOST_THREADS"):
        ctx.set_host_threads(int(os.environ["ORPH_AB_HOST_THREADS"]))
    res = {}
    for route in routes:
        buf, fmt = (pin(ld("packed").view(np.int64)), _lib.READS_PACKED2) if route == "packed" else (pin(ld("flat")), _lib.READS_ASCII)
        times = []
        for rep in range(10):
            t0 = time.perf_counter()
            _lib.check(lib.orph_score_reads(ctx._h, g._h, buf.data_ptr(), fmt, h_off.data_ptr(), n, lut.ctypes.data, 0.5, h_out.data_ptr()))
            times.append(1e3 * (time.perf_counter() - t0))
        res[route] = {"median_ms": float(np.median(times[2:])), "min_ms": float(min(times[2:]))}
        del buf
    print("AB_RESULT " + json.dumps(res), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "run":
        run(sys.argv[2].split(","))
        sys.exit(0)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    routes = sys.argv[2] if len(sys.argv) > 2 else "packed,ascii"
    gen(n) if not os.path.exists(os.path.join(CACHE, "packed.npy")) else None
    for name, env in (CONFIGS if not os.environ.get("ORPH_AB_GEN_ONLY") else []):
        e = dict(os.environ)
        e.update(env)
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "run", routes if not env.get("ORPHEUM_B200_AB_SKIP") else "packed"],
                           env=e, capture_output=True, text=True)
        line = [ln for ln in p.stdout.splitlines() if ln.startswith("AB_RESULT ")]
        print(f"{name:24s} {line[0][10:] if line else 'FAILED: ' + p.stderr[-400:]}", flush=True)
