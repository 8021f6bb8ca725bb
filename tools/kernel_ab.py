#!/usr/bin/env python
"""A/B harness for the scoring kernels on device-resident reads: variants of the library built with -D switches, one
subprocess each on the same cached inputs (tools/e2e_ab.py's cache), per-kernel CUDA-event times per 10 M-read step and a
digest of the records (every variant must produce the same one).
    ORPH_AB_VARIANTS="base:;exp:-DORPH_X=1,-DORPH_Y=0" python tools/kernel_ab.py build   # here: builds csrc/ab/lib_*.so
    ORPH_AB_VARIANTS=... python tools/kernel_ab.py [alphabet ...]                          # on the GPU box
Round 2 used it for profiles/r02_task_kernel_stage_split_experiment.patch (profiles/r02_kernel_ab_stage_split_carve.log)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CACHE = "/dev/shm/orph_ab"
AB_DIR = os.path.join(ROOT, "orpheum_b200", "csrc", "ab")
VARIANTS = {part.split(":")[0]: [f for f in part.split(":", 1)[1].split(",") if f] for part in os.environ.get("ORPH_AB_VARIANTS", "default:").split(";") if part}


def build():
    from concurrent.futures import ThreadPoolExecutor

    from orpheum_b200 import build as b

    os.makedirs(AB_DIR, exist_ok=True)
    with ThreadPoolExecutor(4) as ex:
        list(ex.map(lambda kv: b.build(force=True, extra_flags=tuple(kv[1]), out=os.path.join(AB_DIR, f"lib_{kv[0]}.so")), VARIANTS.items()))


def run(alphabet):
    import numpy as np
    import torch

    from orpheum_b200 import _lib, gpu, sequence_encodings as enc

    ld = lambda name: np.load(os.path.join(CACHE, name + ".npy"))
    residues, poffs, offsets, packed = ld("residues"), ld("poffs"), ld("offsets"), ld("packed")
    n = len(offsets) - 1
    lut = enc.encode_lut(alphabet)
    ksize = enc.BEST_KSIZES[alphabet]
    stream = torch.cuda.Stream()
    ctx = gpu.Context(0, stream=stream.cuda_stream)
    g = gpu.GpuNodegraph(ksize, int(1e8), 4, ctx=ctx)
    g.add_peptides(residues, poffs, lut, count_unique=False)
    d_reads = torch.from_numpy(packed.view(np.int64)).cuda()
    d_off = torch.from_numpy(offsets.view(np.int64)).cuda()
    d_out = torch.zeros(n * 32, dtype=torch.uint8, device="cuda")
    thr = 0.8 if alphabet == "hp" else 0.5
    with torch.cuda.stream(stream):
        for _ in range(3):
            gpu.score_reads_device(g, d_reads.data_ptr(), d_off.data_ptr(), n, 150, lut, thr, d_out.data_ptr(), ctx=ctx)
        ctx.check()
        ctx.set_kernel_timing(True)
        ctx.kernel_times()
        steps = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            gpu.score_reads_device(g, d_reads.data_ptr(), d_off.data_ptr(), n, 150, lut, thr, d_out.data_ptr(), ctx=ctx)
        e1.record(stream)
        torch.cuda.synchronize()
        kt = ctx.kernel_times()
    import hashlib

    digest = hashlib.sha256(d_out.cpu().numpy().tobytes()).hexdigest()[:12]
    print("AB_RESULT " + json.dumps({"ms_per_step": e0.elapsed_time(e1) / steps, "kernels": {k: round(v[0] / steps, 4) for k, v in kt.items() if v[0]}, "sha": digest}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        build()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "run":
        run(sys.argv[2])
        sys.exit(0)
    if not os.path.exists(os.path.join(CACHE, "packed.npy")):
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "e2e_ab.py")], env={**os.environ, "ORPH_AB_GEN_ONLY": "1"}, check=True)
    alphabets = sys.argv[1:] or ["protein"]
    for alphabet in alphabets:
        for name in VARIANTS:
            env = {**os.environ, "ORPHEUM_B200_LIB": os.path.join(AB_DIR, f"lib_{name}.so")}
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "run", alphabet], env=env, capture_output=True, text=True)
            line = [ln for ln in p.stdout.splitlines() if ln.startswith("AB_RESULT ")]
            print(f"{alphabet:8s} {name:14s} {line[0][10:] if line else 'FAILED: ' + p.stderr[-400:]}", flush=True)
