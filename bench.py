#!/usr/bin/env python
"""Headline benchmark: reads/s scored (six-frame translation + Bloom probes) on N B200s.

Workload (BASELINE.json configs[1]): 10 M synthetic 150-nt reads per GPU (half of
them back-translated from the proteome, half uniform ACGT) against a Bloom filter built from a
human-proteome-sized synthetic proteome (20 000 proteins, ~1.07e7 residues), protein alphabet, k=7,
tablesize 1e8 x 4 tables (50 MB).  One step = one pass of the scoring hot path over the whole batch.

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N ...            # the CPU arm (oracle port, all host threads)

Our arm prints one JSON line: `value` = kernel-level reads/s with inputs resident in HBM (CUDA events,
max over ranks); `e2e` = the same metric through the host-buffer C-ABI call (pinned host ASCII reads
in, result records out, H2D and D2H inside the timed region); `roofline` for the dominant kernel;
`cpu_baseline` = the oracle timed on this box's host cores on a bounded sample.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "reads/sec scored (six-frame+Bloom)"
UNIT = "reads/s"
READ_LEN = 150
STREAM_BYTES_PER_READ = (READ_LEN + 3) // 4 + 4 + 32  # 2-bit payload + offset + result record = 74 B


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU per step")
    ap.add_argument("--alphabet", default="protein")
    ap.add_argument("--ksize", type=int, default=None)
    ap.add_argument("--tablesize", type=float, default=None, help="default: 1e8 (L2-resident filter); cfg4: 1e9 (HBM-resident)")
    ap.add_argument("--n-tables", type=int, default=4)
    ap.add_argument("--proteins", type=int, default=20000)
    ap.add_argument("--cpu-sample", type=int, default=400_000, help="reads in the bounded CPU-baseline sample")
    ap.add_argument("--workload", default="r150", choices=["r150", "mixed", "cfg4"],
                    help="r150 = cfg2/cfg3 (fixed 150-nt reads); mixed = cfg5 (50-300 nt, homopolymers, repeats, N, short); "
                         "cfg4 = GPU index build from the 1e8-residue proteome, then --total-reads counter-based reads "
                         "generated in HBM and sharded over the GPUs (strong scaling)")
    ap.add_argument("--total-reads", type=float, default=1e9, help="cfg4: reads of the whole job, split over the GPUs")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-fastq", action="store_true", help="skip the FASTQ-file end-to-end leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the cfg3 / cfg4 / cfg5 legs")
    ap.add_argument("--ref-sample", type=int, default=3000, help="--impl reference: reads per host process per step (unmodified reference)")
    ap.add_argument("--reference-port", action="store_true", help="--impl reference: time the C oracle port even when oracle/_ref is staged")
    return ap.parse_args()


def workload_name(args):
    if args.workload == "mixed":
        return (f"cfg5: {args.reads} synthetic mixed reads/GPU (50-300 nt; 45% coding, 45% random, 4% homopolymer, 3% "
                f"triplet repeats, 2% with N, 1% short) vs {args.proteins}-protein synthetic proteome Bloom filter, "
                f"{args.alphabet} k={args.ksize}, tablesize {int(args.tablesize):.0e} x {args.n_tables}")
    return (f"cfg2: {args.reads} synthetic {READ_LEN}-nt reads/GPU (50% back-translated, 1% subst.) vs "
            f"{args.proteins}-protein synthetic proteome Bloom filter, {args.alphabet} k={args.ksize}, "
            f"tablesize {int(args.tablesize):.0e} x {args.n_tables}")


def default_ksize(alphabet):
    from orpheum_b200.sequence_encodings import BEST_KSIZES

    return BEST_KSIZES[alphabet]


def default_threshold(alphabet):
    return 0.8 if alphabet in ("hp", "hydrophobic-polar") else 0.5  # translate.py:31-37


class _LimitedReader:
    """File-like view of the first `limit` bytes up to the last complete four-line record in them."""

    def __init__(self, fh, limit):
        data = fh.read(limit)
        lines = data.split(b"\n")
        keep = (len(lines) - 1) // 4 * 4
        self._buf = b"\n".join(lines[:keep]) + (b"\n" if keep else b"")
        self._at = 0

    def read(self, n=-1):
        n = len(self._buf) - self._at if n is None or n < 0 else n
        out = self._buf[self._at:self._at + n]
        self._at += len(out)
        return out


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU while the timed region runs (NVML)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            if visible:
                ids = [v for v in visible.split(",") if v.strip() != ""]
                try:
                    index = int(ids[index])
                except (ValueError, IndexError):
                    pass
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _run(self):
        nv = self._nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for name, bit in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.005)

    def __enter__(self):
        if self._nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        return {
            "sm_mhz": float(np.median(self.samples)) if self.samples else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
        }


def n_host_threads():
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    return max(1, cores // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1"))))


def host_memory_read_rate(threads, mb=128, reps=3):
    """GB/s at which `threads` threads of this process sum private arrays (numpy releases the GIL): the host-memory read rate
    the packing lane and the DMA engines share."""
    import threading

    arrays = [np.ones(mb << 20, dtype=np.uint8).view(np.uint64) for _ in range(threads)]
    gate = threading.Barrier(threads + 1)

    def work(a):
        gate.wait()
        for _ in range(reps):
            a.sum()
        gate.wait()

    pool = [threading.Thread(target=work, args=(a,)) for a in arrays]
    for t in pool:
        t.start()
    gate.wait()
    t0 = time.perf_counter()
    gate.wait()
    dt = time.perf_counter() - t0
    for t in pool:
        t.join()
    return {"threads": threads, "read_GBps": threads * reps * (mb << 20) / dt / 1e9}


def make_inputs(args, rank):
    from orpheum_b200 import synth

    residues, poffs = synth.make_proteome(n_proteins=args.proteins, seed=1001)
    if args.workload == "mixed":
        flat, offsets = synth.make_reads_mixed(args.reads, residues, poffs, ksize=args.ksize, seed=5005 + rank)
    else:
        reads2d = synth.make_reads_fixed(args.reads, residues, poffs, length=READ_LEN, seed=2002 + rank)
        flat, offsets = synth.concat_fixed(reads2d)
    return residues, poffs, flat, offsets


def oracle_filter(args, residues, poffs):
    from oracle import oracle as orc

    g = orc.Nodegraph(args.ksize, int(args.tablesize), args.n_tables)
    g.add_peptides(residues, poffs, orc.encode_lut(args.alphabet))
    return g


def time_oracle(args, g, flat, offsets, n_sample, threads):
    from oracle import oracle as orc

    n_sample = min(n_sample, len(offsets) - 1)
    sub_off = offsets[: n_sample + 1]
    t0 = time.perf_counter()
    out, _ = orc.score_reads(g, flat, sub_off, orc.encode_lut(args.alphabet), default_threshold(args.alphabet),
                             n_threads=threads)
    dt = time.perf_counter() - t0
    return n_sample / dt, out


def _reference_worker(conn, ng_path, fq_path, csv_path, alphabet, ksize):
    """One host process of the reference arm: the UNMODIFIED reference CLI (staged by oracle/stage_reference.py, khmer and
    sourmash as native calls) run in-process on this worker's FASTQ file, once per 'go' message."""
    import contextlib

    sys.path.insert(0, ROOT)
    from oracle import stage_reference

    stage_reference.activate()
    import orpheum.translate as ref_translate  # the staged, unmodified module

    argv = ["--peptides-are-bloom-filter", "--alphabet", alphabet, "--peptide-ksize", str(ksize), "--csv", csv_path, ng_path, fq_path]
    conn.send("ready")
    with open(os.devnull, "w") as devnull:
        while conn.recv() == "go":
            with contextlib.redirect_stdout(devnull):
                ref_translate.cli.main(argv, standalone_mode=False)
            conn.send("done")


def time_reference_unmodified(args, cores, per_worker, steps, warmup):
    """orpheum's own `translate` CLI (orpheum/translate.py:385-602, unmodified, staged in oracle/_ref by build()) on `cores`
    host cores -- the reference is single-threaded, so one process per core, each on its own FASTQ slice of the same
    synthetic workload against the same Bloom filter (.nodegraph written by the oracle).  -> (reads/s, seconds, CSV of
    worker 0 identical to the oracle's scores)."""
    import multiprocessing as mp
    import shutil
    import tempfile

    from oracle import oracle as orc

    total = per_worker * cores
    residues, poffs, flat, offsets = make_inputs(argparse.Namespace(**{**vars(args), "reads": total}), 0)
    g = oracle_filter(args, residues, poffs)
    tmp = tempfile.mkdtemp(prefix="orph_ref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        ng = os.path.join(tmp, "filter.nodegraph")
        g.save(ng)
        qual = b"I" * 400
        ctx = mp.get_context("spawn")
        workers = []
        for w in range(cores):
            fq = os.path.join(tmp, f"reads_{w}.fq")
            with open(fq, "wb") as fh:
                for i in range(w * per_worker, (w + 1) * per_worker):
                    seq = flat[int(offsets[i]):int(offsets[i + 1])].tobytes()
                    fh.write(b"@read%d 1:N:0\n%s\n+\n%s\n" % (i, seq, qual[: len(seq)]))
            parent, child = ctx.Pipe()
            proc = ctx.Process(target=_reference_worker, args=(child, ng, fq, os.path.join(tmp, f"scores_{w}.csv"), args.alphabet, args.ksize), daemon=True)
            proc.start()
            workers.append((proc, parent))
        for _, conn in workers:
            assert conn.recv() == "ready"

        def step():
            for _, conn in workers:
                conn.send("go")
            for _, conn in workers:
                assert conn.recv() == "done"

        for _ in range(warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        dt = time.perf_counter() - t0
        for proc, conn in workers:
            conn.send("stop")
            proc.join(timeout=10)
        # the CSV of worker 0 against the oracle on the same reads: the timed thing computes what the oracle computes
        import pandas as pd

        df = pd.read_csv(os.path.join(tmp, "scores_0.csv"), float_precision="round_trip")
        ora, _ = orc.score_reads(g, flat, offsets[: per_worker + 1], orc.encode_lut(args.alphabet), default_threshold(args.alphabet), n_threads=1)
        nk = df["n_kmers"].to_numpy().reshape(per_worker, 6)
        jac = df["jaccard_in_peptide_db"].to_numpy().reshape(per_worker, 6)
        scored = ora["n_kmers"] >= 0
        with np.errstate(invalid="ignore", divide="ignore"):
            want_j = ora["n_hits"] / ora["n_kmers"]
        cat = ora["category"]
        has_j = (cat == 3) | (cat == 4)
        agrees = bool(np.array_equal(np.isnan(nk), ~scored) and (nk[scored] == ora["n_kmers"][scored]).all()
                      and np.array_equal(np.isnan(jac), ~has_j) and (jac[has_j] == want_j[has_j]).all())
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return steps * total / dt, dt, agrees


def run_reference_unmodified(args):
    """CPU arm, kind "reference": the unmodified reference on every host core (time_reference_unmodified)."""
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    per_worker = args.ref_sample
    total = per_worker * cores
    value, dt, agrees = time_reference_unmodified(args, cores, per_worker, args.steps, args.warmup)
    sample = (f"{per_worker} reads per process per step x {cores} single-threaded processes (the same synthetic workload), unmodified "
              f"orpheum/translate.py CLI in-process with --csv, khmer/sourmash calls native (oracle/refnative)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(args), "reads_per_step_measured": total,
                   "note": "same generator, mix and filter as the GPU arm's config; the CPU arm scores a bounded sample per step (rates are comparable, step sizes are not)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample,
                         "per_core": value / cores, "csv_matches_oracle": agrees},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_reference(args):
    """CPU arm.  With oracle/_ref staged (build() does it wherever /root/reference exists) the unmodified reference is
    timed (kind "reference"); otherwise the oracle port (the reference's algorithm restated in C) with every host thread,
    bounded sample per step (kind "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc, stage_reference

    if stage_reference.staged() and not args.reference_port and args.workload == "r150":
        return run_reference_unmodified(args)
    residues, poffs, flat, offsets = make_inputs(argparse.Namespace(**{**vars(args), "reads": args.cpu_sample}), 0)
    g = oracle_filter(args, residues, poffs)
    threads = orc.max_threads()
    for _ in range(args.warmup):
        time_oracle(args, g, flat, offsets, args.cpu_sample // 4, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        time_oracle(args, g, flat, offsets, args.cpu_sample, threads)
    dt = time.perf_counter() - t0
    value = args.steps * args.cpu_sample / dt
    sample = f"{args.cpu_sample} reads per step (prefix of the same synthetic workload), {threads} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(args), "reads_per_step_measured": args.cpu_sample,
                   "note": "same generator, mix and filter as the GPU arm's config; the CPU arm scores a bounded sample per step (rates are comparable, step sizes are not)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def check_prefix_against_oracle(res, ora):
    scored = ora["n_kmers"] >= 0
    return bool((res["category"].astype(np.int64) == ora["category"]).all() and (res["n_kmers"][scored] == ora["n_kmers"][scored]).all()
                and (res["n_hits"][scored] == ora["n_hits"][scored]).all())


def config_leg(name, ctx, stream, dev, alphabet, workload, n_reads, tablesize, n_tables, proteins, steps, warmup, n_check=50_000):
    """One of BASELINE.json's other configs on ONE GPU, reads resident in HBM: the same kernel-level measurement as the
    headline (CUDA events on the launch stream around `steps` passes), a prefix checked against the oracle, and the
    random-gather ceiling measured for THIS filter's footprint."""
    import torch

    from oracle import oracle as orc
    from orpheum_b200 import _lib, gpu, sequence_encodings as enc, synth

    ksize = default_ksize(alphabet)
    thr = default_threshold(alphabet)
    lut, olut = enc.encode_lut(alphabet), orc.encode_lut(alphabet)
    residues, poffs = synth.make_proteome(n_proteins=proteins, seed=1001)
    if workload == "mixed":
        flat, offsets = synth.make_reads_mixed(n_reads, residues, poffs, ksize=ksize, seed=5005)
    else:
        flat, offsets = synth.concat_fixed(synth.make_reads_fixed(n_reads, residues, poffs, length=READ_LEN, seed=2002))
    graph = gpu.GpuNodegraph(ksize, int(tablesize), n_tables, ctx=ctx)
    graph.add_peptides(residues, poffs, lut, count_unique=False)
    packed, needs = gpu.pack_reads(flat, offsets)
    d_offsets = torch.from_numpy(offsets.view(np.int64)).to(dev)
    d_out = torch.zeros(n_reads * 32, dtype=torch.uint8, device=dev)
    max_len = int(np.diff(offsets.astype(np.int64)).max())
    n_plane = None
    if needs.any():
        # reads with N: the device-resident form is the 2-bit stream + N plane, built on the device from the ASCII reads before
        # the clock starts (what a caller with N reads keeps in HBM); bytes other than ACGTN would force the ASCII form
        d_ascii = torch.from_numpy(flat).to(dev)
        total = int(offsets[-1])
        d_reads = torch.zeros((total + 31) // 32 + 2, dtype=torch.int64, device=dev)
        n_plane = (torch.zeros((total + 31) // 32 + 2, dtype=torch.int32, device=dev), torch.zeros(n_reads, dtype=torch.uint8, device=dev))
        d_needs = torch.zeros(n_reads, dtype=torch.uint8, device=dev)
        with torch.cuda.stream(stream):
            gpu.pack_reads_device(ctx, d_ascii.data_ptr(), d_offsets.data_ptr(), n_reads, total, d_reads.data_ptr(), n_plane[0].data_ptr(),
                                  n_plane[1].data_ptr(), d_needs.data_ptr())
        torch.cuda.synchronize()
        fmt = "packed2 + N plane"
        if bool(d_needs.any().item()):
            d_reads, fmt, n_plane = d_ascii, _lib.READS_ASCII, None
        else:
            del d_ascii
    else:
        d_reads, fmt = torch.from_numpy(packed.view(np.int64)).to(dev), _lib.READS_PACKED2

    def step():
        if n_plane is not None:
            gpu.score_reads_device_n(graph, d_reads.data_ptr(), n_plane[0].data_ptr(), n_plane[1].data_ptr(), d_offsets.data_ptr(), n_reads, max_len, lut, thr,
                                     d_out.data_ptr(), ctx=ctx)
        else:
            gpu.score_reads_device(graph, d_reads.data_ptr(), d_offsets.data_ptr(), n_reads, max_len, lut, thr, d_out.data_ptr(), reads_format=fmt, ctx=ctx)

    with torch.cuda.stream(stream):
        for _ in range(warmup):
            step()
        ctx.check()
        torch.cuda.synchronize()
        ctx.set_kernel_timing(True)
        ctx.kernel_times()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        ktimes = ctx.kernel_times()
        ctx.set_kernel_timing(False)
        ctx.check()
    res = d_out.cpu().numpy().view(_lib.READ_RESULT_DTYPE)
    n_check = min(n_check, n_reads)
    g_cpu = orc.Nodegraph(ksize, int(tablesize), n_tables)
    g_cpu.add_peptides(residues, poffs, olut)
    ora, _ = orc.score_reads(g_cpu, flat, offsets[: n_check + 1], olut, thr, n_threads=orc.max_threads())
    parity = check_prefix_against_oracle(res[:n_check], ora)
    if not parity:
        raise SystemExit(f"bench: leg {name}: GPU results differ from the oracle on the checked prefix")
    pmin = float(ora["probes_min"].sum()) / n_check
    _, fbytes, _ = graph.device_buffer()
    peak = max(ctx.gather_peak(fbytes, 512, False), ctx.gather_peak(fbytes, 512, True))
    dominant = max(ktimes, key=lambda k: ktimes[k][0])
    dom_ms = ktimes[dominant][0] / steps  # all launches of the kernel in one step (mixed lengths: one per length tier)
    # the kernels that probe the filter (thread-per-frame, warp-per-read, byte-exact, block-per-read), per step
    probing_ms = sum(ktimes[k][0] for k in ("score_tasks", "score_frames", "score_bytes", "score_long") if k in ktimes) / steps
    probe_rate = n_reads * pmin / (probing_ms * 1e-3)
    out = {"workload": f"{n_reads} {'mixed 50-300 nt (cfg5 recipe)' if workload == 'mixed' else '150-nt'} reads vs {proteins}-protein proteome, {alphabet} k={ksize}, "
                       f"tablesize {int(tablesize):.0e} x {n_tables} ({fbytes / 1e6:.0f} MB filter)",
           "value": n_reads / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "reads": n_reads,
           "input_format": fmt if isinstance(fmt, str) else "ascii (bytes other than ACGTN)" if fmt == _lib.READS_ASCII else "packed2",
           "dominant_kernel": dominant + "_kernel", "dominant_kernel_ms_per_step": dom_ms, "probing_kernels_ms_per_step": probing_ms,
           "kernel_ms_per_step": {k: v[0] / steps for k, v in ktimes.items() if v[0] > 0},
           "probes_min_per_read": pmin, "probe_sectors_per_s": probe_rate, "gather_peak_sectors_per_s": peak,
           "gather_frac": probe_rate / peak, "byte_path_reads": int((res["flags"] & _lib.FLAG_BYTE_PATH != 0).sum()),
           "parity_checked_vs_oracle": parity, "parity_reads_checked": n_check}
    del graph, d_reads, d_offsets, d_out
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    from orpheum_b200 import _lib, build, gpu, parallel, sequence_encodings as enc

    build.build()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    stream = torch.cuda.Stream(device=dev)
    ctx = gpu.Context(local, stream=stream.cuda_stream)
    lut = enc.encode_lut(args.alphabet)
    thr = default_threshold(args.alphabet)

    t_gen = time.perf_counter()
    residues, poffs, flat, offsets = make_inputs(args, rank)
    t_gen = time.perf_counter() - t_gen

    # ---- filter: built on rank 0's GPU, replicated to the other GPUs over NVLink (NCCL broadcast) ----
    graph = gpu.GpuNodegraph(args.ksize, int(args.tablesize), args.n_tables, ctx=ctx)
    t_build = time.perf_counter()
    if rank == 0:
        graph.add_peptides(residues, poffs, lut)
    ctx.synchronize()
    t_build = time.perf_counter() - t_build
    if world > 1:
        parallel.broadcast_filter(graph, src=0)
    fill = graph.n_occupied(0) / graph.hashsizes()[0]

    # ---- device-resident inputs (kernel-level number) ----
    packed, needs = gpu.pack_reads(flat, offsets)
    d_offsets = torch.from_numpy(offsets.view(np.int64)).to(dev)
    d_out = torch.zeros(args.reads * 32, dtype=torch.uint8, device=dev)
    max_len = int(np.diff(offsets.astype(np.int64)).max())
    if needs.any():
        # reads with N / lower case cannot be 2-bit packed: the device-resident input is the ASCII stream
        d_reads, fmt = torch.from_numpy(flat).to(dev), _lib.READS_ASCII
    else:
        d_reads, fmt = torch.from_numpy(packed.view(np.int64)).to(dev), _lib.READS_PACKED2
    torch.cuda.synchronize()

    def step_device():
        gpu.score_reads_device(graph, d_reads.data_ptr(), d_offsets.data_ptr(), args.reads, max_len, lut, thr,
                               d_out.data_ptr(), reads_format=fmt, ctx=ctx)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step_device()
        ctx.check()
        barrier()
        ctx.set_kernel_timing(True)
        ctx.kernel_times()
        launches0 = ctx.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as clocks:
            barrier()
            e0.record(stream)
            for _ in range(args.steps):
                step_device()
            e1.record(stream)
            barrier()
        ms_total = e0.elapsed_time(e1)
        launches = ctx.launch_count - launches0
        ktimes = ctx.kernel_times()
        ctx.set_kernel_timing(False)
        ctx.check()
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = world * args.reads / (ms_per_step * 1e-3)

    # ---- parity spot check against the oracle (not timed) ----
    res_dev = d_out.cpu().numpy().view(_lib.READ_RESULT_DTYPE)
    parity = None
    pmin_mean = None
    cpu = None
    if rank != 0:
        # every rank checks a prefix of its own shard against the oracle
        from oracle import oracle as orc

        n_chk = min(20_000, args.reads)
        _, ora = time_oracle(args, oracle_filter(args, residues, poffs), flat, offsets, n_chk, 2)
        sub, scored = res_dev[:n_chk], ora["n_kmers"] >= 0
        ok = bool((sub["category"].astype(np.int64) == ora["category"]).all()
                  and (sub["n_kmers"][scored] == ora["n_kmers"][scored]).all()
                  and (sub["n_hits"][scored] == ora["n_hits"][scored]).all())
        if not ok:
            raise SystemExit(f"bench: rank {rank}: GPU results differ from the oracle on the checked prefix")
    if rank == 0:
        from oracle import oracle as orc

        g_cpu = oracle_filter(args, residues, poffs)
        n_chk = min(args.cpu_sample, args.reads)
        if not args.no_cpu_baseline and world == 1:
            threads = orc.max_threads()
            rate, ora = time_oracle(args, g_cpu, flat, offsets, n_chk, threads)
            cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"first {n_chk} reads of the same workload, C oracle (oracle/orpheum_oracle.c), {threads} threads"}
            from oracle import stage_reference

            if stage_reference.staged() and args.workload == "r150":
                # BASELINE.md 2's CPU-1 figure: the unmodified reference CLI on ONE core (the all-core figure is the
                # `--impl reference` arm's line)
                try:
                    v1, _, ok1 = time_reference_unmodified(args, 1, 3000, 2, 1)
                    cpu["unmodified_reference_one_core"] = {"value": v1, "unit": UNIT, "cores": 1, "kind": "reference", "csv_matches_oracle": ok1,
                                                            "sample": "3000 reads per step, 2 steps, orpheum/translate.py CLI in-process (oracle/_ref)"}
                except Exception as exc:
                    cpu["unmodified_reference_one_core"] = {"error": f"{type(exc).__name__}: {exc}"}
        else:
            n_chk = min(50_000, args.reads)
            _, ora = time_oracle(args, g_cpu, flat, offsets, n_chk, orc.max_threads())
        sub = res_dev[:n_chk]
        scored = ora["n_kmers"] >= 0
        parity = bool((sub["category"].astype(np.int64) == ora["category"]).all()
                      and (sub["n_kmers"][scored] == ora["n_kmers"][scored]).all()
                      and (sub["n_hits"][scored] == ora["n_hits"][scored]).all())
        pmin_mean = float(ora["probes_min"].sum()) / n_chk
        pmax_mean = float(ora["probes_max"].sum()) / n_chk
        if not parity:
            raise SystemExit("bench: GPU results differ from the oracle on the checked prefix")

    # ---- end to end through the host-buffer C ABI (pinned host ASCII in, records out) ----
    e2e = None
    host_info = None
    if not args.no_e2e:
        # what the host can do, for reading the e2e figures: cores of this process and the rate at which they read memory
        # (every rank's threads run at once, as they do in the legs below)
        host_info = host_memory_read_rate(n_host_threads())
        if world > 1:
            t = torch.tensor([host_info["read_GBps"]], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            host_info["read_GBps_all_ranks"] = float(t.item())
        h_reads = torch.empty(flat.size, dtype=torch.uint8).pin_memory()
        h_reads.numpy()[:] = flat
        h_offsets = torch.empty(offsets.size, dtype=torch.int64).pin_memory()
        h_offsets.numpy()[:] = offsets.view(np.int64)
        h_out = torch.empty(args.reads * 32, dtype=torch.uint8).pin_memory()
        lib = _lib.load()

        def step_host():
            rc = lib.orph_score_reads(ctx._h, graph._h, h_reads.data_ptr(), _lib.READS_ASCII, h_offsets.data_ptr(),
                                      args.reads, lut.ctypes.data, thr, h_out.data_ptr())
            _lib.check(rc)

        for _ in range(max(1, min(args.warmup, 2))):
            step_host()
        barrier()
        stats0 = ctx.ingest_stats()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_host()
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        same = h_out.numpy().view(_lib.READ_RESULT_DTYPE).tobytes() == res_dev.tobytes()
        stats1 = ctx.ingest_stats()
        n_chunks = max(1, stats1["chunks_packed_on_host"] + stats1["chunks_sent_ascii"] - stats0["chunks_packed_on_host"] - stats0["chunks_sent_ascii"])
        frac_packed = (stats1["chunks_packed_on_host"] - stats0["chunks_packed_on_host"]) / n_chunks
        # bytes that actually crossed PCIe: chunks the host packed went out at 2 bits per base
        h2d = int(flat.nbytes * (1 - frac_packed) + packed.nbytes * frac_packed + offsets.nbytes)
        e2e = {"value": world * args.reads * args.steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(args.reads * 32),
               "input": "pinned host ASCII reads + u64 offsets via orph_score_reads(); the library 2-bit packs part of the "
                        "chunks on the host threads while the copy engine ships the others as ASCII",
               "ms_per_step": 1e3 * dt / args.steps, "identical_to_device_path": bool(same),
               "host": host_info,
               "ingest": {"host_threads": stats1["host_threads"], "fraction_of_chunks_packed_on_host": frac_packed,
                          "host_pack_GBps": stats1["pack_bytes_per_s"] / 1e9, "h2d_GBps_measured": stats1["h2d_bytes_per_s"] / 1e9}}
        # the same call with host packing disabled: every byte of ASCII crosses PCIe
        ctx.set_host_threads(0)
        step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_host()
        barrier()
        dta = time.perf_counter() - t0
        ctx.set_host_threads(-1)
        if world > 1:
            t = torch.tensor([dta], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dta = float(t.item())
        e2e["ascii_over_pcie_only"] = {"value": world * args.reads * args.steps / dta, "ms_per_step": 1e3 * dta / args.steps,
                                       "h2d_bytes_per_step": int(flat.nbytes + offsets.nbytes),
                                       "identical_to_device_path": bool(h_out.numpy().view(_lib.READ_RESULT_DTYPE).tobytes() == res_dev.tobytes())}
        if not needs.any():
            # the same call with the host buffer already 2-bit packed (what a packing FASTQ parser would hand over);
            # reported next to the ASCII figure, which stays the headline e2e number
            h_packed = torch.empty(packed.size, dtype=torch.int64).pin_memory()
            h_packed.numpy()[:] = packed.view(np.int64)

            def step_host_packed():
                _lib.check(lib.orph_score_reads(ctx._h, graph._h, h_packed.data_ptr(), _lib.READS_PACKED2,
                                                h_offsets.data_ptr(), args.reads, lut.ctypes.data, thr, h_out.data_ptr()))

            step_host_packed()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                step_host_packed()
            barrier()
            dtp = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dtp], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dtp = float(t.item())
            e2e["packed_input"] = {"value": world * args.reads * args.steps / dtp, "ms_per_step": 1e3 * dtp / args.steps,
                                   "h2d_bytes_per_step": int(packed.nbytes + offsets.nbytes),
                                   "identical_to_device_path": bool(h_out.numpy().view(_lib.READ_RESULT_DTYPE).tobytes() == res_dev.tobytes())}

    # ---- from a FASTQ file: orph_stream_* (parallel parse + 2-bit pack into pinned slots -> H2D -> kernels -> records) ----
    # every rank streams its own file (its shard of the reads) at the same time; value = all reads / slowest rank
    if e2e is not None and args.workload == "r150" and not args.no_fastq:
        import shutil
        import tempfile

        n_fq = min(args.reads, 4_000_000 if world == 1 else 2_000_000)
        need_bytes = n_fq * (2 * READ_LEN + 40) * (world if rank == 0 else 1)
        shm_ok = os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > 2 * need_bytes * max(1, world)
        tmpdir = tempfile.mkdtemp(prefix="orph_bench_", dir="/dev/shm" if shm_ok else None)
        fq_leg = None
        fq_out = np.zeros(n_fq, dtype=_lib.READ_RESULT_DTYPE)  # the caller's result array (pages touched before the clock starts)
        problems = []

        def pass_fastq(path):
            # never raises: every rank must reach the barriers around the timed passes
            got = 0
            try:
                with gpu.ScoringStream(graph, path, lut, thr, ctx=ctx) as stream_:
                    for batch in stream_:
                        got = batch.copy_results_into(fq_out, got)
            except Exception as exc:
                problems.append(f"{type(exc).__name__}: {exc}")
                return -1
            return got

        try:
            fq = os.path.join(tmpdir, "reads.fq")
            fq_bytes = 0
            try:
                qual = b"I" * READ_LEN
                rows = flat[: n_fq * READ_LEN].reshape(n_fq, READ_LEN)
                with open(fq, "wb") as fh:
                    for lo in range(0, n_fq, 200_000):
                        fh.write(b"".join(b"@read%d 1:N:0\n%s\n+\n%s\n" % (i, rows[i].tobytes(), qual) for i in range(lo, min(n_fq, lo + 200_000))))
                fq_bytes = os.path.getsize(fq)
            except Exception as exc:
                problems.append(f"{type(exc).__name__}: {exc}")
            pass_fastq(fq)
            reps = 3
            barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                got = pass_fastq(fq)
            barrier()
            dtf = (time.perf_counter() - t0) / reps
            same_fq = got == n_fq and fq_out.tobytes() == res_dev[:n_fq].tobytes()
            if world > 1:
                t = torch.tensor([dtf, 0.0 if same_fq else 1.0, float(len(problems))], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dtf, same_fq = float(t[0].item()), bool(t[1].item() == 0.0)
                if t[2].item() > 0 and not problems:
                    problems.append("another rank failed")
            if problems:
                raise RuntimeError("; ".join(problems))
            fq_leg = {"value": world * n_fq / dtf, "unit": UNIT, "reads_per_gpu": n_fq, "fastq_bytes_per_gpu": fq_bytes, "ms": 1e3 * dtf,
                      "fastq_GBps": world * fq_bytes / dtf / 1e9, "identical_to_device_path": bool(same_fq),
                      "path": "uncompressed four-line FASTQ on tmpfs (one file per rank) -> orph_stream_* (mapped text parsed and 2-bit packed "
                              "into pinned slots in one parallel pass, next batch parsed while this one is copied and scored) -> records "
                              "copied into the caller's array; includes opening and closing the stream"}
            if world == 1:
                # the same reads as gzip (screed reads .gz transparently, translate.py:370): one member, zlib level 1
                import gzip

                n_gz = min(n_fq, 1_000_000)
                gz_path = os.path.join(tmpdir, "reads.fq.gz")
                with open(fq, "rb") as fh, gzip.open(gz_path, "wb", compresslevel=1) as gzf:
                    shutil.copyfileobj(_LimitedReader(fh, n_gz * (fq_bytes // n_fq + 2)), gzf, 1 << 24)
                got = pass_fastq(gz_path)
                t0 = time.perf_counter()
                got = pass_fastq(gz_path)
                dtg = time.perf_counter() - t0
                fq_leg["gzip"] = {"value": got / dtg, "unit": UNIT, "reads": got, "gz_bytes": os.path.getsize(gz_path), "ms": 1e3 * dtg,
                                  "layout": "one gzip member (zlib level 1): a single deflate stream, inflated sequentially",
                                  "identical_to_device_path": bool(fq_out[:got].tobytes() == res_dev[:got].tobytes())}
                # the same text as blocked gzip (BGZF: bgzip / samtools / sequencing pipelines): members inflated in parallel
                from orpheum_b200 import synth

                bg_path = os.path.join(tmpdir, "reads.bgzf.fq.gz")
                with gzip.open(gz_path, "rb") as gzf:
                    synth.write_bgzf(bg_path, gzf.read())
                got = pass_fastq(bg_path)
                t0 = time.perf_counter()
                got = pass_fastq(bg_path)
                dtb = time.perf_counter() - t0
                fq_leg["bgzf"] = {"value": got / dtb, "unit": UNIT, "reads": got, "gz_bytes": os.path.getsize(bg_path), "ms": 1e3 * dtb,
                                  "layout": "BGZF (64 KiB members, zlib level 1), members inflated in parallel on the host pool",
                                  "identical_to_device_path": bool(fq_out[:got].tobytes() == res_dev[:got].tobytes())}
        except Exception as exc:  # a measurement extra: never lose the bench line over it (disk full, ...)
            fq_leg = {"error": f"{type(exc).__name__}: {exc}"}
        finally:
            shutil.rmtree(tmpdir, ignore_errors=True)
        e2e["from_fastq"] = fq_leg

    # ---- cfg5 output parity over the ranks (BASELINE configs[4]): records gathered in read order, CSV / parquet / JSON ----
    output_parity = None
    if world > 1:
        import importlib.util

        spec = importlib.util.spec_from_file_location("orph_mgpu_parity", os.path.join(ROOT, "tests", "scripts", "multi_gpu_output_parity.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        try:
            output_parity = mod.check(50_000, rank, world, local, ctx)
        except Exception as exc:
            output_parity = {"error": f"{type(exc).__name__}: {exc}"}

    # ---- measured random-gather ceiling for this filter footprint ----
    gather = None
    if rank == 0:
        _, fbytes, _ = graph.device_buffer()
        try:
            gather = {"footprint_bytes": fbytes, "sectors_per_s_plain": ctx.gather_peak(fbytes, 512, False),
                      "sectors_per_s_l2_persist": ctx.gather_peak(fbytes, 512, True)}
        except Exception as exc:  # measurement helper only
            gather = {"error": str(exc)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak_gbs, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak_gbs, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    dominant = max(ktimes, key=lambda name: ktimes[name][0])
    sf_ms, sf_n = ktimes[dominant]
    sf_ms_per_launch = sf_ms / args.steps  # the kernel's launches of one step (one, unless the batch has two length tiers)
    # The dominant kernel is bound by the random 32-byte-sector gather rate of the Bloom probes (L1TEX tag stage / L2), not by
    # the HBM stream: SURVEY 8(d)'s per-unit figure is P_min probes per read (counted exactly by the oracle) x 32 B, the roof is
    # the gather ceiling measured on this GPU for this filter's footprint (orph_bench_gather_peak), both in GB/s of sectors.
    stream_bytes = ((packed.nbytes if fmt == _lib.READS_PACKED2 else flat.nbytes) / args.reads) + 4 + 32
    probe_rate = args.reads * pmin_mean / (sf_ms_per_launch * 1e-3)
    roofline = {"kernel": dominant + "_kernel", "ms_per_launch": sf_ms_per_launch, "share_of_step": sf_ms_per_launch / ms_per_step,
                "probes_min_per_read": pmin_mean, "probes_max_per_read": pmax_mean, "probe_sectors_per_s": probe_rate,
                "kernel_ms_per_step": {k: v[0] / args.steps for k, v in ktimes.items()}}
    if gather and "sectors_per_s_plain" in gather:
        best = max(gather["sectors_per_s_plain"], gather["sectors_per_s_l2_persist"])
        roofline.update({"bound": "l2-gather", "achieved": probe_rate * 32 / 1e9, "peak": best * 32 / 1e9, "unit": "GB/s",
                         "frac": probe_rate / best, "gather_frac": probe_rate / best, "gather_peak_sectors_per_s": best,
                         "peak_source": f"measured in this run: orph_bench_gather_peak over the filter's {gather['footprint_bytes']} bytes "
                                        "(random 32-byte sector loads, 512 per thread); not in MEASURED_PEAKS.json, cross-check: "
                                        "148 SMs x 1 sector/clk x SM clock",
                         "algorithmic_bytes_per_read": 32.0 * pmin_mean,
                         "note": "achieved = P_min required probes per read x 32-byte sectors / dominant kernel time"})
    else:
        roofline.update({"bound": "l2-gather", "achieved": probe_rate * 32 / 1e9, "peak": None, "unit": "GB/s", "frac": None})
    # secondary: the same kernel time against the HBM roof (74 B/read of stream + 32 B per required probe): exceeds 1 when the
    # filter is L2-resident, because the probes never reach HBM
    algo_bytes = args.reads * (stream_bytes + 32.0 * pmin_mean)
    roofline["hbm_stream"] = {"bound": "hbm", "achieved": algo_bytes / (sf_ms_per_launch * 1e-3) / 1e9, "peak": peak_gbs, "unit": "GB/s",
                              "frac": algo_bytes / (sf_ms_per_launch * 1e-3) / 1e9 / peak_gbs, "peak_source": peak_src,
                              "stream_bytes_per_read": stream_bytes}
    # DRAM bytes of the dominant kernel per launch are NOT measured by this run (that needs ncu): the figure below is the
    # committed `ncu --set full` capture of the same kernel and workload on 2 M reads, scaled to this launch's reads
    if args.workload == "r150" and dominant == "score_tasks" and args.alphabet == "protein":
        roofline["traffic"] = 115.4 * args.reads
        roofline["traffic_source"] = ("ncu --set full, profiles/r02b_ncu_full_tasks.csv: dram__bytes_read.sum 187.3 MB + "
                                      "dram__bytes_write.sum 43.5 MB per 2 M-read launch = 115.4 B/read, scaled by reads")
        roofline["l2_sectors_per_read_ncu"] = 115.1
        roofline["l2_sectors_source"] = ("same capture: lts__t_sectors 2.302e8 per 2 M reads; l1tex__m_l1tex2xbar_req_cycles_active 90.2 % "
                                         "(the binding unit); all sectors / gather peak = 0.94")
    else:
        roofline["traffic"] = None
    configs = None
    if not args.no_configs and args.workload == "r150" and args.alphabet == "protein":
        # the other BASELINE.json configs, on this rank's GPU (the other ranks have left): kernel-level, parity-checked
        configs = {}
        for name, kw in (("cfg3_dayhoff", dict(alphabet="dayhoff", workload="r150", n_reads=2_000_000, tablesize=1e8)),
                         ("cfg3_hp", dict(alphabet="hp", workload="r150", n_reads=2_000_000, tablesize=1e8)),
                         ("cfg5_mixed", dict(alphabet="protein", workload="mixed", n_reads=2_000_000, tablesize=1e8)),
                         ("cfg4_hbm_filter", dict(alphabet="protein", workload="r150", n_reads=2_000_000, tablesize=1e9))):
            try:
                configs[name] = config_leg(name, ctx, stream, dev, n_tables=args.n_tables, proteins=args.proteins, steps=5, warmup=2, **kw)
            except SystemExit:
                raise
            except Exception as exc:
                configs[name] = {"error": f"{type(exc).__name__}: {exc}"}
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic",
        "config": {"workload": workload_name(args), "reads_per_gpu": args.reads, "read_len": READ_LEN,
                   "filter_fill_table0": fill, "l2_flush": "inputs (375 MB packed + 80 MB offsets + 320 MB results per step) exceed the 126 MB L2",
                   "parallelism": f"reads sharded over {world} GPU(s), filter replicated (NCCL broadcast), no steady-state collectives"},
        "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        "gather_peak": gather, "parity_checked_vs_oracle": parity, "configs": configs, "output_parity": output_parity,
        "setup_s": {"generate_inputs": t_gen, "bloom_build_gpu": t_build},
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_cfg4(args):
    """BASELINE configs[3]: `orpheum index` Bloom build from the 1e8-residue synthetic proteome on the GPU, then
    --total-reads counter-based 150-nt reads, generated in HBM shard by shard (never staged from the host) and scored;
    rank r owns the global read indices [r*T/W, (r+1)*T/W).  Timed: the scoring launches only (CUDA events around every
    batch, summed, max over ranks); read generation is input synthesis.  Every rank checks a sample of its last batch
    against the oracle on reads regenerated on the host from their global indices."""
    import torch
    import torch.distributed as dist

    from oracle import oracle as orc
    from orpheum_b200 import _lib, build, gpu, parallel, sequence_encodings as enc, synth

    build.build()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    stream = torch.cuda.Stream(device=dev)
    ctx = gpu.Context(local, stream=stream.cuda_stream)
    lib = _lib.load()
    lut, thr = enc.encode_lut(args.alphabet), default_threshold(args.alphabet)
    residues, poffs = synth.make_proteome(n_proteins=187000, seed=1004)
    tablesize = int(args.tablesize)
    graph = gpu.GpuNodegraph(args.ksize, tablesize, args.n_tables, ctx=ctx)
    ctx.synchronize()
    t_build = time.perf_counter()
    if rank == 0:
        graph.add_peptides(residues, poffs, lut, count_unique=False)
    ctx.synchronize()
    t_build = time.perf_counter() - t_build
    if world > 1:
        parallel.broadcast_filter(graph, src=0)
    total = int(args.total_reads)
    bounds = parallel.shard_bounds(total, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    batch = args.reads
    d_res = torch.from_numpy(residues).to(dev)
    d_poffs = torch.from_numpy(poffs.view(np.int64)).to(dev)
    d_syn = torch.from_numpy(synth.syn_table().view(np.int64)).to(dev)
    d_reads = torch.zeros((batch * READ_LEN + 31) // 32 + 2, dtype=torch.int64, device=dev)
    d_offsets = torch.arange(batch + 1, dtype=torch.int64, device=dev) * READ_LEN
    d_out = torch.zeros(batch * 32, dtype=torch.uint8, device=dev)
    seed = 3003
    events, n_cat = [], np.zeros(6, dtype=np.int64)
    launches0 = None

    def gen(first, m):
        _lib.check(lib.orph_synth_reads_device(ctx._h, seed, first, m, READ_LEN, d_res.data_ptr(), d_poffs.data_ptr(), len(poffs) - 1,
                                               d_syn.data_ptr(), d_reads.data_ptr()))

    with torch.cuda.stream(stream), ClockSampler(local) as clocks:
        gen(lo, min(batch, hi - lo))
        for _ in range(max(1, args.warmup)):
            gpu.score_reads_device(graph, d_reads.data_ptr(), d_offsets.data_ptr(), min(batch, hi - lo), READ_LEN, lut, thr, d_out.data_ptr(), ctx=ctx)
        ctx.check()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        launches0 = ctx.launch_count
        t_wall = time.perf_counter()
        first = lo
        while first < hi:
            m = min(batch, hi - first)
            gen(first, m)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            gpu.score_reads_device(graph, d_reads.data_ptr(), d_offsets.data_ptr(), m, READ_LEN, lut, thr, d_out.data_ptr(), ctx=ctx)
            e1.record(stream)
            events.append((e0, e1))
            last_first, last_m = first, m
            first += m
        ctx.check()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t_wall = time.perf_counter() - t_wall
    launches = ctx.launch_count - launches0
    ms_score = sum(a.elapsed_time(b) for a, b in events)
    if world > 1:
        t = torch.tensor([ms_score, t_wall], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_score, t_wall = float(t[0].item()), float(t[1].item())
    # parity: a sample of the LAST batch, regenerated on the host from the global indices
    res = d_out.cpu().numpy().view(_lib.READ_RESULT_DTYPE)[:last_m]
    pick = np.sort(np.random.default_rng(rank).choice(last_m, min(20_000, last_m), replace=False))
    host_reads = synth.counter_reads(last_first + pick.astype(np.uint64), residues, poffs, seed=seed, length=READ_LEN)
    dev_words = d_reads.cpu().numpy().view(np.uint64)
    flat, offs = synth.concat_fixed(host_reads)
    hp, _ = gpu.pack_reads(flat, offs)
    bits = lambda words, base, n: np.array([(int(words[(base + q) // 32]) >> (2 * ((base + q) % 32))) & 3 for q in range(n)])
    for s_i in (0, len(pick) // 2, len(pick) - 1):  # the device generator and its numpy twin agree base for base
        assert (bits(dev_words, int(pick[s_i]) * READ_LEN, READ_LEN) == bits(hp, s_i * READ_LEN, READ_LEN)).all(), "generator mismatch"
    g_cpu = orc.Nodegraph(args.ksize, tablesize, args.n_tables)
    g_cpu.add_peptides(residues, poffs, orc.encode_lut(args.alphabet))
    ora, _ = orc.score_reads(g_cpu, flat, offs, orc.encode_lut(args.alphabet), thr, n_threads=orc.max_threads() if world == 1 else 2)
    sub, scored = res[pick], ora["n_kmers"] >= 0
    ok = bool((sub["category"].astype(np.int64) == ora["category"]).all() and (sub["n_kmers"][scored] == ora["n_kmers"][scored]).all()
              and (sub["n_hits"][scored] == ora["n_hits"][scored]).all())
    if not ok:
        raise SystemExit(f"bench cfg4: rank {rank}: GPU results differ from the oracle on the sampled reads")
    if rank == 0:
        value = total / (ms_score * 1e-3)
        n_batches = len(events)
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": n_batches, "warmup": max(1, args.warmup),
            "ms_per_step": ms_score / n_batches, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": {"workload": f"cfg4: {total} counter-based synthetic 150-nt reads generated in HBM, sharded over {world} GPU(s) in batches "
                                   f"of {batch}, vs a Bloom filter built on the GPU from a {int(poffs[-1])}-residue synthetic proteome, "
                                   f"{args.alphabet} k={args.ksize}, tablesize {tablesize:.0e} x {args.n_tables} ({graph.device_buffer()[1] / 1e6:.0f} MB)",
                       "l2_flush": "each batch is freshly generated (375 MB packed) and exceeds L2"},
            "clocks": clocks.summary(), "gpu_launches": int(launches), "wall_s_including_generation": t_wall,
            "index_build_s": t_build, "index_residues_per_s": float(poffs[-1]) / t_build,
            "probes_min_per_read_sample": float(ora["probes_min"].sum()) / len(pick), "parity_checked_vs_oracle": ok,
            "e2e": None, "roofline": None, "cpu_baseline": None,
        }))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.ksize is None:
        args.ksize = default_ksize(args.alphabet)
    if args.tablesize is None:
        args.tablesize = 1e9 if args.workload == "cfg4" else 1e8
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "cfg4":
        run_cfg4(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
