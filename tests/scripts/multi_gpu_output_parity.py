#!/usr/bin/env python
"""cfg5 output parity on N GPUs (BASELINE configs[4]): mixed-length reads (50-300 nt, homopolymers, triplet repeats,
N, short) sharded over the ranks of one box, records gathered in global read order on rank 0 and written as CSV,
parquet and JSON summary through the columnar writer -- compared with (a) the files the row-list CreateSaveSummary
(the reference's object interface, itself pinned to the reference's golden outputs by tests/test_host_logic.py)
writes from ORACLE scores of the same reads, and (b) the single-GPU records.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/scripts/multi_gpu_output_parity.py [n_reads]
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def check(n_reads, rank, world, local, ctx=None):
    """Runs the comparison inside an initialised process group (or a single process); every rank calls it, rank 0 gets the
    report (the others None).  bench.py calls it at N > 1 (`output_parity`)."""
    import torch

    from oracle import oracle as orc
    from orpheum_b200 import _lib, gpu, parallel, sequence_encodings as enc, synth
    from orpheum_b200.columnar import ScoreTable
    from orpheum_b200.constants_translate import FRAMES, category_text
    from orpheum_b200.create_save_summary import CreateSaveSummary

    ctx = ctx or gpu.Context(local)
    alphabet, ksize, thr = "protein", 7, 0.5
    lut = enc.encode_lut(alphabet)
    residues, poffs = synth.make_proteome(n_proteins=20000, seed=1001)
    flat, offsets = synth.make_reads_mixed(n_reads, residues, poffs, ksize=ksize, seed=5005)
    graph = gpu.GpuNodegraph(ksize, int(1e8), 4, ctx=ctx)
    if rank == 0:
        graph.add_peptides(residues, poffs, lut)
    ctx.synchronize()
    if world > 1:
        parallel.broadcast_filter(graph, src=0)
    lo, hi, view = parallel.shard_reads(offsets, rank, world)
    mine = gpu.score_reads(graph, flat, view, lut, thr, ctx=ctx)
    full = parallel.gather_records(mine, dst=0, device=f"cuda:{local}") if world > 1 else mine
    if rank != 0:
        return None
    names = [f"read{i} mixed/{i % 7}" + (', "quoted"' if i % 1000 == 3 else "") for i in range(n_reads)]
    tmp = tempfile.mkdtemp(prefix="orph_mgpu_")
    try:
        table = ScoreTable(alphabet)
        table.append(names, full, "reads.fq")
        table.write_csv(os.path.join(tmp, "gpu.csv"))
        table.write_parquet(os.path.join(tmp, "gpu.parquet"))
        table.write_json_summary(os.path.join(tmp, "gpu.json"), ["reads.fq"], "filter.nodegraph", ksize, thr)
        # oracle -> the reference's six-element rows -> the row-list writers
        go = orc.Nodegraph(ksize, int(1e8), 4)
        go.add_peptides(residues, poffs, orc.encode_lut(alphabet))
        ora, _ = orc.score_reads(go, flat, offsets, orc.encode_lut(alphabet), thr, n_threads=orc.max_threads())
        rows = []
        for i in range(n_reads):
            for f in range(6):
                cat = int(ora["category"][i, f])
                nk, nh = int(ora["n_kmers"][i, f]), int(ora["n_hits"][i, f])
                jac = nh / nk if cat in (3, 4) else np.nan
                rows.append([names[i], jac, nk if cat in (2, 3, 4) else np.nan, category_text(cat, alphabet), FRAMES[f], "reads.fq"])
        ref = CreateSaveSummary(["reads.fq"], os.path.join(tmp, "ref.csv"), os.path.join(tmp, "ref.parquet"), os.path.join(tmp, "ref.json"),
                                "filter.nodegraph", alphabet, ksize, thr, rows)
        ref.maybe_write_csv()
        ref.maybe_write_parquet()
        ref.maybe_write_json_summary()
        import pyarrow.parquet as pq

        csv_same = open(os.path.join(tmp, "gpu.csv"), "rb").read() == open(os.path.join(tmp, "ref.csv"), "rb").read()
        a, b = pq.read_table(os.path.join(tmp, "gpu.parquet")), pq.read_table(os.path.join(tmp, "ref.parquet"))
        pq_same = a.schema.equals(b.schema) and a.to_pandas().equals(b.to_pandas())  # NaN-aware
        json_same = json.load(open(os.path.join(tmp, "gpu.json"))) == json.load(open(os.path.join(tmp, "ref.json")))
        single = gpu.score_reads(graph, flat, offsets, lut, thr, ctx=ctx)
        same_as_single = single.tobytes() == full.tobytes()
        cats = np.bincount(full["category"].reshape(-1), minlength=6).tolist()
        csv_bytes = os.path.getsize(os.path.join(tmp, "gpu.csv"))
    finally:
        import shutil

        shutil.rmtree(tmp, ignore_errors=True)
    return {"n_gpus": world, "n_reads": n_reads, "csv_identical": csv_same, "parquet_identical": bool(pq_same),
            "json_summary_identical": json_same, "records_identical_to_single_gpu": same_as_single, "csv_bytes": csv_bytes,
            "frames_per_category_code": cats, "byte_path_reads": int((full["flags"] & _lib.FLAG_BYTE_PATH != 0).sum()),
            "compared_with": "row-list CreateSaveSummary writers fed with oracle scores of the same reads"}


def main():
    import torch
    import torch.distributed as dist

    n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    report = check(n_reads, rank, world, local)
    ok = True
    if rank == 0:
        print(json.dumps(report))
        ok = all(report[k] for k in ("csv_identical", "parquet_identical", "json_summary_identical", "records_identical_to_single_gpu"))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if not ok:
        raise SystemExit("multi-GPU outputs differ")


if __name__ == "__main__":
    main()
