"""The fused file-to-records path (orph_stream_*: parallel FASTQ parse + 2-bit pack + H2D + kernels) must give exactly
the records of the batch call on the same reads -- which test_gpu_parity.py pins against the oracle -- for every file
layout: strict four-line FASTQ (plain, gzip, CRLF, no final newline), wrapped / blank-line FASTQ and FASTA through
the general parser, reads with N / lower case (byte path), empty files, and files far larger than one batch."""
import gzip
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    from oracle import oracle as orc
    from orpheum_b200 import build, gpu, sequence_encodings as enc, synth

    build.build()
    residues, poffs = synth.make_proteome(n_proteins=2000, seed=41)
    lut = enc.encode_lut("protein")
    graph = gpu.GpuNodegraph(7, 1e7, 4)
    graph.add_peptides(residues, poffs, lut)
    go = orc.Nodegraph(7, int(1e7), 4)
    go.add_peptides(residues, poffs, orc.encode_lut("protein"))
    return {"gpu": gpu, "orc": orc, "synth": synth, "lut": lut, "graph": graph, "oracle_graph": go, "residues": residues, "poffs": poffs}


def mixed_reads(env, n, seed):
    flat, offsets = env["synth"].make_reads_mixed(n, env["residues"], env["poffs"], ksize=7, seed=seed)
    flat = flat.copy()
    lower = np.random.default_rng(seed).choice(n, max(1, n // 500), replace=False)
    for r in lower:
        a, b = int(offsets[r]), int(offsets[r + 1])
        flat[a:b] = np.frombuffer(flat[a:b].tobytes().lower(), dtype=np.uint8)
    return flat, offsets


def write_fastq(path, flat, offsets, eol=b"\n", final_newline=True, wrap=0, blank_every=0, opener=open):
    data = flat.tobytes()
    parts = []
    n = len(offsets) - 1
    for i in range(n):
        s = data[int(offsets[i]):int(offsets[i + 1])]
        q = bytes([33 + (i + j) % 40 for j in range(len(s))])  # quality lines that start with '@' or '+' now and then
        if wrap:
            s_lines = eol.join(s[j:j + wrap] for j in range(0, len(s), wrap))
            q_lines = eol.join(q[j:j + wrap] for j in range(0, len(q), wrap))
        else:
            s_lines, q_lines = s, q
        parts.append(b"@r%d some description/1" % i + eol + s_lines + eol + b"+" + eol + q_lines + eol)
        if blank_every and i % blank_every == blank_every - 1 and i + 1 < n:
            parts.append(eol)
    blob = b"".join(parts)
    if not final_newline:
        blob = blob[: -len(eol)]
    with opener(path, "wb") as f:
        f.write(blob)


def stream_all(env, path, **kw):
    gpu = env["gpu"]
    names, parts, byte_path = [], [], 0
    with gpu.ScoringStream(env["graph"], path, env["lut"], 0.5, **kw) as st:
        for batch in st:
            assert batch.status == 0
            parts.append(batch.results.copy())
            names.extend(batch.names().tolist())
            byte_path += batch.n_byte_path
    res = np.concatenate(parts) if parts else np.zeros(0, dtype=gpu.READ_RESULT_DTYPE)
    return names, res, byte_path


@pytest.mark.parametrize("layout", ["plain", "crlf", "no_final_newline", "gz", "bgzf", "wrapped", "blank_lines_tail"])
def test_stream_matches_batch_call(env, tmp_path, layout):
    gpu = env["gpu"]
    n = 120_000
    flat, offsets = mixed_reads(env, n, seed=3)
    want = gpu.score_reads(env["graph"], flat, offsets, env["lut"], 0.5)
    path = str(tmp_path / ("reads.fq.gz" if layout in ("gz", "bgzf") else "reads.fq"))
    expect_n = n
    if layout == "plain":
        write_fastq(path, flat, offsets)
    elif layout == "crlf":
        write_fastq(path, flat, offsets, eol=b"\r\n")
    elif layout == "no_final_newline":
        write_fastq(path, flat, offsets, final_newline=False)
    elif layout == "gz":
        write_fastq(path, flat, offsets, opener=gzip.open)
    elif layout == "bgzf":  # blocked gzip: members inflated in parallel
        plain = str(tmp_path / "plain.fq")
        write_fastq(plain, flat, offsets)
        env["synth"].write_bgzf(path, open(plain, "rb").read())
    elif layout == "wrapped":
        write_fastq(path, flat, offsets, wrap=60)
    else:  # strict records, then a blank line in the middle of the file: the general parser takes over there
        write_fastq(path, flat, offsets, blank_every=70_000)
    for batch_reads in (1 << 20, 30_000):
        names, res, byte_path = stream_all(env, path, batch_reads=batch_reads)
        assert len(res) == expect_n, layout
        assert names[0] == "r0 some description/1" and names[-1] == f"r{n - 1} some description/1"
        assert res.tobytes() == want.tobytes(), (layout, batch_reads)
        assert byte_path >= n // 500
    # a sample against the oracle as well
    sample = np.sort(np.random.default_rng(1).choice(n, 5000, replace=False))
    sflat, soffs = gpu.concat_reads([flat[int(offsets[r]):int(offsets[r + 1])].tobytes() for r in sample])
    ora, _ = env["orc"].score_reads(env["oracle_graph"], sflat, soffs, env["orc"].encode_lut("protein"), 0.5, n_threads=0)
    cat = res[sample]["category"].astype(np.int64)
    np.testing.assert_array_equal(cat, ora["category"])
    scored = ora["n_kmers"] >= 0
    np.testing.assert_array_equal(res[sample]["n_hits"][scored], ora["n_hits"][scored])
    np.testing.assert_array_equal(res[sample]["n_kmers"][scored], ora["n_kmers"][scored])


def test_stream_fasta_empty_and_errors(env, tmp_path):
    gpu = env["gpu"]
    flat, offsets = mixed_reads(env, 5000, seed=9)
    want = gpu.score_reads(env["graph"], flat, offsets, env["lut"], 0.5)
    fa = str(tmp_path / "reads.fa")
    data = flat.tobytes()
    with open(fa, "wb") as f:
        for i in range(len(offsets) - 1):
            s = data[int(offsets[i]):int(offsets[i + 1])]
            f.write(b">q%d\n" % i + b"\n".join(s[j:j + 70] for j in range(0, len(s), 70)) + b"\n")
    names, res, _ = stream_all(env, fa)
    assert names[:2] == ["q0", "q1"] and res.tobytes() == want.tobytes()
    empty = str(tmp_path / "empty.fq")
    open(empty, "wb").close()
    assert stream_all(env, empty)[1].size == 0
    junk = str(tmp_path / "junk.txt")
    with open(junk, "w") as f:
        f.write("not a sequence file\n")
    with pytest.raises(ValueError):
        gpu.ScoringStream(env["graph"], junk, env["lut"], 0.5)
    with pytest.raises(OSError):
        gpu.ScoringStream(env["graph"], str(tmp_path / "missing.fq"), env["lut"], 0.5)
    # quality shorter than the sequence: the reference's reader raises; so does the stream, at that record
    bad = str(tmp_path / "bad.fq")
    with open(bad, "w") as f:
        f.write("@a\nACGTACGTACGTACGTACGTACGTACGTACGT\n+\nIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII\n@b\nACGTACGT\n+\nIII\n")
    with pytest.raises(OSError):
        stream_all(env, bad)
    # an empty read is the reference's ZeroDivisionError: reported through the batch status
    zero = str(tmp_path / "zero.fa")
    with open(zero, "w") as f:
        f.write(">a\nACGTACGTACGTACGTACGTACGTACGTACGTAAAAAAAACCCCCGGGGGTTTTTACGATCGATCGACTG\n>b\n\n>c\nACGTAGCTAGCTAGCTAGCTAGCTGATCGATCGTAGCTAGCTAGCTAGCTAGCTAGCTGATCG\n")
    with gpu.ScoringStream(env["graph"], zero, env["lut"], 0.5) as st:
        statuses = [b.status for b in st]
    from orpheum_b200 import _lib

    assert statuses == [_lib.ORPH_E_EMPTY_READ]


def test_stream_early_close_and_reads_accessor(env, tmp_path):
    gpu = env["gpu"]
    flat, offsets = mixed_reads(env, 200_000, seed=5)
    path = str(tmp_path / "reads.fq")
    write_fastq(path, flat, offsets)
    st = gpu.ScoringStream(env["graph"], path, env["lut"], 0.5, batch_reads=20_000)
    it = iter(st)
    first = next(it)
    sub_flat, sub_off = first.reads(np.array([0, 7, first.n_reads - 1]))
    data = flat.tobytes()
    assert sub_flat.tobytes() == b"".join(data[int(offsets[r]):int(offsets[r + 1])] for r in (0, 7, first.n_reads - 1))
    assert sub_off.tolist()[0] == 0 and len(sub_off) == 4
    st.close()  # with batches still in flight
    # two streams at once on one device, and the caller's context used in between
    a = gpu.ScoringStream(env["graph"], path, env["lut"], 0.5, batch_reads=50_000)
    b = gpu.ScoringStream(env["graph"], path, env["lut"], 0.5, batch_reads=70_000)
    ra, rb = [], []
    for x, y in zip(a, b):
        ra.append(x.results.copy())
        rb.append(y.results.copy())
        gpu.score_reads(env["graph"], flat[:1500], offsets[:11], env["lut"], 0.5)
    a.close()
    b.close()
    want = gpu.score_reads(env["graph"], flat, offsets, env["lut"], 0.5)
    na, nb = sum(len(x) for x in ra), sum(len(x) for x in rb)
    assert np.concatenate(ra).tobytes() == want[:na].tobytes() and np.concatenate(rb).tobytes() == want[:nb].tobytes()


@pytest.mark.parametrize("n_lanes", [2, 3])
def test_multi_lane_stream_delivers_the_same_batches_in_file_order(env, tmp_path, n_lanes):
    """orph_stream_open_multi: one parser dealing batches round-robin to several device lanes (here: further private contexts
    on the one GPU of the test box, the filter replicated with orph_filter_replicate) must deliver exactly the single-lane
    stream -- same records, same names, same order -- including byte-path reads and a short last batch."""
    gpu = env["gpu"]
    n = 150_000
    flat, offsets = mixed_reads(env, n, seed=17)
    path = str(tmp_path / "reads.fq")
    write_fastq(path, flat, offsets)
    names1, res1, bp1 = stream_all(env, path, batch_reads=8192)
    lanes = gpu.replica_lanes(env["graph"], n_lanes, devices=[0] * (n_lanes - 1))
    for _, g in lanes:
        assert [g.n_occupied(i) for i in range(4)] == [env["graph"].n_occupied(i) for i in range(4)]
    names2, res2, bp2 = stream_all(env, path, batch_reads=8192, replicas=lanes)
    assert len(res2) == n and names2 == names1 and bp2 == bp1 > 0
    assert res2.tobytes() == res1.tobytes()
    want = gpu.score_reads(env["graph"], flat, offsets, env["lut"], 0.5)
    assert res2.tobytes() == want.tobytes()
    # a caller that abandons the stream half way: close() must drain every lane
    with gpu.ScoringStream(env["graph"], path, env["lut"], 0.5, batch_reads=8192, replicas=lanes) as st:
        for i, batch in enumerate(st):
            if i == 3:
                break


def test_stream_emit_matches_the_python_side_formatting(env, tmp_path):
    """orph_stream_emit (emitting frames found in the records, reads gathered from the text, one orph_translate_frames
    call, block-parallel formatting) against the pieces it replaces, called one by one from Python: orph_gather_reads ->
    orph_translate_frames -> orph_format_fasta, for all four kinds, with and without low-complexity peptides among the
    translated frames, on batches with N / lower-case reads, short reads and reads that emit nothing."""
    gpu = env["gpu"]
    from orpheum_b200 import _lib
    from orpheum_b200.constants_translate import FRAMES

    n = 60_000
    flat, offsets = mixed_reads(env, n, seed=23)
    path = str(tmp_path / "reads.fq")
    write_fastq(path, flat, offsets)
    ctx = env["graph"].ctx
    for kinds in (15, 1, 9, 6):
        with_lcp = bool(kinds & 8)
        total = [0, 0, 0, 0]
        with gpu.ScoringStream(env["graph"], path, env["lut"], 0.5, batch_reads=16384) as st:
            for batch in st:
                got = st.emit(ctx, kinds)
                results = batch.results.copy()
                names = batch.names()
                rflat, roffs = batch.reads()
                cat = results["category"]
                want = cat == _lib.CAT_CODING
                if with_lcp:
                    want = want | (cat == _lib.CAT_LOW_COMPLEXITY_PEPTIDE)
                reads_idx, frames_idx = np.nonzero(want)
                keys = np.asarray(FRAMES, dtype=np.int8)[frames_idx]
                peptides, pep_offsets = gpu.translate_frames_raw(ctx, rflat, roffs, reads_idx, keys)
                blob = np.frombuffer(names.blob, dtype=np.uint8) if len(names.blob) else np.zeros(1, dtype=np.uint8)
                for kind in range(4):
                    if not kinds & (1 << kind):
                        assert got[kind] == b""
                        continue
                    expect = gpu.format_fasta(kind, blob, names.offsets, results, rflat, roffs, peptides, pep_offsets, with_lcp)
                    assert got[kind] == expect, (kinds, kind)
                    total[kind] += len(expect)
        assert (not kinds & 1 or total[0] > 100_000) and (not kinds & 4 or total[2] > 100_000) and (not kinds & 8 or total[3] > 0)
