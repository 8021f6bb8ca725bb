"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the committed golden
vectors.  Integer / byte work: everything is compared bit-exactly."""
import hashlib
import math
import os

import numpy as np
import pandas as pd
import pytest

from tests.conftest import SCORE_VECTORS, load_score_vector, read_fastx

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from orpheum_b200 import build, gpu as g

    build.build()
    return g


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle

    return oracle


@pytest.fixture(scope="module")
def enc():
    from orpheum_b200 import sequence_encodings

    return sequence_encodings


def fasta_arrays(path):
    recs = read_fastx(path)
    seqs = [s.encode() for _, s in recs]
    offsets = np.zeros(len(seqs) + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum([len(s) for s in seqs])
    return np.frombuffer(b"".join(seqs), dtype=np.uint8), offsets


def oracle_filter(orc, fasta, alphabet, ksize, tablesize=1e6, n_tables=4):
    g = orc.Nodegraph(ksize, tablesize, n_tables)
    res, offs = fasta_arrays(fasta)
    g.add_peptides(res, offs, orc.encode_lut(alphabet))
    return g


def gpu_filter(gpu, enc, fasta, alphabet, ksize, tablesize=1e6, n_tables=4):
    g = gpu.GpuNodegraph(ksize, tablesize, n_tables)
    res, offs = fasta_arrays(fasta)
    g.add_peptides(res, offs, enc.encode_lut(alphabet))
    return g


def assert_same_as_oracle(res, ora, where=""):
    """res: READ_RESULT_DTYPE [n]; ora: oracle FRAME_SCORE_DTYPE [n, 6]."""
    from orpheum_b200 import _lib

    cat = res["category"].astype(np.int64)
    np.testing.assert_array_equal(cat, ora["category"], err_msg=f"category {where}")
    scored = ora["n_kmers"] >= 0
    np.testing.assert_array_equal(res["n_kmers"][scored], ora["n_kmers"][scored], err_msg=f"n_kmers {where}")
    np.testing.assert_array_equal(res["n_hits"][scored], ora["n_hits"][scored], err_msg=f"n_hits {where}")
    assert (res["n_kmers"][~scored] == 0).all() and (res["n_hits"][~scored] == 0).all()
    # best frame: exact rational argmax over coding / non-coding frames, first on ties
    frames = np.array(_lib.FRAMES)
    for i in range(len(res)):
        best, bh, bn = 0, 0, 0
        for f in range(6):
            if ora["category"][i, f] in (3, 4):
                h, n = int(ora["n_hits"][i, f]), int(ora["n_kmers"][i, f])
                if best == 0 or h * bn > bh * n:
                    best, bh, bn = int(frames[f]), h, n
        assert int(res["best_frame"][i]) == best, (where, i)


# ---- hash ------------------------------------------------------------------------------------------
def test_device_murmur(gpu, orc, misc_kats):
    ctx = gpu.default_context()
    assert int(ctx.hash_kmers(["ACG"])[0]) == 1731421407650554201
    for key, value in misc_kats["hash_murmur"].items():
        assert int(ctx.hash_kmers([key])[0]) == value, key
    rng = np.random.default_rng(1)
    for length in list(range(0, 40)) + [47, 48, 49, 63, 64]:
        keys = [bytes(rng.integers(0, 256, length, dtype=np.uint8)) for _ in range(50)]
        got = ctx.hash_kmers(keys)
        assert [int(x) for x in got] == [orc.hash_murmur(k) for k in keys], length


# ---- filter ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("alphabet,ksize", [("protein", 7), ("dayhoff", 12)])
def test_load_query_save_golden_nodegraph(gpu, orc, refdata, tmp_path, alphabet, ksize):
    path = os.path.join(refdata, "index", f"Homo_sapiens.GRCh38.pep.subset.alphabet-{alphabet}_ksize-{ksize}.bloomfilter.nodegraph")
    g = gpu.GpuNodegraph.load(path)
    o = orc.Nodegraph.load(path)
    assert g.ksize() == ksize and g.hashsizes() == o.hashsizes() and g.n_tables() == 4
    assert g.n_occupied() == o.n_occupied()
    rng = np.random.default_rng(2)
    hashes = rng.integers(0, 2**64, 20000, dtype=np.uint64)
    # plus hashes that are certainly present
    lut = orc.encode_lut(alphabet)
    recs = [r for r in read_fastx(os.path.join(refdata, "index", "Homo_sapiens.GRCh38.pep.subset.fa.gz")) if "*" not in r[1]][:30]
    present = []
    for _, s in recs:
        e = bytes(lut[np.frombuffer(s.encode(), dtype=np.uint8)])
        present += [orc.hash_murmur(e[i:i + ksize]) for i in range(len(e) - ksize + 1)]
    allh = np.concatenate([hashes, np.array(present, dtype=np.uint64)])
    got = g.get_many(allh)
    want = np.array([o.get(int(h)) for h in allh], dtype=np.uint8)
    np.testing.assert_array_equal(got, want)
    assert got[len(hashes):].mean() > 0.5
    assert g.get(int(allh[-1])) == o.get(int(allh[-1]))
    out = tmp_path / "resaved.nodegraph"
    g.save(str(out))
    assert out.read_bytes() == open(path, "rb").read()
    with pytest.raises(OSError):
        gpu.GpuNodegraph.load(os.path.join(refdata, "index", "Homo_sapiens.GRCh38.pep.first1000lines.fa"))


@pytest.mark.parametrize("alphabet,ksize", [("protein", 7), ("dayhoff", 12)])
def test_build_golden_nodegraph_byte_identical(gpu, enc, refdata, tmp_path, alphabet, ksize):
    """orpheum index on the GPU reproduces the reference's golden .nodegraph files byte for byte."""
    golden = os.path.join(refdata, "index", f"Homo_sapiens.GRCh38.pep.subset.alphabet-{alphabet}_ksize-{ksize}.bloomfilter.nodegraph")
    g = gpu_filter(gpu, enc, os.path.join(refdata, "index", "Homo_sapiens.GRCh38.pep.subset.fa.gz"), alphabet, ksize)
    out = tmp_path / "built.nodegraph"
    g.save(str(out))
    assert out.read_bytes() == open(golden, "rb").read()
    expected = {("protein", 7): 506352, ("dayhoff", 12): 488469}[(alphabet, ksize)]
    np.testing.assert_allclose(g.n_unique_kmers(), expected, rtol=1e-3)  # tests/test_index.py:54


def test_build_index_kats(gpu, enc, refdata, index_kats, tmp_path):
    fasta = os.path.join(refdata, "index", "Homo_sapiens.GRCh38.pep.first1000lines.fa")
    for kat in index_kats:
        g = gpu_filter(gpu, enc, fasta, kat["alphabet"], kat["ksize"], kat["tablesize"], kat["n_tables"])
        assert [g.n_occupied(i) for i in range(kat["n_tables"])] == kat["popcounts"], kat
        out = tmp_path / "k.nodegraph"
        g.save(str(out))
        assert hashlib.sha256(out.read_bytes()).hexdigest() == kat["sha256"], kat
        # n_unique_kmers is insertion-order dependent through Bloom false positives: the
        # reference's own tolerance at tablesize 1e6; looser for the deliberately over-full tiny geometry
        rtol = 1e-3 if kat["tablesize"] >= 1000000 else 2e-2
        np.testing.assert_allclose(g.n_unique_kmers(), kat["n_unique_kmers"], rtol=rtol)


def test_add_get_single_hash_and_replicate(gpu):
    g = gpu.GpuNodegraph(7, 1e4, 4)
    assert g.get(12345) == 0
    assert g.add(12345) is True
    assert g.add(12345) is False
    assert g.get(12345) == 1 and g.n_unique_kmers() == 1
    assert g.n_occupied() == 1
    copy = g.replicate(gpu.Context(0))
    assert copy.get(12345) == 1 and copy.get(54321) == 0 and copy.hashsizes() == g.hashsizes()
    ptr, nbytes, offs = g.device_buffer()
    assert ptr != 0 and nbytes >= sum(s // 8 + 1 for s in g.hashsizes()) and offs[0] == 0


# ---- scoring ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("alphabet,ksize", [("protein", 7), ("dayhoff", 12)])
def test_golden_csv_23_reads(gpu, enc, refdata, reads_n22, misc_kats, alphabet, ksize):
    """The reference's own end-to-end golden (tests/test_translate.py:279-315), through the GPU."""
    from orpheum_b200 import _lib

    g = gpu.GpuNodegraph.load(os.path.join(
        refdata, "index", f"Homo_sapiens.GRCh38.pep.subset.alphabet-{alphabet}_ksize-{ksize}.bloomfilter.nodegraph"))
    true = pd.read_csv(os.path.join(
        refdata, "translate", f"SRR306838_GSM752691_hsa_br_F_1_trimmed_subsampled_n22__alphabet-{alphabet}_ksize-{ksize}.csv"))
    recs = read_fastx(reads_n22)
    flat, offsets = gpu.concat_reads([s for _, s in recs])
    res = gpu.score_reads(g, flat, offsets, enc.encode_lut(alphabet), 0.5)
    jac = gpu.jaccard_column(res)
    text = {0: "Translation frame has stop codon(s)", 1: "Translation is shorter than peptide k-mer size + 1",
            2: misc_kats["low_complexity_categories"][alphabet], 3: "Coding", 4: "Non-coding",
            5: "Read length was shorter than 3 * peptide k-mer size"}
    rows = [(recs[i][0], jac[i, f], int(res["n_kmers"][i, f]), text[int(res["category"][i, f])], _lib.FRAMES[f])
            for i in range(len(recs)) for f in range(6)]
    assert len(rows) == len(true) == 138
    for (name, j, n, cat, frame), t in zip(rows, true.itertuples()):
        assert (name, cat, frame) == (t.read_id, t.category, t.translation_frame)
        if math.isnan(t.n_kmers):
            assert n == 0
        else:
            assert n == int(t.n_kmers)
        if math.isnan(t.jaccard_in_peptide_db):
            assert math.isnan(j)
        else:
            assert j == pytest.approx(t.jaccard_in_peptide_db, abs=1e-12)


@pytest.mark.parametrize("force_exact", [False, True])
@pytest.mark.parametrize("alphabet,ksize", SCORE_VECTORS)
def test_reference_vectors(gpu, enc, orc, refdata, misc_kats, alphabet, ksize, force_exact):
    """Rows produced by the UNMODIFIED reference (tests/golden/ref_scores_*.json): categories, n_kmers,
    integer hit counts and jaccard, for every read incl. N / lower-case / short / low-complexity ones."""
    n_bytes_path, n_reads = check_reference_vector(gpu, enc, refdata, misc_kats, load_score_vector(alphabet, ksize), alphabet, ksize, force_exact)
    assert 0 < n_bytes_path < n_reads  # both the packed fast path and the byte path were exercised


@pytest.mark.parametrize("force_exact", [False, True])
def test_reference_vectors_long_reads(gpu, enc, refdata, misc_kats, force_exact):
    """Reads of 6.1 - 9 kb scored by the UNMODIFIED reference (tests/golden/ref_scores_long_protein_k7.json): the
    block-per-read kernel against the reference itself, not only against the oracle."""
    vec = load_score_vector("protein", 7, tag="long_")
    n_bytes_path, n_reads = check_reference_vector(gpu, enc, refdata, misc_kats, vec, "protein", 7, force_exact)
    assert n_bytes_path == n_reads == 8


def check_reference_vector(gpu, enc, refdata, misc_kats, vec, alphabet, ksize, force_exact):
    g = gpu_filter(gpu, enc, os.path.join(refdata, vec["peptides"]), alphabet, ksize, vec["tablesize"], vec["n_tables"])
    assert [g.n_occupied(i) for i in range(4)] == vec["filter_popcounts"]
    flat, offsets = gpu.concat_reads([r["seq"] for r in vec["reads"]])
    ctx = g.ctx
    ctx.set_force_exact_dedup(force_exact)
    try:
        res = gpu.score_reads(g, flat, offsets, enc.encode_lut(alphabet), vec["jaccard_threshold"])
    finally:
        ctx.set_force_exact_dedup(False)
    jac = gpu.jaccard_column(res)
    text = {0: "Translation frame has stop codon(s)", 1: "Translation is shorter than peptide k-mer size + 1",
            2: misc_kats["low_complexity_categories"][alphabet], 3: "Coding", 4: "Non-coding",
            5: "Read length was shorter than 3 * peptide k-mer size"}
    n_bytes_path = 0
    for i, read in enumerate(vec["reads"]):
        n_bytes_path += bool(res["flags"][i] & 0x80)
        for f, (j, n, cat, frame, hits) in enumerate(read["rows"]):
            assert text[int(res["category"][i, f])] == cat, (read["name"], frame)
            assert int(res["n_kmers"][i, f]) == (n or 0), (read["name"], frame)
            assert int(res["n_hits"][i, f]) == (hits or 0), (read["name"], frame)
            if j is None:
                assert math.isnan(jac[i, f])
            else:
                assert jac[i, f] == j
    return n_bytes_path, len(vec["reads"])


def synthetic_case(gpu, enc, orc, refdata, alphabet, ksize, thr, n, mixed):
    from orpheum_b200 import synth

    fasta = os.path.join(refdata, "index", "Homo_sapiens.GRCh38.pep.subset.fa.gz")
    res_, offs_ = fasta_arrays(fasta)
    # proteome without the 4 records that contain '*'
    go = oracle_filter(orc, fasta, alphabet, ksize, 1e6)
    gg = gpu_filter(gpu, enc, fasta, alphabet, ksize, 1e6)
    standard = np.zeros(256, dtype=bool)
    standard[np.frombuffer(synth.AMINO_ACIDS.encode(), dtype=np.uint8)] = True
    # back-translation needs the 20 standard residues only (the real proteome has X / U / '*')
    keep = [i for i in range(len(offs_) - 1)
            if offs_[i + 1] - offs_[i] >= 120 and standard[res_[int(offs_[i]):int(offs_[i + 1])]].all()]
    residues = np.concatenate([res_[int(offs_[i]):int(offs_[i + 1])] for i in keep])
    poffs = np.zeros(len(keep) + 1, dtype=np.uint64)
    poffs[1:] = np.cumsum([int(offs_[i + 1] - offs_[i]) for i in keep])
    if mixed:
        flat, offsets = synth.make_reads_mixed(n, residues, poffs, ksize=ksize, seed=5005)
        nz = np.nonzero(np.diff(offsets.astype(np.int64)) > 0)[0]
        assert len(nz) == n
    else:
        flat, offsets = synth.concat_fixed(synth.make_reads_fixed(n, residues, poffs, length=150, seed=2002))
    return go, gg, flat, offsets


@pytest.mark.parametrize("alphabet,ksize,thr", [("protein", 7, 0.5), ("dayhoff", 12, 0.5), ("hp", 31, 0.8)])
def test_synthetic_r150_vs_oracle(gpu, enc, orc, refdata, alphabet, ksize, thr):
    """cfg2 / cfg3 shape: 150-nt reads, half back-translated from the proteome; ASCII and packed inputs."""
    n = 20000
    go, gg, flat, offsets = synthetic_case(gpu, enc, orc, refdata, alphabet, ksize, thr, n, mixed=False)
    lut = enc.encode_lut(alphabet)
    ora, n_empty = orc.score_reads(go, flat, offsets, orc.encode_lut(alphabet), thr, n_threads=0)
    assert n_empty == 0
    res = gpu.score_reads(gg, flat, offsets, lut, thr)
    assert_same_as_oracle(res, ora, f"ascii {alphabet}")
    assert (ora["category"] == 3).sum() > n // 10
    packed, needs = gpu.pack_reads(flat, offsets)
    assert not needs.any()
    res2 = gpu.score_reads(gg, packed, offsets, lut, thr, reads_format=1)
    assert res2.tobytes() == res.tobytes()
    assert not (res["flags"] & 0x80).any()
    # the warp-per-read kernel (used for reads longer than 320 nt) must agree with the thread-per-frame kernel
    gg.ctx.set_thread_path(False)
    try:
        res3 = gpu.score_reads(gg, packed, offsets, lut, thr, reads_format=1)
    finally:
        gg.ctx.set_thread_path(True)
    assert res3.tobytes() == res.tobytes()


@pytest.mark.parametrize("alphabet,ksize,thr", [("protein", 7, 0.5), ("dayhoff", 12, 0.5), ("hp", 31, 0.8)])
def test_synthetic_mixed_vs_oracle(gpu, enc, orc, refdata, alphabet, ksize, thr):
    """cfg5 shape: 50-300 nt, homopolymers, triplet repeats, N bases, too-short reads."""
    n = 20000
    go, gg, flat, offsets = synthetic_case(gpu, enc, orc, refdata, alphabet, ksize, thr, n, mixed=True)
    ora, n_empty = orc.score_reads(go, flat, offsets, orc.encode_lut(alphabet), thr, n_threads=0)
    assert n_empty == 0
    res = gpu.score_reads(gg, flat, offsets, enc.encode_lut(alphabet), thr)
    assert_same_as_oracle(res, ora, f"mixed {alphabet}")
    gg.ctx.set_thread_path(False)
    try:
        res_warp = gpu.score_reads(gg, flat, offsets, enc.encode_lut(alphabet), thr)
    finally:
        gg.ctx.set_thread_path(True)
    # identical but for the route flag: reads with N are scored thread-per-frame with an N plane on the 2-bit path, and by the
    # byte-exact kernel when that path is switched off
    strip = lambda r: np.concatenate([r[name].reshape(len(r), -1).astype(np.int64) for name in ("n_kmers", "n_hits", "category", "best_frame")]
                                     + [(r["flags"] & 0x7F).reshape(-1, 1).astype(np.int64)], axis=1)
    np.testing.assert_array_equal(strip(res_warp), strip(res))
    cats = set(np.unique(ora["category"]).tolist())
    assert {0, 1, 2, 3, 4, 5} <= cats or alphabet == "hp"
    has_n = np.array([b"N" in flat[int(offsets[i]):int(offsets[i + 1])].tobytes() for i in range(n)])
    assert has_n.any() and not (res["flags"] & 0x80).any()           # N reads of <= 320 nt stay on the 2-bit path ...
    np.testing.assert_array_equal((res_warp["flags"] & 0x80) != 0, has_n)  # ... and are the byte path's without the thread kernel


@pytest.mark.parametrize("alphabet,ksize", [("protein", 7), ("dayhoff", 12), ("hp", 31)])
def test_reads_with_n_on_the_two_bit_path(gpu, enc, orc, refdata, alphabet, ksize):
    """Reads that hold N (and no other unusual byte) are scored by the thread-per-frame kernel with an N plane: every N
    placement that matters -- third base of each of the 16 two-base prefixes (seven of them translate, CCN and the rest are
    X), first and second codon positions, both strands, the first and last base of the read, runs of N, all-N reads, N in
    reads of every length up to 320 nt and just beyond (byte path) -- against the oracle."""
    from orpheum_b200 import synth

    fasta = os.path.join(refdata, "index", "Homo_sapiens.GRCh38.pep.subset.fa.gz")
    go = oracle_filter(orc, fasta, alphabet, ksize, 1e6)
    gg = gpu_filter(gpu, enc, fasta, alphabet, ksize, 1e6)
    res_, offs_ = fasta_arrays(fasta)
    standard = np.zeros(256, dtype=bool)
    standard[np.frombuffer(synth.AMINO_ACIDS.encode(), dtype=np.uint8)] = True
    keep = [i for i in range(len(offs_) - 1) if offs_[i + 1] - offs_[i] >= 120 and standard[res_[int(offs_[i]):int(offs_[i + 1])]].all()]
    residues = np.concatenate([res_[int(offs_[i]):int(offs_[i + 1])] for i in keep])
    poffs = np.zeros(len(keep) + 1, dtype=np.uint64)
    poffs[1:] = np.cumsum([int(offs_[i + 1] - offs_[i]) for i in keep])
    rng = np.random.default_rng(77)
    reads = []
    coding = synth.make_reads_fixed(6000, residues, poffs, length=330, seed=78)  # even rows are back-translated, odd rows random
    for i in range(6000):
        L = int(rng.choice([31, 64, 95, 96, 150, 151, 191, 192, 193, 250, 319, 320, 321, 330]))
        s = bytearray(coding[i, :L].tobytes())
        mode = (i // 2) % 8  # (even rows are coding reads, odd rows random: both kinds get every mode)
        if mode == 0:
            s[int(rng.integers(0, L))] = ord("N")
        elif mode == 1:
            for p in rng.integers(0, L, int(rng.integers(2, 7))):
                s[int(p)] = ord("N")
        elif mode == 2:
            a = int(rng.integers(0, L))
            b_ = min(L, a + int(rng.integers(2, 40)))
            s[a:b_] = b"N" * (b_ - a)
        elif mode == 3:
            s[0] = ord("N")
            s[L - 1] = ord("N")
        elif mode == 4:  # every third base from a random phase: one N per codon of one frame
            for p in range(int(rng.integers(0, 3)), L, 3):
                if rng.random() < 0.3:
                    s[p] = ord("N")
        elif mode == 5:
            s = bytearray(b"N" * L)
        elif mode == 6:  # all 16 prefixes followed by N, back to back
            pre = b"".join(bytes([a, b]) + b"N" for a in b"ACGT" for b in b"ACGT")
            at = int(rng.integers(0, max(1, L - len(pre))))
            s[at:at + len(pre)] = pre[: L - at]
        reads.append(bytes(s))
    flat, offsets = gpu.concat_reads(reads)
    olut = orc.encode_lut(alphabet)
    thr = 0.8 if alphabet == "hp" else 0.5
    ora, n_empty = orc.score_reads(go, flat, offsets, olut, thr, n_threads=0)
    assert n_empty == 0
    res = gpu.score_reads(gg, flat, offsets, enc.encode_lut(alphabet), thr)
    assert_same_as_oracle(res, ora, f"N reads {alphabet}")
    lengths = np.diff(offsets.astype(np.int64))
    has_n = np.array([b"N" in r for r in reads])
    byte_path = (res["flags"] & 0x80) != 0
    np.testing.assert_array_equal(byte_path, has_n & (lengths > 320))
    scored_with_n = ((ora["category"] == 3) | (ora["category"] == 4)).any(axis=1) & has_n & ~byte_path
    assert scored_with_n.sum() > 300  # frames of N reads were really scored on the new path
    # the same batch as 2-bit words + N plane resident in HBM (orph_pack_reads_device -> orph_score_reads_device_n)
    import torch

    from orpheum_b200._lib import READ_RESULT_DTYPE

    n = len(reads)
    dev = torch.device("cuda", gg.ctx.device)
    d_ascii = torch.from_numpy(flat.copy()).to(dev)
    d_off = torch.from_numpy(offsets.view(np.int64)).to(dev)
    total = int(offsets[-1])
    d_packed = torch.zeros((total + 31) // 32 + 2, dtype=torch.int64, device=dev)
    d_nmask = torch.zeros((total + 31) // 32 + 2, dtype=torch.int32, device=dev)
    d_has_n = torch.zeros(n, dtype=torch.uint8, device=dev)
    d_needs = torch.zeros(n, dtype=torch.uint8, device=dev)
    d_out = torch.zeros(n * 32, dtype=torch.uint8, device=dev)
    gpu.pack_reads_device(gg.ctx, d_ascii.data_ptr(), d_off.data_ptr(), n, total, d_packed.data_ptr(), d_nmask.data_ptr(), d_has_n.data_ptr(), d_needs.data_ptr())
    gpu.score_reads_device_n(gg, d_packed.data_ptr(), d_nmask.data_ptr(), d_has_n.data_ptr(), d_off.data_ptr(), n, int(lengths.max()),
                             enc.encode_lut(alphabet), thr, d_out.data_ptr())
    gg.ctx.check()
    assert not d_needs.cpu().numpy().any()
    np.testing.assert_array_equal(d_has_n.cpu().numpy().astype(bool), has_n)
    assert d_out.cpu().numpy().view(READ_RESULT_DTYPE).tobytes() == res.tobytes()


def test_sub_batch_offsets_and_device_api(gpu, enc, orc, refdata):
    """offsets[0] != 0 (a view into a larger buffer) and the device-pointer entry point."""
    import torch

    go, gg, flat, offsets = synthetic_case(gpu, enc, orc, refdata, "protein", 7, 0.5, 3000, mixed=True)
    lut = enc.encode_lut("protein")
    ora, _ = orc.score_reads(go, flat, offsets, orc.encode_lut("protein"), 0.5)
    lo, hi = 1001, 2500
    res = gpu.score_reads(gg, flat, offsets[lo:hi + 1], lut, 0.5)
    assert_same_as_oracle(res, ora[lo:hi], "sub-batch ascii")
    # packed sub-batch on clean reads only
    go2, gg2, flat2, offsets2 = synthetic_case(gpu, enc, orc, refdata, "protein", 7, 0.5, 3000, mixed=False)
    ora2, _ = orc.score_reads(go2, flat2, offsets2, orc.encode_lut("protein"), 0.5)
    packed, _ = gpu.pack_reads(flat2, offsets2)
    res2 = gpu.score_reads(gg2, packed, offsets2[lo:hi + 1], lut, 0.5, reads_format=1)
    assert_same_as_oracle(res2, ora2[lo:hi], "sub-batch packed")
    # device API (torch only provides the memory)
    dev = torch.device("cuda", 0)
    d_packed = torch.from_numpy(packed.view(np.int64)).to(dev)
    d_offsets = torch.from_numpy(offsets2.view(np.int64)).to(dev)
    d_out = torch.zeros(3000 * 32, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    gpu.score_reads_device(gg2, d_packed.data_ptr(), d_offsets.data_ptr(), 3000, 150, lut, 0.5, d_out.data_ptr())
    gg2.ctx.check()
    from orpheum_b200._lib import READ_RESULT_DTYPE
    res3 = d_out.cpu().numpy().view(READ_RESULT_DTYPE)
    assert_same_as_oracle(res3, ora2, "device packed")
    d_ascii = torch.from_numpy(flat2).to(dev)
    d_out.zero_()
    torch.cuda.synchronize()
    gpu.score_reads_device(gg2, d_ascii.data_ptr(), d_offsets.data_ptr(), 3000, 0, lut, 0.5, d_out.data_ptr(), reads_format=0)
    gg2.ctx.check()
    assert d_out.cpu().numpy().view(READ_RESULT_DTYPE).tobytes() == res3.tobytes()


def test_edge_cases(gpu, enc, orc, refdata):
    fasta = os.path.join(refdata, "index", "Homo_sapiens.GRCh38.pep.first1000lines.fa")
    gg = gpu_filter(gpu, enc, fasta, "protein", 7)
    go = oracle_filter(orc, fasta, "protein", 7)
    lut = enc.encode_lut("protein")
    # no reads at all
    assert len(gpu.score_reads(gg, np.zeros(0, np.uint8), np.zeros(1, np.uint64), lut, 0.5)) == 0
    # an empty read: the reference raises ZeroDivisionError (translate.py:83)
    flat, offsets = gpu.concat_reads(["ACGTACGTAGCTAGCTAGCATCGATCGATCAGCTAC", "", "ACGT"])
    with pytest.raises(ZeroDivisionError):
        gpu.score_reads(gg, flat, offsets, lut, 0.5)
    # long reads: beyond the packed fast path (1536) -> byte path; beyond ORPH_MAX_READ_LEN -> ValueError
    # (reads between 6144 and ORPH_MAX_READ_LEN: test_long_reads_block_kernel)
    rng = np.random.default_rng(3)
    long_reads = ["".join(rng.choice(list("ACGT"), size=n)) for n in (1536, 1537, 3000, 6144)]
    # make one of them free of stop codons in frame 1 so it is actually scored
    long_reads.append("GCTGAAAAGCTGGTT" * 200)
    flat, offsets = gpu.concat_reads(long_reads)
    res = gpu.score_reads(gg, flat, offsets, lut, 0.5)
    ora, _ = orc.score_reads(go, flat, offsets, orc.encode_lut("protein"), 0.5)
    assert_same_as_oracle(res, ora, "long")
    flat, offsets = gpu.concat_reads(["ACGT" * 49152])  # 196608 = ORPH_MAX_READ_LEN + 1
    with pytest.raises(ValueError):
        gpu.score_reads(gg, flat, offsets, lut, 0.5)
    # mismatched arguments fail loudly
    with pytest.raises(ValueError):
        gpu.score_reads(gg, flat, offsets, lut, float("nan"))


def test_emitted_peptides_on_gpu(gpu, orc, misc_kats):
    """orph_translate_frames: the un-encoded peptides the reference prints, incl. N / lower-case / IUPAC codons."""
    vec = load_score_vector("protein", 7)
    seqs = [r["seq"] for r in vec["reads"]] + list(misc_kats["six_frame"].keys())
    flat, offsets = gpu.concat_reads(seqs)
    item_read = np.repeat(np.arange(len(seqs)), 6)
    item_frame = np.tile(np.array([1, 2, 3, -1, -2, -3], dtype=np.int8), len(seqs))
    got = gpu.translate_frames(gpu.default_context(), flat, offsets, item_read, item_frame)
    for i, seq in enumerate(seqs):
        want = orc.six_frame_translation(seq)
        assert got[6 * i: 6 * i + 6] == [want[f] for f in (1, 2, 3, -1, -2, -3)], seq
    for seq, frames in misc_kats["six_frame"].items():  # produced by the unmodified reference
        i = seqs.index(seq)
        assert got[6 * i: 6 * i + 6] == [frames[str(f)] for f in (1, 2, 3, -1, -2, -3)]
    # a sub-batch view (offsets[0] != 0) and an empty item list
    sub = gpu.translate_frames(gpu.default_context(), flat, offsets[5:9], [0, 2], [2, -3])
    assert sub == [orc.six_frame_translation(seqs[5])[2], orc.six_frame_translation(seqs[7])[-3]]
    assert gpu.translate_frames(gpu.default_context(), flat, offsets, [], []) == []
    with pytest.raises(ValueError):
        gpu.translate_frames(gpu.default_context(), flat, offsets, [0], [4])


@pytest.mark.parametrize("ksize,n_tables,tablesize", [
    (7, 1, 1e6), (7, 2, 30011), (7, 7, 1e6), (5, 4, 1e5), (13, 3, 1e6), (8, 4, 1e6), (20, 4, 1e6), (32, 2, 1e6), (40, 4, 1e6),
])
def test_other_geometries_and_ksizes(gpu, enc, orc, refdata, ksize, n_tables, tablesize):
    """k-mer sizes with and without a compiled thread-per-frame kernel (generic warp kernel otherwise), 1 / 2 / 3 / 7
    tables (the non-unrolled probe loops), tiny over-full tables."""
    from orpheum_b200 import synth

    fasta = os.path.join(refdata, "index", "Homo_sapiens.GRCh38.pep.first1000lines.fa")
    res_, offs_ = fasta_arrays(fasta)
    lut = enc.encode_lut("dayhoff" if ksize >= 13 else "protein")
    olut = orc.encode_lut("dayhoff" if ksize >= 13 else "protein")
    go = orc.Nodegraph(ksize, tablesize, n_tables)
    go.add_peptides(res_, offs_, olut)
    gg = gpu.GpuNodegraph(ksize, tablesize, n_tables)
    gg.add_peptides(res_, offs_, lut)
    assert [gg.n_occupied(i) for i in range(n_tables)] == [go.n_occupied(i) for i in range(n_tables)]
    standard = np.zeros(256, dtype=bool)
    standard[np.frombuffer(synth.AMINO_ACIDS.encode(), dtype=np.uint8)] = True
    keep = [i for i in range(len(offs_) - 1) if offs_[i + 1] - offs_[i] >= 120 and standard[res_[int(offs_[i]):int(offs_[i + 1])]].all()]
    residues = np.concatenate([res_[int(offs_[i]):int(offs_[i + 1])] for i in keep])
    poffs = np.zeros(len(keep) + 1, dtype=np.uint64)
    poffs[1:] = np.cumsum([int(offs_[i + 1] - offs_[i]) for i in keep])
    flat, offsets = synth.make_reads_mixed(3000, residues, poffs, ksize=ksize, seed=11 + ksize, min_len=30, max_len=400)
    ora, n_empty = orc.score_reads(go, flat, offsets, olut, 0.5, n_threads=0)
    assert n_empty == 0
    res = gpu.score_reads(gg, flat, offsets, lut, 0.5)
    assert_same_as_oracle(res, ora, f"k={ksize} tables={n_tables}")


def test_table_larger_than_2_31_bits(gpu, enc, orc):
    """Tables of more than 2^31 bits take the 64-bit remainder path of the Barrett reduction."""
    from orpheum_b200 import synth

    residues, poffs = synth.make_proteome(n_proteins=300, seed=21)
    tablesize = 2_300_000_000
    gg = gpu.GpuNodegraph(7, tablesize, 1)
    assert gg.hashsizes()[0] > 2**31
    gg.add_peptides(residues, poffs, enc.encode_lut("protein"))
    go = orc.Nodegraph(7, tablesize, 1)
    go.add_peptides(residues, poffs, orc.encode_lut("protein"))
    assert gg.n_occupied(0) == go.n_occupied(0)
    flat, offsets = synth.concat_fixed(synth.make_reads_fixed(4000, residues, poffs, length=150, seed=22))
    ora, _ = orc.score_reads(go, flat, offsets, orc.encode_lut("protein"), 0.5, n_threads=0)
    res = gpu.score_reads(gg, flat, offsets, enc.encode_lut("protein"), 0.5)
    assert_same_as_oracle(res, ora, "2^31")
    hashes = np.random.default_rng(5).integers(0, 2**64, 5000, dtype=np.uint64)
    np.testing.assert_array_equal(gg.get_many(hashes), np.array([go.get(int(h)) for h in hashes], dtype=np.uint8))


def test_table_larger_than_2_32_bits(gpu, enc, orc):
    """SURVEY App. B lists tablesize 4e9: one table of more than 2^32 bits (550 MB) -- bins no longer fit 32 bits, byte offsets
    pass 2^29 -- built on the GPU, queried hash by hash and probed by every scoring kernel tier."""
    from orpheum_b200 import synth

    residues, poffs = synth.make_proteome(n_proteins=400, seed=23)
    tablesize = 4_400_000_000
    gg = gpu.GpuNodegraph(7, tablesize, 1)
    assert gg.hashsizes()[0] > 2**32
    gg.add_peptides(residues, poffs, enc.encode_lut("protein"))
    go = orc.Nodegraph(7, tablesize, 1)
    assert go.hashsizes()[0] == gg.hashsizes()[0]
    go.add_peptides(residues, poffs, orc.encode_lut("protein"))
    assert gg.n_occupied(0) == go.n_occupied(0)
    flat, offsets = synth.make_reads_mixed(6000, residues, poffs, ksize=7, seed=24, min_len=30, max_len=900)
    ora, _ = orc.score_reads(go, flat, offsets, orc.encode_lut("protein"), 0.5, n_threads=0)
    res = gpu.score_reads(gg, flat, offsets, enc.encode_lut("protein"), 0.5)
    assert_same_as_oracle(res, ora, "2^32")
    assert int((res["category"] == 3).sum()) > 500
    rng = np.random.default_rng(6)
    hashes = rng.integers(0, 2**64, 5000, dtype=np.uint64)
    np.testing.assert_array_equal(gg.get_many(hashes), np.array([go.get(int(h)) for h in hashes], dtype=np.uint8))


@pytest.mark.parametrize("reads_format", ["ascii", "packed"])
def test_reads_beyond_the_scoring_limit_are_classified_from_their_stop_codons(gpu, enc, orc, reads_format):
    """The reference has no read-length limit (translate.py:323-331).  Reads beyond ORPH_MAX_READ_LEN -- contigs, long
    transcripts -- are classified exactly when no frame needs scoring (stop codons in every frame, or a low-complexity
    read); only a longer read with a stop-free frame is refused, by name, with every other record filled in."""
    from orpheum_b200 import _lib, synth

    residues, poffs = synth.make_proteome(n_proteins=300, seed=31)
    lut, olut = enc.encode_lut("protein"), orc.encode_lut("protein")
    gg = gpu.GpuNodegraph(7, 1e6, 4)
    gg.add_peptides(residues, poffs, lut)
    go = orc.Nodegraph(7, int(1e6), 4)
    go.add_peptides(residues, poffs, olut)
    rng = np.random.default_rng(32)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    random_read = lambda n: acgt[rng.integers(0, 4, n)].tobytes().decode()
    low = list("A" * 250_000)
    for p in rng.integers(0, len(low), 2000):
        low[int(p)] = "C"
    reads = [random_read(150), random_read(_lib.MAX_READ_LEN + 1), random_read(300_000), "".join(low), random_read(7000),
             random_read(1_000_003), random_read(151)]
    if reads_format == "ascii":
        with_n = list(random_read(220_000))
        for p in rng.integers(0, len(with_n), 50):
            with_n[int(p)] = "N"
        reads += ["".join(with_n)]  # (a soft-masked read of that length translates to X only: no stop codon, see below)
    flat = np.frombuffer("".join(reads).encode(), dtype=np.uint8)
    offsets = np.zeros(len(reads) + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum([len(r) for r in reads])
    ora, _ = orc.score_reads(go, flat, offsets, olut, 0.5, n_threads=0)
    if reads_format == "ascii":
        res = gpu.score_reads(gg, flat, offsets, lut, 0.5)
    else:
        packed, needs = gpu.pack_reads(flat, offsets)
        assert not needs.any()
        res = gpu.score_reads(gg, packed, offsets, lut, 0.5, reads_format=_lib.READS_PACKED2)
    assert_same_as_oracle(res, ora, "over-long reads")
    assert (res["category"][3] == _lib.CAT_LOW_COMPLEXITY_NUCLEOTIDE).all() and (res["category"][2] == _lib.CAT_STOP_CODONS).all()
    # a stop-free frame of more than 65535 residues: refused, the other records are still there
    orf = "GCTGAA" * 40_000  # 240 kb, frame 1 = (AE)^40000
    if reads_format == "ascii":
        orf = random_read(200_000).lower()  # lower case translates to X: six stop-free frames
    reads2 = [random_read(150), orf, random_read(210_000)]
    flat2 = np.frombuffer("".join(reads2).encode(), dtype=np.uint8)
    offs2 = np.zeros(4, dtype=np.uint64)
    offs2[1:] = np.cumsum([len(r) for r in reads2])
    out = np.zeros(3, dtype=_lib.READ_RESULT_DTYPE)
    if reads_format == "ascii":
        buf, fmt = flat2, _lib.READS_ASCII
    else:
        buf, fmt = gpu.pack_reads(flat2, offs2)[0], _lib.READS_PACKED2
    rc = _lib.load().orph_score_reads(gg.ctx._h, gg._h, buf.ctypes.data, fmt, offs2.ctypes.data, 3, lut.ctypes.data, 0.5, out.ctypes.data)
    assert rc == _lib.ORPH_E_TOO_LONG
    assert (out["category"][1] == _lib.CAT_INVALID).all()
    ora2, _ = orc.score_reads(go, flat2, offs2, olut, 0.5, n_threads=0)
    assert_same_as_oracle(out[[0, 2]], ora2[[0, 2]], "records next to a refused read")


@pytest.mark.parametrize("tablesize,tiny", [(1_431_655_766, True), (1_431_656_200, False)])
def test_tables_at_the_32_bit_reduction_limit(gpu, enc, orc, tablesize, tiny):
    """Four tables just below / across (2^32 - 1) / 3 bits: the largest geometry whose bins come from mod_prime_tiny
    (device_common.cuh: the un-corrected remainder must stay below 3p < 2^32) and the smallest that must not."""
    from orpheum_b200 import synth

    limit = 0xFFFFFFFF // 3
    residues, poffs = synth.make_proteome(n_proteins=300, seed=27)
    lut, olut = enc.encode_lut("protein"), orc.encode_lut("protein")
    gg = gpu.GpuNodegraph(7, tablesize, 4)
    sizes = gg.hashsizes()
    assert all(s <= limit for s in sizes) == tiny and max(sizes) > limit - 400
    gg.add_peptides(residues, poffs, lut)
    go = orc.Nodegraph(7, tablesize, 4)
    assert list(go.hashsizes()) == list(sizes)
    go.add_peptides(residues, poffs, olut)
    assert [gg.n_occupied(i) for i in range(4)] == [go.n_occupied(i) for i in range(4)]
    flat, offsets = synth.make_reads_mixed(5000, residues, poffs, ksize=7, seed=28, min_len=30, max_len=900)
    ora, _ = orc.score_reads(go, flat, offsets, olut, 0.5, n_threads=0)
    res = gpu.score_reads(gg, flat, offsets, lut, 0.5)
    assert_same_as_oracle(res, ora, f"tablesize {tablesize}")
    assert int((res["category"] == 3).sum()) > 300
    rng = np.random.default_rng(7)
    hashes = np.concatenate([rng.integers(0, 2**64, 5000, dtype=np.uint64),
                             np.array([0, 1, 2**32 - 1, 2**32, 2**64 - 1, 2**64 - 2] + [s * q + d for s in sizes for q in (1, 2**32 // s, 2**64 // s - 1)
                                                                                 for d in (0, s - 1)], dtype=np.uint64)])
    np.testing.assert_array_equal(gg.get_many(hashes), np.array([go.get(int(h)) for h in hashes], dtype=np.uint8))


def test_scoring_against_hbm_resident_filter(gpu, enc, orc):
    """BASELINE cfg4's filter geometry (tablesize 1e9 x 4 = 500 MB, four times the L2): 1e5 reads scored on the GPU against
    the oracle -- the probes of every table go to HBM, the persisting-L2 window covers only part of the buffer."""
    from orpheum_b200 import _lib, synth

    residues, poffs = synth.make_proteome(n_proteins=4000, seed=25)
    lut, olut = enc.encode_lut("protein"), orc.encode_lut("protein")
    gg = gpu.GpuNodegraph(7, int(1e9), 4)
    gg.add_peptides(residues, poffs, lut)
    go = orc.Nodegraph(7, int(1e9), 4)
    go.add_peptides(residues, poffs, olut)
    assert [gg.n_occupied(i) for i in range(4)] == [go.n_occupied(i) for i in range(4)]
    flat, offsets = synth.concat_fixed(synth.make_reads_fixed(100_000, residues, poffs, length=150, seed=26))
    ora, _ = orc.score_reads(go, flat, offsets, olut, 0.5, n_threads=0)
    res = gpu.score_reads(gg, flat, offsets, lut, 0.5)
    assert_same_as_oracle(res, ora, "hbm-resident filter")
    packed, needs = gpu.pack_reads(flat, offsets)
    assert not needs.any()
    res2 = gpu.score_reads(gg, packed, offsets, lut, 0.5, reads_format=_lib.READS_PACKED2)
    assert res2.tobytes() == res.tobytes()


def _orf_read(rng, residues, poffs, n_codons, subst=0.0):
    """A read with one long open frame: back-translated proteome residues (one fixed codon per amino acid)."""
    codon = {"A": "GCT", "R": "CGT", "N": "AAC", "D": "GAC", "C": "TGC", "Q": "CAG", "E": "GAA", "G": "GGT", "H": "CAC",
             "I": "ATC", "L": "CTG", "K": "AAA", "M": "ATG", "F": "TTC", "P": "CCG", "S": "TCT", "T": "ACC", "W": "TGG",
             "Y": "TAC", "V": "GTT"}
    out = []
    while len(out) < n_codons:
        p = int(rng.integers(0, len(poffs) - 1))
        seq = residues[int(poffs[p]):int(poffs[p + 1])].tobytes().decode()
        out.extend(codon.get(a, "GCT") for a in seq)
    s = np.frombuffer("".join(out[:n_codons]).encode(), dtype=np.uint8).copy()
    if subst:
        hit = rng.random(s.size) < subst
        s[hit] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(hit.sum()))]
    return s.tobytes().decode()


@pytest.mark.parametrize("reads_format", ["ascii", "packed"])
@pytest.mark.parametrize("alphabet,ksize", [("protein", 7), ("dayhoff", 12), ("hp", 31)])
def test_long_reads_block_kernel(gpu, enc, orc, refdata, alphabet, ksize, reads_format):
    """Reads of 6145 .. ORPH_MAX_READ_LEN bases (the reference has no length limit) take the block-per-read kernel:
    long open reading frames with thousands of distinct and repeated k-mers, random reads, N bytes, lower case, a
    low-complexity read, the maximum length, mixed with short reads in one batch."""
    from orpheum_b200 import _lib, synth

    residues, poffs = synth.make_proteome(n_proteins=400, seed=31)
    lut, olut = enc.encode_lut(alphabet), orc.encode_lut(alphabet)
    gg = gpu.GpuNodegraph(ksize, 2e6, 4)
    gg.add_peptides(residues, poffs, lut)
    go = orc.Nodegraph(ksize, 2e6, 4)
    go.add_peptides(residues, poffs, olut)
    rng = np.random.default_rng(17)
    rand = lambda n: "".join(rng.choice(list("ACGT"), size=n))
    reads = [
        _orf_read(rng, residues, poffs, 2049),            # 6147 nt, open frame 1, every k-mer in the filter
        _orf_read(rng, residues, poffs, 3000, 0.002),     # a few substitutions: some frames may close
        rand(6145), rand(20000),                          # random: stops everywhere
        "ACGTAC" + _orf_read(rng, residues, poffs, 7000), # open frame 1 again (6 = 0 mod 3), 21 kb
        "G" + _orf_read(rng, residues, poffs, 5000),      # open frame 2
        _orf_read(rng, residues, poffs, 1200) * 3,        # tandem repeat of a 3.6 kb unit: 2/3 duplicates
        "GCTGAAAAGCTGGTT" * 1000,                         # 15 kb of a 5-residue repeat: low-complexity peptide
        "A" * 5000 + "C" * 5000,                          # fastp low complexity
        rand(150), _orf_read(rng, residues, poffs, 50),   # short reads in the same batch
        _orf_read(rng, residues, poffs, 65535) + "AC",    # ORPH_MAX_READ_LEN = 196607 bases, 65535-residue frames
    ]
    if reads_format == "ascii":
        with_n = list(_orf_read(rng, residues, poffs, 4000))
        for pos in rng.integers(0, len(with_n), 40):
            with_n[int(pos)] = "N"
        reads.append("".join(with_n))                     # N inside codons: X residues and xxN codons
        reads.append(_orf_read(rng, residues, poffs, 2500).lower())
        reads.append(_orf_read(rng, residues, poffs, 2500)[:7000] + "acgtnRYK" * 10)
    flat, offsets = gpu.concat_reads(reads)
    ora, _ = orc.score_reads(go, flat, offsets, olut, 0.5, n_threads=0)
    assert (ora["n_kmers"] > 50000).any() and (ora["n_kmers"] > 2000).sum() >= 5  # the long frames were really scored
    if reads_format == "ascii":
        res = gpu.score_reads(gg, flat, offsets, lut, 0.5)
    else:
        packed, needs = gpu.pack_reads(flat, offsets)
        assert not needs.any()
        res = gpu.score_reads(gg, packed, offsets, lut, 0.5, reads_format=_lib.READS_PACKED2)
    assert_same_as_oracle(res, ora, f"long {alphabet} {reads_format}")
    # the exact all-pairs restatement of the de-duplication agrees (without the 196 kb read: O(n^2))
    if reads_format == "ascii":
        flat2, offsets2 = gpu.concat_reads(reads[:4])
        gg.ctx.set_force_exact_dedup(True)
        try:
            res2 = gpu.score_reads(gg, flat2, offsets2, lut, 0.5)
        finally:
            gg.ctx.set_force_exact_dedup(False)
        assert res2.tobytes() == res[:4].tobytes()


def test_full_size_properties_cfg2(gpu, enc, orc):
    """BASELINE configs[1] at full size (10 M x 150 nt, protein k=7, tablesize 1e8 x 4) through size-independent
    properties: (1) every input form and kernel path gives byte-identical records (host ASCII, host 2-bit, warp-per-read
    instead of thread-per-frame on a 1 M slice); (2) scoring uneven sub-batches and concatenating equals scoring the
    whole batch; (3) reverse-complement symmetry: frame f of rc(read) is frame f+3 of the read; (4) a random sample
    drawn from the whole batch equals the oracle."""
    from orpheum_b200 import _lib, synth

    n = 10_000_000
    residues, poffs = synth.make_proteome(n_proteins=20000, seed=1001)
    reads2d = synth.make_reads_fixed(n, residues, poffs, length=150, seed=2002)
    flat, offsets = synth.concat_fixed(reads2d)
    lut = enc.encode_lut("protein")
    gg = gpu.GpuNodegraph(7, 1e8, 4)
    gg.add_peptides(residues, poffs, lut, count_unique=False)
    res = gpu.score_reads(gg, flat, offsets, lut, 0.5)
    # the batch is not degenerate: every category of an ACGT read occurs, and many frames were scored
    cats = np.bincount(res["category"].reshape(-1), minlength=6)
    assert cats[_lib.CAT_CODING] > 4_000_000 and cats[_lib.CAT_NON_CODING] > 1_000_000 and cats[0] > 30_000_000
    digest = lambda a: (int(a.view(np.uint64).sum(dtype=np.uint64)), int(np.bitwise_xor.reduce(a.view(np.uint64))))
    want = digest(res)
    # (1) 2-bit host input
    packed, needs = gpu.pack_reads(flat, offsets)
    assert not needs.any()
    assert digest(gpu.score_reads(gg, packed, offsets, lut, 0.5, reads_format=_lib.READS_PACKED2)) == want
    # (1) warp-per-read kernel on a slice
    lo, hi = 3_000_000, 4_000_000
    gg.ctx.set_thread_path(False)
    try:
        sl = gpu.score_reads(gg, flat, offsets[lo:hi + 1], lut, 0.5)
    finally:
        gg.ctx.set_thread_path(True)
    assert sl.tobytes() == res[lo:hi].tobytes()
    # (2) uneven sub-batches
    cuts = [0, 1, 31, 32, 1000, 1_048_577, 2_500_001, 7_000_003, n]
    for a, b in zip(cuts[:-1], cuts[1:]):
        assert gpu.score_reads(gg, flat, offsets[a:b + 1], lut, 0.5).tobytes() == res[a:b].tobytes(), (a, b)
    # (3) reverse complement: forward frames of rc(read) are the read's reverse frames and vice versa
    m = 2_000_000
    comp = np.zeros(256, dtype=np.uint8)
    comp[np.frombuffer(b"ACGT", dtype=np.uint8)] = np.frombuffer(b"TGCA", dtype=np.uint8)
    rc = np.ascontiguousarray(comp[reads2d[:m, ::-1]])
    rres = gpu.score_reads(gg, rc.reshape(-1), offsets[:m + 1], lut, 0.5)
    swap = [3, 4, 5, 0, 1, 2]
    for field in ("n_kmers", "n_hits", "category"):
        np.testing.assert_array_equal(rres[field], res[field][:m][:, swap], err_msg=field)
    np.testing.assert_array_equal(rres["flags"] & 0x40, res["flags"][:m] & 0x40)
    # (4) random sample against the oracle
    pick = np.sort(np.random.default_rng(9).choice(n, 50_000, replace=False))
    sflat, soffs = synth.concat_fixed(reads2d[pick])
    go = orc.Nodegraph(7, int(1e8), 4)
    go.add_peptides(residues, poffs, orc.encode_lut("protein"))
    ora, _ = orc.score_reads(go, sflat, soffs, orc.encode_lut("protein"), 0.5, n_threads=0)
    assert_same_as_oracle(res[pick], ora, "full-size sample")


def test_host_packing_ingest_paths(gpu, enc, orc):
    """orph_score_reads on large ASCII batches: chunks packed to 2 bits on the host threads, chunks sent as ASCII,
    the second pass over unpackable reads (N, lower case), the mostly-unpackable fallback, pageable and pinned input --
    every route must give the records of the plain ASCII path (host_threads = 0) and of the oracle."""
    import torch

    from orpheum_b200 import _lib, synth

    residues, poffs = synth.make_proteome(n_proteins=2000, seed=41)
    lut, olut = enc.encode_lut("protein"), orc.encode_lut("protein")
    gg = gpu.GpuNodegraph(7, 1e7, 4)
    gg.add_peptides(residues, poffs, lut)
    go = orc.Nodegraph(7, int(1e7), 4)
    go.add_peptides(residues, poffs, olut)
    ctx = gg.ctx
    # 2.6 M mixed reads (~450 MB of ASCII, 2 % with N, several chunks): every category, three kernel tiers, byte path
    flat, offsets = synth.make_reads_mixed(2_600_000, residues, poffs, ksize=7, seed=77)
    flat = flat.copy()
    lower = np.random.default_rng(5).choice(len(offsets) - 1, 3000, replace=False)  # plus some soft-masked reads
    for r in lower:
        a, b = int(offsets[r]), int(offsets[r + 1])
        flat[a:b] = np.frombuffer(flat[a:b].tobytes().lower(), dtype=np.uint8)
    try:
        ctx.set_host_threads(0)
        start = ctx.ingest_stats()
        want = gpu.score_reads(gg, flat, offsets, lut, 0.5)
        before = ctx.ingest_stats()
        assert before["chunks_packed_on_host"] == start["chunks_packed_on_host"]  # nothing is packed with host_threads = 0
        pinned = torch.empty(flat.size, dtype=torch.uint8).pin_memory()
        pinned.numpy()[:] = flat
        for threads, buf in ((1, flat), (4, flat), (-1, flat), (-1, pinned.numpy()), (3, pinned.numpy())):
            ctx.set_host_threads(threads)
            got = gpu.score_reads(gg, buf, offsets, lut, 0.5)
            assert got.tobytes() == want.tobytes(), threads
        after = ctx.ingest_stats()
        assert after["chunks_packed_on_host"] - before["chunks_packed_on_host"] >= 6 and after["pack_bytes_per_s"] > 0
        # sub-range with a non-zero first offset and a non-multiple-of-32 start
        ctx.set_host_threads(-1)
        a, b = 12_345, 2_222_222
        assert gpu.score_reads(gg, flat, offsets[a:b + 1], lut, 0.5).tobytes() == want[a:b].tobytes()
        # an empty read / an over-long read inside a large host-packed batch: same behaviour as on the small-batch route
        cut = 1_000_000
        offs_empty = np.concatenate([offsets[:cut + 1], offsets[cut:cut + 200_001]])  # read `cut` has length 0
        with pytest.raises(ZeroDivisionError):
            gpu.score_reads(gg, flat, offs_empty, lut, 0.5)
        offs_long = np.concatenate([offsets[:cut], offsets[cut + 2000:cut + 100_000]])  # one read of ~350 k bases
        assert int(np.diff(offs_long.astype(np.int64)).max()) > 196_607
        got_long = gpu.score_reads(gg, flat, offs_long, lut, 0.5)  # beyond ORPH_MAX_READ_LEN, stop codons in every frame: classified
        assert (got_long["category"][cut - 1] == _lib.CAT_STOP_CODONS).all() and got_long[:cut - 1].tobytes() == want[:cut - 1].tobytes()
        assert got_long[cut:].tobytes() == want[cut + 2000:cut + 99_999].tobytes()
        assert gpu.score_reads(gg, flat, offsets[:cut + 1], lut, 0.5).tobytes() == want[:cut].tobytes()  # the context recovers
        # mostly unpackable input: soft-masked everything
        low = np.frombuffer(flat.tobytes().lower(), dtype=np.uint8)
        ctx.set_host_threads(0)
        want_low = gpu.score_reads(gg, low, offsets, lut, 0.5)
        ctx.set_host_threads(-1)
        assert gpu.score_reads(gg, low, offsets, lut, 0.5).tobytes() == want_low.tobytes()
    finally:
        ctx.set_host_threads(-1)
    sample = np.sort(np.random.default_rng(6).choice(len(offsets) - 1, 30_000, replace=False))
    sample = np.union1d(sample, lower[:500])
    sflat, soffs = gpu.concat_reads([flat[int(offsets[r]):int(offsets[r + 1])].tobytes() for r in sample])
    ora, _ = orc.score_reads(go, sflat, soffs, olut, 0.5, n_threads=0)
    assert_same_as_oracle(want[sample], ora, "ingest sample")


def test_full_size_cfg4_index_build(gpu, enc, orc):
    """BASELINE configs[3], index side: Bloom build from the 1e8-residue synthetic proteome (P100M) on the GPU at
    tablesize 1e8 (L2 regime, ~63 % fill) and 1e9 (HBM regime).  Bit-exact against the oracle on a 1e6-residue prefix
    proteome with the same geometry; at full size through properties: the build is idempotent and order-independent
    (two halves in either order, with and without the n_unique_kmers bookkeeping), every bit of the prefix filter is set
    in the full one, and every k-mer of sampled proteins is found (no false negatives)."""
    from orpheum_b200 import synth

    residues, poffs = synth.make_proteome(n_proteins=187000, seed=1004)
    assert 9.5e7 < int(poffs[-1]) < 1.1e8
    lut, olut = enc.encode_lut("protein"), orc.encode_lut("protein")
    n_prefix = int(np.searchsorted(poffs, 1_000_000))
    half = len(poffs) // 2
    for tablesize in (int(1e8), int(1e9)):
        full = gpu.GpuNodegraph(7, tablesize, 4)
        full.add_peptides(residues, poffs, lut)
        pops = [full.n_occupied(i) for i in range(4)]
        # order independence + idempotence
        other = gpu.GpuNodegraph(7, tablesize, 4)
        other.add_peptides(residues, poffs[half:], lut, count_unique=False)
        other.add_peptides(residues, poffs[: half + 1], lut, count_unique=False)
        other.add_peptides(residues, poffs[: n_prefix + 1], lut)
        assert [other.n_occupied(i) for i in range(4)] == pops
        ta, _ = full.tables()
        tb, _ = other.tables()
        assert all((a == b).all() for a, b in zip(ta, tb))
        del other, tb
        # prefix proteome: identical to the oracle, and a subset of the full filter
        small = gpu.GpuNodegraph(7, tablesize, 4)
        small.add_peptides(residues, poffs[: n_prefix + 1], lut)
        o_small = orc.Nodegraph(7, tablesize, 4)
        o_small.add_peptides(residues, poffs[: n_prefix + 1], olut)
        ts, _ = small.tables()
        for i in range(4):
            assert (ts[i] == o_small.table(i)).all(), (tablesize, i)
            assert ((ts[i] & ta[i]) == ts[i]).all()
        assert abs(small.n_unique_kmers() - o_small.n_unique_kmers()) <= 1e-3 * o_small.n_unique_kmers()  # tests/test_index.py:54 rtol
        # no false negatives on proteins sampled from the whole proteome
        rng = np.random.default_rng(tablesize % 1000)
        kmers = []
        for p in rng.choice(len(poffs) - 1, 300, replace=False):
            seq = residues[int(poffs[p]):int(poffs[p + 1])].tobytes().decode()
            kmers.extend(seq[j:j + 7] for j in range(0, len(seq) - 6, 5))
        assert full.get_many(full.ctx.hash_kmers(kmers)).all()
        del full, small, ta, ts, o_small


def test_counter_based_read_generator(gpu, enc, orc):
    """The HBM read generator of bench.py --workload cfg4 (orph_synth_reads_device) and its numpy twin
    (synth.counter_reads) produce the same bases for the same global indices, for any split of the index range, and the
    generated batch scores identically to the host copy."""
    import torch

    from orpheum_b200 import _lib, synth

    residues, poffs = synth.make_proteome(n_proteins=3000, seed=8)
    ctx = gpu.default_context()
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    d_res = torch.from_numpy(residues).to(dev)
    d_poffs = torch.from_numpy(poffs.view(np.int64)).to(dev)
    d_syn = torch.from_numpy(synth.syn_table().view(np.int64)).to(dev)
    for length, first, n in ((150, 0, 50_000), (150, 999_999_999_000, 4097), (149, 12_345, 1000), (64, 7, 333), (320, 1 << 40, 200)):
        d_reads = torch.zeros((n * length + 31) // 32 + 2, dtype=torch.int64, device=dev)
        _lib.check(lib.orph_synth_reads_device(ctx._h, 3003, first, n, length, d_res.data_ptr(), d_poffs.data_ptr(), len(poffs) - 1,
                                               d_syn.data_ptr(), d_reads.data_ptr()))
        ctx.synchronize()
        torch.cuda.synchronize()
        host = synth.counter_reads(np.arange(first, first + n, dtype=np.uint64), residues, poffs, seed=3003, length=length)
        flat, offs = synth.concat_fixed(host)
        want, needs = gpu.pack_reads(flat, offs)
        assert not needs.any()
        got = d_reads.cpu().numpy().view(np.uint64)
        n_words = (n * length + 31) // 32
        np.testing.assert_array_equal(got[:n_words], want[:n_words], err_msg=f"L={length} first={first}")
    # half of the reads are coding: the generated batch is a meaningful workload
    lut = enc.encode_lut("protein")
    gg = gpu.GpuNodegraph(7, 1e7, 4)
    gg.add_peptides(residues, poffs, lut)
    host = synth.counter_reads(np.arange(0, 20_000, dtype=np.uint64), residues, poffs, seed=3003, length=150)
    flat, offs = synth.concat_fixed(host)
    res = gpu.score_reads(gg, flat, offs, lut, 0.5)
    coding = (res["category"] == _lib.CAT_CODING).any(axis=1)
    assert coding[0::2].mean() > 0.9 and coding[1::2].mean() < 0.05
    go = orc.Nodegraph(7, int(1e7), 4)
    go.add_peptides(residues, poffs, orc.encode_lut("protein"))
    ora, _ = orc.score_reads(go, flat, offs, orc.encode_lut("protein"), 0.5, n_threads=0)
    assert_same_as_oracle(res, ora, "counter reads")


def test_host_packer_on_the_gpu_box(gpu):
    """The SIMD level is chosen at run time (AVX2 on the GPU boxes, SSSE3 in the build container): run the numpy
    comparison of the host packer on this host too."""
    from tests.test_cabi_symbols import test_host_packer_simd_against_numpy

    test_host_packer_simd_against_numpy(None)


def test_ksize_one_and_two(gpu, enc, orc):
    """Degenerate k-mer sizes through every kernel tier (no thread-per-frame kernel is compiled for them): k = 1 makes a
    2048-residue frame reach position 2047 of the warp kernel's de-duplication slots, which is why reads of 6142+ bases
    go to the block-per-read kernel there."""
    from orpheum_b200 import synth

    residues, poffs = synth.make_proteome(n_proteins=50, seed=61)
    rng = np.random.default_rng(62)
    reads = [_orf_read(rng, residues, poffs, n) for n in (50, 107, 400, 2046, 2047, 2048, 2049, 3000)]
    reads += ["".join(rng.choice(list("ACGT"), size=n)) for n in (30, 150, 6141, 6142, 6144)]
    flat, offsets = gpu.concat_reads(reads)
    for ksize in (1, 2):
        gg = gpu.GpuNodegraph(ksize, 1e4, 4)
        gg.add_peptides(residues[:200], np.array([0, 200], dtype=np.uint64), enc.encode_lut("protein"))  # part of the alphabet only
        go = orc.Nodegraph(ksize, int(1e4), 4)
        go.add_peptides(residues[:200], np.array([0, 200], dtype=np.uint64), orc.encode_lut("protein"))
        ora, _ = orc.score_reads(go, flat, offsets, orc.encode_lut("protein"), 0.5, n_threads=0)
        res = gpu.score_reads(gg, flat, offsets, enc.encode_lut("protein"), 0.5)
        assert_same_as_oracle(res, ora, f"k={ksize}")


def test_two_host_threads_two_contexts(gpu, enc, orc):
    """The ABI is thread-compatible per orph_ctx: two host threads, each with its own context and filter replica on the
    same GPU, score different batches at the same time (ctypes releases the GIL) and get the serial results."""
    import threading

    from orpheum_b200 import synth

    residues, poffs = synth.make_proteome(n_proteins=1000, seed=71)
    lut = enc.encode_lut("protein")
    base = gpu.GpuNodegraph(7, 1e7, 4)
    base.add_peptides(residues, poffs, lut)
    batches = [synth.make_reads_mixed(300_000, residues, poffs, ksize=7, seed=72 + i) for i in range(2)]
    want = [gpu.score_reads(base, f, o, lut, 0.5) for f, o in batches]
    ctxs = [gpu.Context(0), gpu.Context(0)]
    graphs = [base.replicate(c) for c in ctxs]
    errors, got = [], [None, None]

    def work(i):
        try:
            for _ in range(6):
                res = gpu.score_reads(graphs[i], batches[i][0], batches[i][1], lut, 0.5, ctx=ctxs[i])
                if res.tobytes() != want[i].tobytes():
                    raise AssertionError(f"thread {i}: records differ")
            got[i] = True
        except BaseException as exc:  # noqa: BLE001 - reported in the main thread
            errors.append(exc)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert got == [True, True]
    del graphs
    for c in ctxs:
        c.close()


def test_threads_with_different_batch_geometries(gpu, enc, orc):
    """Two host threads whose batches need DIFFERENT amounts of dynamic shared memory from the same kernels (reads of up to
    120 nt with N in one thread, of up to 320 and 1500 nt with N in the other), many small calls each: a kernel's
    shared-memory allowance is a property of the kernel, not of the launch, so a thread that lowers it to its own size
    between another thread's set and launch makes that launch fail (found as a 1-in-25 flake of the two-streams test)."""
    import threading

    from orpheum_b200 import synth

    residues, poffs = synth.make_proteome(n_proteins=300, seed=81)
    lut, olut = enc.encode_lut("protein"), orc.encode_lut("protein")
    base = gpu.GpuNodegraph(7, 1e6, 4)
    base.add_peptides(residues, poffs, lut)
    go = orc.Nodegraph(7, int(1e6), 4)
    go.add_peptides(residues, poffs, olut)
    rng = np.random.default_rng(82)

    def batch(max_len, seed):
        flat, offsets = synth.make_reads_mixed(400, residues, poffs, ksize=7, seed=seed, min_len=40, max_len=max_len)
        flat = flat.copy()
        for r in range(0, 400, 3):  # every third read holds an N
            flat[int(offsets[r]) + int(rng.integers(0, int(offsets[r + 1] - offsets[r])))] = ord("N")
        return flat, offsets

    batches = [[batch(120, 100 + j) for j in range(4)], [batch(320 if j % 2 else 1500, 200 + j) for j in range(4)]]
    want = [[orc.score_reads(go, f, o, olut, 0.5, n_threads=0)[0] for f, o in bs] for bs in batches]
    ctxs = [gpu.Context(0), gpu.Context(0)]
    graphs = [base.replicate(c) for c in ctxs]
    errors = []

    def work(i):
        try:
            for it in range(150):
                f, o = batches[i][it % 4]
                res = gpu.score_reads(graphs[i], f, o, lut, 0.5, ctx=ctxs[i])
                assert_same_as_oracle(res, want[i][it % 4], f"thread {i} call {it}")
        except BaseException as exc:  # noqa: BLE001 - reported in the main thread
            errors.append(exc)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    del graphs
    for c in ctxs:
        c.close()


def test_context_and_filter_lifecycle_churn(gpu, enc, orc):
    """Contexts and filters created and destroyed over and over on recycled device memory (nothing the library reads may
    depend on cudaMalloc returning zeroed memory): every fresh context scores correctly on its first call, small and
    large batches alternate so that every staging buffer is re-allocated."""
    import torch

    from orpheum_b200 import synth

    residues, poffs = synth.make_proteome(n_proteins=300, seed=81)
    lut, olut = enc.encode_lut("protein"), orc.encode_lut("protein")
    go = orc.Nodegraph(7, int(1e6), 4)
    go.add_peptides(residues, poffs, olut)
    small = synth.make_reads_mixed(2000, residues, poffs, ksize=7, seed=82)
    large = synth.make_reads_mixed(120_000, residues, poffs, ksize=7, seed=83)
    want = {}
    for name, (flat, offs) in (("small", small), ("large", large)):
        ora, _ = orc.score_reads(go, flat, offs, olut, 0.5, n_threads=0)
        want[name] = ora
    for it in range(12):
        # dirty the allocator's free lists: a buffer full of ones, released right before the context is created
        junk = torch.full(((64 + 16 * it) << 20,), 0xFF, dtype=torch.uint8, device="cuda:0")
        torch.cuda.synchronize()
        del junk
        torch.cuda.empty_cache()
        ctx = gpu.Context(0)
        g = gpu.GpuNodegraph(7, 1e6, 4, ctx=ctx)
        g.add_peptides(residues, poffs, lut)
        order = ("small", "large") if it % 2 else ("large", "small")
        for name in order:
            flat, offs = small if name == "small" else large
            res = gpu.score_reads(g, flat, offs, lut, 0.5, ctx=ctx)
            assert_same_as_oracle(res[:3000], want[name][:3000], f"churn {it} {name}")
            scored = want[name]["n_kmers"] >= 0
            assert (res["n_hits"][scored] == want[name]["n_hits"][scored]).all()
        ctx.check()
        del g
        ctx.close()


def test_randomised_differential_smoke():
    """A short run of tests/scripts/fuzz_parity.py (random alphabets, k, geometries, read-length tiers, dirty bytes, input
    formats, kernel routes; GPU records, filter tables and emitted peptides against the oracle) with a fresh seed per day of
    the year, so that the driver's runs keep sampling new cases.  The seed is printed and is part of every failure message;
    ORPHEUM_B200_FUZZ_SEED=<seed> re-runs exactly that case list.  Longer recorded runs: profiles/r01_fuzz_parity_summary.json."""
    import datetime
    import json
    import subprocess
    import sys

    seed = int(os.environ.get("ORPHEUM_B200_FUZZ_SEED", 1000 + datetime.date.today().timetuple().tm_yday))
    print(f"fuzz seed {seed} (re-run with ORPHEUM_B200_FUZZ_SEED={seed})")
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "scripts", "fuzz_parity.py")
    proc = subprocess.run([sys.executable, script, "20", str(seed)], capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, f"fuzz seed {seed} (ORPHEUM_B200_FUZZ_SEED={seed} reproduces it)\n" + proc.stdout[-3000:] + proc.stderr[-3000:]
    summary = json.loads(proc.stdout.strip().splitlines()[-1])
    assert summary["differences"] == 0 and summary["fuzz_rounds"] > 20, f"fuzz seed {seed}: {summary}"
