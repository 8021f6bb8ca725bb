"""The CPU oracle (oracle/orpheum_oracle.c) against every golden vector the reference's own
tests hold for the hot path, and against vectors produced by the unmodified reference
(oracle/gen_golden.py).  CPU only."""
import hashlib
import math
import os

import numpy as np
import pandas as pd
import pytest

from oracle import oracle as orc
from tests.conftest import SCORE_VECTORS, load_score_vector, read_fastx

CATEGORY_TEXT = {
    orc.CAT_STOP: "Translation frame has stop codon(s)",
    orc.CAT_SHORT_PEPTIDE: "Translation is shorter than peptide k-mer size + 1",
    orc.CAT_CODING: "Coding",
    orc.CAT_NONCODING: "Non-coding",
    orc.CAT_LOWCPLX_NT: "Read length was shorter than 3 * peptide k-mer size",
}


def category_text(code, misc_kats, alphabet):
    if code == orc.CAT_LOWCPLX_PEPTIDE:
        return misc_kats["low_complexity_categories"][alphabet]
    return CATEGORY_TEXT[code]


def build_filter(fasta, alphabet, ksize, tablesize=1e6, n_tables=4):
    g = orc.Nodegraph(ksize, tablesize, n_tables)
    lut = orc.encode_lut(alphabet)
    for _, seq in read_fastx(fasta):
        g.add_peptide(seq, lut)
    return g


def test_hash_murmur_kats(misc_kats):
    # SURVEY App. C + sourmash's own unit-test value for "ACG"
    assert orc.hash_murmur("ACG") == 1731421407650554201
    for key, value in misc_kats["hash_murmur"].items():
        assert orc.hash_murmur(key) == value, key


def test_table_sizes(misc_kats):
    for x, primes in misc_kats["table_sizes"].items():
        assert orc.table_sizes(int(x), 4) == primes
    assert orc.table_sizes(1e6, 4) == [999983, 999979, 999961, 999959]


def test_codons_and_six_frames(misc_kats):
    # tests/test_translate_single_seq.py:23-104
    for codon, aa in misc_kats["codons"].items():
        assert orc.translate_codon(codon) == aa, codon
    assert orc.translate_codon("AC") == ""
    assert orc.reverse_complement("AACCGGTT") == "AACCGGTT"
    assert orc.reverse_complement("AACC") == "GGTT"
    for seq, frames in misc_kats["six_frame"].items():
        got = orc.six_frame_translation(seq)
        assert {str(k): v for k, v in got.items()} == frames
        assert list(got) == [1, 2, 3, -1, -2, -3]


def test_alphabets(misc_kats):
    # tests/test_sequence_encodings.py:81-155 (through the unmodified encode_peptide)
    probe = "ACDEFGHIKLMNPQRSTVWYXUBZOJ*acdxy"
    for name, encoded in misc_kats["alphabets"].items():
        lut = orc.encode_lut(name)
        assert bytes(lut[np.frombuffer(probe.encode(), dtype=np.uint8)]).decode() == encoded, name
    with pytest.raises(ValueError):
        orc.encode_lut("not-an-alphabet")


def test_complexities(misc_kats):
    # tests/test_translate.py:83-124
    for seq, value in misc_kats["fastp_complexity"].items():
        assert orc.fastp_complexity(seq) == value
    assert orc.count_distinct_kmers("GATTACA", 3) == 5  # tests/test_compare_kmer_content.py:52-57


@pytest.mark.parametrize("alphabet,ksize", [("protein", 7), ("dayhoff", 12)])
def test_rebuild_golden_nodegraph_byte_identical(refdata, tmp_path, alphabet, ksize):
    golden = os.path.join(
        refdata, "index", f"Homo_sapiens.GRCh38.pep.subset.alphabet-{alphabet}_ksize-{ksize}.bloomfilter.nodegraph"
    )
    g = build_filter(os.path.join(refdata, "index", "Homo_sapiens.GRCh38.pep.subset.fa.gz"), alphabet, ksize)
    out = tmp_path / "rebuilt.nodegraph"
    g.save(str(out))
    assert out.read_bytes() == open(golden, "rb").read()
    # tests/test_index.py:32-50 (rtol is the reference's own)
    expected = {("protein", 7): 506352, ("dayhoff", 12): 488469}[(alphabet, ksize)]
    np.testing.assert_allclose(g.n_unique_kmers(), expected, rtol=1e-3)
    # load path round trip
    again = orc.Nodegraph.load(golden)
    assert again.ksize() == ksize and again.hashsizes() == [999983, 999979, 999961, 999959]
    assert again.n_occupied() == g.n_occupied()
    with pytest.raises(OSError):
        orc.Nodegraph.load(os.path.join(refdata, "index", "Homo_sapiens.GRCh38.pep.first1000lines.fa"))


def test_index_kats_first1000lines(refdata, index_kats, tmp_path):
    fasta = os.path.join(refdata, "index", "Homo_sapiens.GRCh38.pep.first1000lines.fa")
    for kat in index_kats:
        g = build_filter(fasta, kat["alphabet"], kat["ksize"], kat["tablesize"], kat["n_tables"])
        assert [g.n_occupied(i) for i in range(kat["n_tables"])] == kat["popcounts"], kat
        out = tmp_path / "k.nodegraph"
        g.save(str(out))
        assert hashlib.sha256(out.read_bytes()).hexdigest() == kat["sha256"], kat
        np.testing.assert_allclose(g.n_unique_kmers(), kat["n_unique_kmers"], rtol=1e-3)


@pytest.mark.parametrize("alphabet,ksize", [("protein", 7), ("dayhoff", 12)])
def test_golden_csv_23_reads(refdata, reads_n22, misc_kats, alphabet, ksize):
    """tests/test_translate.py:279-315: the golden score CSVs, row by row."""
    g = orc.Nodegraph.load(os.path.join(
        refdata, "index", f"Homo_sapiens.GRCh38.pep.subset.alphabet-{alphabet}_ksize-{ksize}.bloomfilter.nodegraph"))
    true = pd.read_csv(os.path.join(
        refdata, "translate",
        f"SRR306838_GSM752691_hsa_br_F_1_trimmed_subsampled_n22__alphabet-{alphabet}_ksize-{ksize}.csv"))
    lut = orc.encode_lut(alphabet)
    rows = []
    coding = []
    for name, seq in read_fastx(reads_n22):
        res = orc.score_read(g, seq, lut, 0.5)
        pep = orc.six_frame_translation(seq)
        for r in res:
            rows.append((name, r["jaccard"], r["n_kmers"], category_text(int(r["category"]), misc_kats, alphabet),
                         int(r["frame"])))
            if r["category"] == orc.CAT_CODING:
                coding.append((name, int(r["frame"]), float(r["jaccard"]), pep[int(r["frame"])]))
    assert len(rows) == len(true) == 138
    for (name, j, n, cat, frame), t in zip(rows, true.itertuples()):
        assert name == t.read_id and cat == t.category and frame == t.translation_frame
        if math.isnan(t.n_kmers):
            assert n == -1
        else:
            assert n == int(t.n_kmers)
        if math.isnan(t.jaccard_in_peptide_db):
            assert math.isnan(j)
        else:
            assert j == pytest.approx(t.jaccard_in_peptide_db, abs=1e-12)
    if alphabet == "protein":
        fasta = "".join(
            ">{}\n{}\n".format(f"{n}__translation_frame:{f}__jaccard:{j}".replace(" ", "__"), p) for n, f, j, p in coding)
        true_fasta = open(os.path.join(refdata, "translate", "true_protein_coding.fasta")).read()
        assert true_fasta in fasta


@pytest.mark.parametrize("alphabet,ksize", SCORE_VECTORS)
def test_against_unmodified_reference_vectors(refdata, misc_kats, alphabet, ksize):
    check_reference_vector(load_score_vector(alphabet, ksize), refdata, misc_kats, alphabet, ksize)


def test_against_unmodified_reference_long_reads(refdata, misc_kats):
    """Reads of 6.1 - 9 kb scored by the unmodified reference (no length limit there): long ORFs, N, soft-masked, tandem repeat."""
    vec = load_score_vector("protein", 7, tag="long_")
    assert min(len(r["seq"]) for r in vec["reads"]) > 6144
    check_reference_vector(vec, refdata, misc_kats, "protein", 7)


def check_reference_vector(vec, refdata, misc_kats, alphabet, ksize):
    g = build_filter(os.path.join(refdata, vec["peptides"]), alphabet, ksize, vec["tablesize"], vec["n_tables"])
    assert [g.n_occupied(i) for i in range(4)] == vec["filter_popcounts"]
    lut = orc.encode_lut(alphabet)
    stdout = []
    for read in vec["reads"]:
        res = orc.score_read(g, read["seq"], lut, vec["jaccard_threshold"])
        pep = orc.six_frame_translation(read["seq"])
        for r, (j, n, cat, frame, hits) in zip(res, read["rows"]):
            assert int(r["frame"]) == frame
            assert category_text(int(r["category"]), misc_kats, alphabet) == cat, read["name"]
            assert (int(r["n_kmers"]) if r["n_kmers"] >= 0 else None) == n
            assert (int(r["n_hits"]) if r["n_hits"] >= 0 else None) == hits
            if j is None:
                assert math.isnan(r["jaccard"])
            else:
                assert float(r["jaccard"]) == j  # bit-exact: one IEEE division of the same integers
            if r["category"] == orc.CAT_CODING:
                h = f"{read['name']}__translation_frame:{frame}__jaccard:{float(r['jaccard'])}".replace(" ", "__")
                stdout.append(f">{h}\n{pep[frame]}\n")
            assert r["probes_min"] <= r["probes_max"]
    assert "".join(stdout) == vec["stdout"]


def test_empty_read_raises(refdata):
    g = orc.Nodegraph(7, 1e4, 4)
    with pytest.raises(ZeroDivisionError):
        orc.score_read(g, "", orc.encode_lut("protein"), 0.5)


def test_batch_matches_single_and_threads(refdata):
    vec = load_score_vector("protein", 7)
    g = build_filter(os.path.join(refdata, vec["peptides"]), "protein", 7)
    lut = orc.encode_lut("protein")
    seqs = [r["seq"].encode() for r in vec["reads"]]
    offsets = np.zeros(len(seqs) + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum([len(s) for s in seqs])
    flat = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    one, e1 = orc.score_reads(g, flat, offsets, lut, 0.5, n_threads=1)
    many, e2 = orc.score_reads(g, flat, offsets, lut, 0.5, n_threads=4)
    assert e1 == e2 == 0
    assert one.tobytes() == many.tobytes()
    for i, s in enumerate(seqs):
        assert orc.score_read(g, s, lut, 0.5).tobytes() == one[i].tobytes()


def test_staged_reference_with_native_shims_reproduces_the_golden_cli_outputs(golden_dir, tmp_path):
    """The thing `bench.py --impl reference` times -- the unmodified reference modules staged in oracle/_ref with khmer's
    Nodegraph and sourmash's hash_murmur as a native extension over the oracle's C code (oracle/refnative) -- must produce
    the committed outputs of the reference CLI (tests/golden/ref_cli, generated through the independent pure-python stand-ins
    of oracle/shims): same CSV, same stdout, same JSON summary.  Runs in a subprocess (the staged package is called
    `orpheum`); skipped where oracle/_ref has not been staged (no /root/reference and no prebuilt copy)."""
    import json
    import shutil
    import subprocess
    import sys

    from oracle import stage_reference

    if not stage_reference.staged() and not stage_reference.stage():
        pytest.skip("oracle/_ref is not staged and /root/reference is absent")
    root = os.path.dirname(golden_dir.rstrip("/"))  # tests/
    work = tmp_path / "golden_copy"
    for sub in ("reference_data", "ref_cli"):
        shutil.copytree(os.path.join(golden_dir, sub), work / sub)
    (work / "ours").mkdir()
    runs = json.load(open(os.path.join(golden_dir, "ref_cli", "manifest.json")))
    code = (
        "import sys, json, io, contextlib\n"
        f"sys.path.insert(0, {os.path.dirname(root)!r})\n"
        "from oracle import stage_reference\n"
        "stage_reference.activate()\n"
        "import khmer, orpheum.translate as T\n"
        "assert type(khmer.Nodegraph).__module__ == 'builtins' and khmer.Nodegraph.__module__ == '_orph_refnative'\n"
        "argv = json.loads(sys.argv[1])\n"
        "buf = io.StringIO()\n"
        "with contextlib.redirect_stdout(buf):\n"
        "    T.cli.main(argv, standalone_mode=False)\n"
        "open(sys.argv[2], 'w').write(buf.getvalue())\n"
    )
    for run in runs:
        tag = run["tag"]
        argv = [a.replace("ref_cli/" + tag, "ours/" + tag) for a in run["argv"]]
        proc = subprocess.run([sys.executable, "-c", code, json.dumps(argv), f"ours/{tag}.stdout.txt"], cwd=work, capture_output=True, text=True, timeout=600)
        assert proc.returncode == 0, proc.stderr[-2000:]
        ref = lambda ext: open(work / "ref_cli" / (tag + ext)).read()
        ours = lambda ext: open(work / "ours" / (tag + ext)).read()
        ref_stdout = ref(".stdout.txt")
        ref_stdout = ref_stdout[ref_stdout.index(">"):] if ">" in ref_stdout else ""
        assert ours(".stdout.txt") == ref_stdout, tag
        assert ours(".csv") == ref(".csv"), tag
        assert json.loads(ours(".json")) == json.loads(ref(".json")), tag
