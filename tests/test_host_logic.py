"""CPU tests of the host-side mirror of the reference interface (no GPU work): alphabets, constants,
emission-side translation, FASTA/FASTQ reading, output writers, click wiring.  Modelled on the
reference's tests/test_sequence_encodings.py, test_translate_single_seq.py, test_create_save_summary.py,
test_index.py:7-22,152-190 and test_translate.py:62-149,236-252."""
import gzip
import json
import math
import os

import numpy as np
import pandas as pd
import pytest
from click.testing import CliRunner

from orpheum_b200 import constants_index, constants_translate, fastx, sequence_encodings as enc
from orpheum_b200.create_save_summary import CreateSaveSummary
from orpheum_b200.translate_single_seq import TranslateSingleSeq
from tests.conftest import load_score_vector

SEQ = "CGCTTGCTTAATACTGACATCAATAATATTAGGAAAATCGCAATATAACTGTAAATCCTGTTCTGTC"


def test_alphabets_match_reference(misc_kats):
    probe = "ACDEFGHIKLMNPQRSTVWYXUBZOJ*acdxy"
    for name, encoded in misc_kats["alphabets"].items():
        assert enc.encode_peptide(probe, name) == encoded, name
        lut = enc.encode_lut(name)
        assert bytes(lut[np.frombuffer(probe.encode(), dtype=np.uint8)]).decode() == encoded
    with pytest.raises(ValueError):
        enc.encode_peptide("ACD", "not-an-alphabet")
    with pytest.raises(ValueError):
        enc.encode_lut("not-an-alphabet")
    assert enc.encode_peptide("SASHAFIERCE", "dayhoff") == "bbbdbfecdac"  # tests/test_sequence_encodings.py
    assert enc.encode_peptide("SASHAFIERCE", "hp") == "phpphhhpppp"
    assert enc.BEST_KSIZES == misc_kats["best_ksizes"]
    assert constants_translate.LOW_COMPLEXITY_CATEGORIES == misc_kats["low_complexity_categories"]
    for aa in "RHKDESTNQCGPAVILMFYW":
        for mapping in (enc.DAYHOFF_MAPPING, enc.HP_MAPPING, enc.BOTVINNIK_MAPPING, enc.AA9_MAPPING, enc.GBMR4_MAPPING,
                        enc.SDM12_MAPPING, enc.HSDM17_MAPPING, enc.DAYHOFF_V2_MAPPING):
            assert aa in mapping


def test_translate_single_seq(misc_kats):
    t = TranslateSingleSeq(SEQ, True)
    for codon, aa in misc_kats["codons"].items():
        assert t._translate_codon(codon) == aa
    assert t._translate_codon("AC") == ""
    assert t._reverse_complement("AACCGGTT") == "AACCGGTT" and t._reverse_complement("AACC") == "GGTT"
    for seq, frames in misc_kats["six_frame"].items():
        got = TranslateSingleSeq(seq, False).six_frame_translation()
        assert {str(k): v for k, v in got.items()} == frames
        assert list(got) == [1, 2, 3, -1, -2, -3]
        for f, pep in got.items():
            assert TranslateSingleSeq(seq, False).frame_translation(f) == pep
    assert TranslateSingleSeq(SEQ, True).six_frame_translation_no_stops() == {
        2: "ACLILTSIILGKSQYNCKSCSV", -2: "TEQDLQLYCDFPNIIDVSIKQA", -3: "QNRIYSYIAIFLILLMSVLSK"}


def test_thresholds_and_helpers(misc_kats):
    from orpheum_b200 import translate

    for alphabet, thr in misc_kats["thresholds"].items():
        assert translate.get_jaccard_threshold(None, alphabet) == thr
    assert translate.get_jaccard_threshold(0.5, "hp") == 0.5
    for seq, value in misc_kats["fastp_complexity"].items():
        assert translate.compute_fastp_complexity(seq) == value
    low = "CCCCCCCCCACCACCACCCCCCCCACCCCCCCCCCCCCCCCCCCCCCCCCCACCCCCCCACACACCCCCAACACCC"
    assert not translate.evaluate_is_fastp_low_complexity(SEQ) and translate.evaluate_is_fastp_low_complexity(low)
    assert translate.compute_kmer_complexity(SEQ, 7) == 30.5 and translate.compute_kmer_complexity(low, 7) == 35.0
    assert translate.evaluate_is_kmer_low_complexity(SEQ, 7) is False
    assert translate.evaluate_is_kmer_low_complexity(low, 7) is True


def test_index_helpers(misc_kats):
    from orpheum_b200 import index

    assert index.per_translation_false_positive_rate(14, 4e7) == misc_kats["fpr"]["per_translation_14_4e7"] == 3.275785921922898e-06
    assert index.per_read_false_positive_coding_rate(60, 7) == misc_kats["fpr"]["per_read_60_7"] == 3.884682410293139e-05
    assert index.get_peptide_ksize("protein") == 7 and index.get_peptide_ksize("dayhoff") == 12
    assert index.get_peptide_ksize("hydrophobic-polar") == 31 and index.get_peptide_ksize("protein", 11) == 11
    with pytest.raises(ValueError):
        index.get_peptide_ksize("not-a-molecule")
    assert constants_index.BASED_INT.convert("1e8", None, None) == 100000000
    assert constants_index.BASED_INT.convert("12345", None, None) == 12345
    assert constants_index.BASED_INT.convert(77, None, None) == 77


def test_fastx_reader(refdata, tmp_path):
    fq = os.path.join(refdata, "SRR306838_GSM752691_hsa_br_F_1_trimmed_subsampled_n22.fq")
    recs = list(fastx.open(fq))
    assert len(recs) == 23
    assert recs[0]["name"].startswith("SRR306838.") and " " in recs[0]["name"]  # full header line, not split
    gz = list(fastx.open(fq + ".gz"))
    # the gzipped fixture predates the 23rd (adversarial) read
    assert [(r["name"], r["sequence"]) for r in gz] == [(r["name"], r["sequence"]) for r in recs[:22]]
    fa = list(fastx.open(os.path.join(refdata, "index", "Homo_sapiens.GRCh38.pep.first1000lines.fa")))
    assert len(fa) > 100 and all("\n" not in r["sequence"] for r in fa)
    assert list(fastx.open(os.path.join(refdata, "empty_fasta.fasta"))) == []
    with pytest.raises(ValueError):
        fastx.open(os.path.join(refdata, "index", "Homo_sapiens.GRCh38.pep.subset.alphabet-protein_ksize-7.bloomfilter.nodegraph"))
    mixed = tmp_path / "lower.fa"
    mixed.write_text(">a b c\nacgt\nACGT\n>second\nNNNN\n")
    assert [(r["name"], r["sequence"]) for r in fastx.open(str(mixed))] == [("a b c", "acgtACGT"), ("second", "NNNN")]
    names, flat, offsets = next(fastx.read_batches(str(mixed)))
    assert names == ["a b c", "second"] and flat.tobytes() == b"acgtACGTNNNN" and offsets.tolist() == [0, 8, 12]
    batches = list(fastx.read_batches(fq, max_records=10))
    assert [len(b[0]) for b in batches] == [10, 10, 3]


def golden_rows(path):
    """Rows as Translate produces them (float / NaN jaccard, int / NaN n_kmers), from a golden CSV."""
    df = pd.read_csv(path, keep_default_na=True, float_precision="round_trip")
    rows = []
    for t in df.itertuples():
        j = float(t.jaccard_in_peptide_db)
        n = t.n_kmers
        rows.append([t.read_id, j, (np.nan if (isinstance(n, float) and math.isnan(n)) else int(n)), t.category,
                     int(t.translation_frame), t.filename])
    return rows


@pytest.mark.parametrize("tag,alphabet,ksize", [("n22_protein7", "protein", 7), ("n22_dayhoff12", "dayhoff", 12),
                                                ("synth_protein7", "protein", 7), ("synth_hp21", "hp", 21)])
def test_create_save_summary_against_reference_outputs(golden_dir, tmp_path, tag, alphabet, ksize):
    """CSV text and JSON summary written by the unmodified reference CLI (tests/golden/ref_cli) from the same rows."""
    ref_csv = os.path.join(golden_dir, "ref_cli", tag + ".csv")
    ref_json = json.load(open(os.path.join(golden_dir, "ref_cli", tag + ".json")))
    rows = golden_rows(ref_csv)
    out_csv, out_json, out_pq = tmp_path / "o.csv", tmp_path / "o.json", tmp_path / "o.parquet"
    summary = CreateSaveSummary([ref_json["input_files"][0]], str(out_csv), str(out_pq), str(out_json),
                                ref_json["peptide_bloom_filter"], alphabet, ksize, ref_json["jaccard_threshold"], rows)
    summary.maybe_write_csv()
    summary.maybe_write_parquet()
    summary.maybe_write_json_summary()
    assert out_csv.read_text() == open(ref_csv).read()
    got = json.load(open(out_json))
    assert got == ref_json
    pd.testing.assert_frame_equal(pd.read_parquet(out_pq), pd.read_csv(ref_csv))


def test_json_summary_kat_and_empty(golden_dir, tmp_path):
    # tests/test_create_save_summary.py:193-261 known answers for the 23 golden reads, protein k=7
    ref = json.load(open(os.path.join(golden_dir, "ref_cli", "n22_protein7.json")))
    assert ref["categorization_counts"]["Coding"] == 3 and ref["categorization_counts"]["Non-coding"] == 14
    assert ref["jaccard_info"]["count"] == 44
    out = tmp_path / "empty.json"
    s = CreateSaveSummary(["empty_fasta.fasta"], None, None, str(out), "f.nodegraph", "protein", 7, 0.5, [])
    s.maybe_write_csv()
    s.maybe_write_parquet()
    summary = s.maybe_write_json_summary()
    assert summary["jaccard_info"] == {"count": 0, "mean": None, "std": None, "min": None, "25%": None, "50%": None,
                                       "75%": None, "max": None}
    assert summary["categorization_counts"]["Coding"] == 0 and json.load(open(out)) == summary


def test_cli_wiring_errors():
    from orpheum_b200 import commandline, translate

    runner = CliRunner()
    assert runner.invoke(commandline.cli, ["--help"]).exit_code == 0
    assert runner.invoke(commandline.cli, ["translate", "--help"]).exit_code == 0
    assert runner.invoke(commandline.cli, ["index", "--help"]).exit_code == 0
    r = runner.invoke(translate.cli, ["--jaccard-threshold", "3.14", "pep.fa", "reads.fq"])
    assert r.exit_code == 2
    assert "--jaccard-threshold needs to be a number between 0 and 1, but 3.14 was provided" in r.output
    r = runner.invoke(translate.cli, ["--jaccard-threshold", "beyonce", "pep.fa", "reads.fq"])
    assert r.exit_code == 2 and "'beyonce' is not a valid float" in r.output
    r = runner.invoke(translate.cli, ["--long-reads", "pep.fa", "reads.fq"])
    assert isinstance(r.exception, NotImplementedError)


def test_native_parser_matches_python_reader(refdata, golden_dir, tmp_path):
    """csrc/fastx.cpp against the pure-Python reader on every fixture and on awkward files."""
    import gzip as gz

    from orpheum_b200 import build

    build.build()
    files = [os.path.join(refdata, "SRR306838_GSM752691_hsa_br_F_1_trimmed_subsampled_n22.fq"),
             os.path.join(refdata, "SRR306838_GSM752691_hsa_br_F_1_trimmed_subsampled_n22.fq.gz"),
             os.path.join(refdata, "SRR306838_GSM752691_hsa_br_F_1_trimmed_subsampled_first20.fq.gz"),
             os.path.join(refdata, "low_complexity_nucleotides.fastq"), os.path.join(refdata, "low_complexity_peptides.fastq"),
             os.path.join(refdata, "index", "Homo_sapiens.GRCh38.pep.first1000lines.fa"),
             os.path.join(refdata, "index", "Homo_sapiens.GRCh38.pep.subset.fa.gz"),
             os.path.join(refdata, "empty_fasta.fasta"), os.path.join(golden_dir, "ref_cli", "synth_mixed.fq")]
    odd = tmp_path / "odd.fa"
    odd.write_bytes(b">crlf header  \r\nACGT\r\nacgtn\r\n\r\n>no-sequence\n>last one\nAC\nGT")  # CRLF, blank line, no final newline
    wrapped = tmp_path / "wrapped.fq"
    wrapped.write_text("@r1 x\nACGT\nACGT\n+\nIIII\nIIII\n@r2\nNN\n+r2\n@@\n")  # multi-line record, '@' in the quality line
    files += [str(odd), str(wrapped)]
    with gz.open(tmp_path / "wrapped.fq.gz", "wt") as f:
        f.write(wrapped.read_text())
    files.append(str(tmp_path / "wrapped.fq.gz"))
    for path in files:
        py = [(r["name"], r["sequence"]) for r in fastx.open(path, native=False)]
        native = [(r["name"], r["sequence"]) for r in fastx.open(path, native=True)]
        assert native == py, path
        # batch form, small batches to exercise the resume logic
        names, seqs = [], []
        for n, flat, offsets in fastx.read_batches(path, max_records=7, native=True):
            data = flat.tobytes()
            names += n
            seqs += [data[int(offsets[i]):int(offsets[i + 1])].decode("latin-1") for i in range(len(n))]
        assert list(zip(names, seqs)) == py, path
    assert [(r["name"], r["sequence"]) for r in fastx.open(str(odd))] == [("crlf header", "ACGTacgtn"), ("no-sequence", ""), ("last one", "ACGT")]
    assert [(r["name"], r["sequence"]) for r in fastx.open(str(wrapped))] == [("r1 x", "ACGTACGT"), ("r2", "NN")]
    with pytest.raises(ValueError):
        fastx.open(os.path.join(refdata, "index", "Homo_sapiens.GRCh38.pep.subset.alphabet-protein_ksize-7.bloomfilter.nodegraph"))
    bad = tmp_path / "bad.fq"
    bad.write_text("@r1\nACGT\n+\nII\n")
    with pytest.raises(OSError):
        list(fastx.open(str(bad)))
    with pytest.raises(OSError):
        list(fastx.open(str(bad), native=False))


def test_native_float_repr_matches_python():
    """The CSV's float text: every jaccard ratio h/n the 16-bit counters can form for n <= 400, plus extremes."""
    import ctypes

    from orpheum_b200 import _lib, build

    build.build()
    lib = _lib.load()
    buf = ctypes.create_string_buffer(32)

    def native(v):
        assert lib.orph_format_float_repr(v, buf) == 0
        return buf.value.decode()

    for n in range(1, 401):
        for h in range(0, n + 1):
            assert native(h / n) == repr(h / n), (h, n)
    for n in (65535, 65521, 4099, 12345):
        for h in (0, 1, 2, 3, 7, n // 3, n - 1, n):
            assert native(h / n) == repr(h / n), (h, n)
    for v in (0.0, 1.0, 0.5, 1e-4, 1e-5, 9.999e-5, 123456.789, 1e15, 1e16, 1.5e17, 2.0**-40, 3.0, 100.0, -0.25, 1e22, 5e-324):
        assert native(v) == repr(v), v
    assert native(float("nan")) == "nan"


def test_parallel_fastq_fast_path(tmp_path):
    """The strict four-line FASTQ fast path (newline index + validation + copies spread over the host cores) against the
    pure-Python reader: files large enough for several segments and the thread pool, CRLF, spaces around fields, no final
    newline, '@' and '+' as first quality characters, and layouts that must fall back to the general parser in the middle
    of a file (a wrapped record, a blank line, an empty sequence)."""
    from orpheum_b200 import build

    build.build()
    rng = np.random.default_rng(4)
    bases = np.frombuffer(b"ACGTN", dtype=np.uint8)

    def records(n, start=0):
        out = []
        for i in range(start, start + n):
            L = int(rng.integers(30, 300))
            seq = bases[rng.integers(0, 5 if i % 50 == 0 else 4, L)].tobytes()
            qual = bytes(rng.integers(33, 74, L, dtype=np.uint8))
            if i % 7 == 0:
                qual = b"@" + qual[1:]
            if i % 11 == 0:
                qual = b"+" + qual[1:]
            out.append((b"read%d lane:%d  extra" % (i, i % 8), seq, qual))
        return out

    def render(recs, eol=b"\n", final=True):
        text = b"".join(b"@" + n + eol + s + eol + b"+" + (n if i % 3 == 0 else b"") + eol + q + eol for i, (n, s, q) in enumerate(recs))
        return text if final else text[: -len(eol)]

    def check(path, **kw):
        py = [(r["name"], r["sequence"]) for r in fastx.open(path, native=False)]
        got_names, got_seqs = [], []
        for names, flat, offsets in fastx.read_batches(path, native=True, **kw):
            data = flat.tobytes()
            got_names += names.tolist()
            got_seqs += [data[int(offsets[i]):int(offsets[i + 1])].decode("latin-1") for i in range(len(names))]
        assert len(py) == len(got_names)
        assert list(zip(got_names, got_seqs)) == py, path
        return len(py)

    big = records(60_000)
    p = tmp_path / "big.fq"
    p.write_bytes(render(big))
    assert p.stat().st_size > (16 << 20)
    assert check(str(p)) == 60_000
    assert check(str(p), max_records=25_000) == 60_000
    assert check(str(p), max_bases=1_000_000) == 60_000
    p.write_bytes(render(big[:5000], eol=b"\r\n", final=False))
    assert check(str(p)) == 5000
    # padded fields
    p.write_bytes(b"@ r1 \t\n  ACGT \n+\n IIII\t\n@r2\nAC\n+\nII")
    assert check(str(p)) == 2
    # layouts the fast path must hand over to the general parser, at the start, in the middle and at the end
    head, tail = render(big[:30_000]), render(big[30_000:40_000])
    wrapped = b"@w x\nACGT\nACGT\n+\nIIII\nIIII\n"
    for odd in (wrapped, b"\n", b"@empty\n\n+\n\n", b"@hash\nACGT\n#\nIIII\n"):
        for layout in (odd + head, head + odd + tail, head + odd):
            if layout[:1] != b"@":
                continue  # a file that starts with a blank line is not FASTQ for either reader
            p.write_bytes(layout)
            check(str(p), max_records=8192)
    # a truncated file raises like the general parser
    p.write_bytes(head + b"@r\nACGT\n+\nII\n")
    with pytest.raises(OSError):
        check(str(p))


def test_native_parser_random_files(tmp_path):
    """Seeded random FASTQ / FASTA files (strict and sloppy layouts mixed: CRLF, blank lines, wrapped sequences and
    qualities, '+name' separators, '@' / '+' / '>' as first quality characters, trailing spaces, missing final newline)
    through the native parser -- fast path, general parser and the hand-over between them -- against the Python reader,
    at several batch sizes."""
    from orpheum_b200 import build

    build.build()
    rng = np.random.default_rng(12)
    alphabet = np.frombuffer(b"ACGTNacgtn", dtype=np.uint8)

    def fastq(n_rec, sloppy):
        out = []
        for i in range(n_rec):
            L = int(rng.integers(1, 120))
            seq = alphabet[rng.integers(0, len(alphabet), L)].tobytes()
            qual = bytes(rng.integers(33, 75, L, dtype=np.uint8))
            eol = b"\r\n" if sloppy and rng.random() < 0.2 else b"\n"
            name = b"r%d %s" % (i, bytes(rng.integers(48, 90, int(rng.integers(0, 12)), dtype=np.uint8)))
            if sloppy and rng.random() < 0.15:  # wrapped record
                cut = int(rng.integers(0, L + 1))
                seq_lines = [s for s in (seq[:cut], seq[cut:]) if s]
                qual_lines = [q for q in (qual[:cut], qual[cut:]) if q]
            else:
                seq_lines, qual_lines = [seq], [qual]
            rec = b"@" + name + (b"  " if sloppy and rng.random() < 0.1 else b"") + eol
            rec += b"".join(s + eol for s in seq_lines) + b"+" + (name if rng.random() < 0.3 else b"") + eol
            rec += b"".join(q + eol for q in qual_lines)
            if sloppy and rng.random() < 0.1:
                rec += eol  # blank line between records
            out.append(rec)
        text = b"".join(out)
        if rng.random() < 0.3 and text.endswith(b"\n") and not text.endswith(b"\n\n") and not text.endswith(b"\r\n"):
            text = text[:-1]
        return text

    def fasta(n_rec):
        out = []
        for i in range(n_rec):
            L = int(rng.integers(0, 200))
            seq = alphabet[rng.integers(0, len(alphabet), L)].tobytes()
            width = int(rng.integers(1, 80))
            out.append(b">p%d some description\n" % i + b"".join(seq[j:j + width] + b"\n" for j in range(0, L, width)))
        return b"".join(out)

    p = tmp_path / "f.fx"
    for case in range(120):
        kind = case % 3
        text = fastq(int(rng.integers(1, 400)), sloppy=False) if kind == 0 else fastq(int(rng.integers(1, 400)), sloppy=True) if kind == 1 \
            else fasta(int(rng.integers(1, 100)))
        p.write_bytes(text)
        try:
            want = [(r["name"], r["sequence"]) for r in fastx.open(str(p), native=False)]
        except OSError:
            with pytest.raises(OSError):
                list(fastx.read_batches(str(p)))
            continue
        for max_records in (1 << 20, 37):
            got = []
            for names, flat, offsets in fastx.read_batches(str(p), max_records=max_records, prefetch=case % 2):
                data = flat.tobytes()
                got += [(names[i], data[int(offsets[i]):int(offsets[i + 1])].decode("latin-1")) for i in range(len(names))]
            assert got == want, (case, max_records)


def test_columnar_writers_against_row_list_writers_random(tmp_path):
    """Random result records (every category mix, duplicate and awkward read ids, several input files) through the
    columnar writers (native CSV formatter, arrow table, vectorised JSON summary) and through the row-list
    CreateSaveSummary that mirrors the reference's object interface: CSV bytes, parquet tables and JSON summaries agree."""
    import json as js

    import pyarrow.parquet as pq

    from orpheum_b200 import _lib, build
    from orpheum_b200.columnar import ScoreTable
    from orpheum_b200.constants_translate import FRAMES, category_text
    from orpheum_b200.create_save_summary import CreateSaveSummary

    build.build()
    rng = np.random.default_rng(21)
    for case in range(12):
        alphabet = ["protein", "dayhoff", "hp"][case % 3]
        table, rows, files = ScoreTable(alphabet), [], []
        for fi in range(int(rng.integers(1, 4))):
            filename = f"dir {fi}/reads,{fi}.fq" if case % 4 == 0 else f"reads{fi}.fq"
            n = int(rng.integers(1, 300))
            rec = np.zeros(n, dtype=_lib.READ_RESULT_DTYPE)
            style = rng.integers(0, 4)
            probs = [[.5, .1, .1, .15, .15, 0], [0, 0, 0, .5, .5, 0], [.9, .1, 0, 0, 0, 0], [.3, .1, .2, .2, .2, 0]][style]
            cat = rng.choice(6, size=(n, 6), p=probs)
            low = rng.random(n) < 0.1
            cat[low] = 5  # a fastp-low-complexity read carries the same label on all six frames
            rec["category"] = cat
            nk = rng.integers(1, 200, size=(n, 6))
            nh = (rng.random((n, 6)) * (nk + 1)).astype(np.int64)
            has_n = (cat == 2) | (cat == 3) | (cat == 4)
            scored = (cat == 3) | (cat == 4)
            rec["n_kmers"] = np.where(has_n, nk, 0)
            rec["n_hits"] = np.where(scored, np.minimum(nh, nk), 0)
            names = []
            for i in range(n):
                kind = rng.integers(0, 12)
                name = f"read{i} lane:{fi} file{fi}"
                if kind == 0:
                    name = f'id,with,commas "{i}" {fi}'
                elif kind == 1 and names and case % 4 != 3:  # (every fourth case keeps all ids unique: the summary's fast path)
                    name = names[-1]  # duplicate of the previous id (groupby merges the rows)
                elif kind == 2 and len(names) > 3 and case % 4 != 3:
                    name = names[0]   # duplicate of a distant id (Counter merges the coding counts)
                elif kind == 3:
                    name = f"caf\xe9 \u4e2d {i}"  # UTF-8 ids keep their bytes in the CSV and their characters in the parquet
                names.append(name)
            table.append(names, rec, filename)
            files.append(filename)
            for i in range(n):
                for f in range(6):
                    c = int(cat[i, f])
                    jac = int(rec["n_hits"][i, f]) / int(rec["n_kmers"][i, f]) if c in (3, 4) else np.nan
                    rows.append([names[i], jac, int(rec["n_kmers"][i, f]) if c in (2, 3, 4) else np.nan, category_text(c, alphabet), FRAMES[f],
                                 filename])
        d = tmp_path / f"case{case}"
        d.mkdir()
        ref = CreateSaveSummary(files, str(d / "r.csv"), str(d / "r.parquet"), str(d / "r.json"), "f.nodegraph", alphabet, 7, 0.5, rows)
        ref.maybe_write_csv()
        ref.maybe_write_parquet()
        ref.maybe_write_json_summary()
        table.write_csv(str(d / "c.csv"))
        table.write_parquet(str(d / "c.parquet"))
        table.write_json_summary(str(d / "c.json"), files, "f.nodegraph", 7, 0.5)
        assert (d / "c.csv").read_bytes() == (d / "r.csv").read_bytes(), case
        a, b = pq.read_table(str(d / "c.parquet")), pq.read_table(str(d / "r.parquet"))
        assert a.schema.equals(b.schema) and a.to_pandas().equals(b.to_pandas()), case
        assert js.load(open(d / "c.json")) == js.load(open(d / "r.json")), case
        if case % 4 == 3:
            # the same table with the parser's name blobs (fastx.NameList): the summary is computed straight from the record
            # arrays (ids proven distinct by their hashes) and must be the same dict, floats to the last bit
            fast = ScoreTable(alphabet, background_summary=case % 8 == 3)  # (half of them through the per-batch background shares)
            for names_, rec_, (filename_, _) in zip(table._names, table._records, table._files):
                enc = [nm.encode(*fastx.TEXT_CODEC) for nm in names_]
                offs = np.zeros(len(enc) + 1, dtype=np.uint64)
                offs[1:] = np.cumsum([len(e) for e in enc])
                fast.append(fastx.NameList(b"".join(enc), offs), rec_, filename_)
            every_id = [nm for names_ in table._names for nm in names_]
            assert fast._ids_all_distinct() == (len(set(every_id)) == len(every_id))
            fast.write_json_summary(str(d / "f.json"), files, "f.nodegraph", 7, 0.5)
            assert js.load(open(d / "f.json")) == js.load(open(d / "r.json")), case
            assert (d / "f.json").read_text() == (d / "c.json").read_text(), case


def test_non_ascii_read_ids_keep_their_bytes(tmp_path):
    """Read ids are decoded as UTF-8 (screed does) with surrogateescape: UTF-8 ids come out as the same characters, any
    other byte string survives byte for byte into the CSV."""
    from orpheum_b200 import _lib, build
    from orpheum_b200.columnar import ScoreTable

    build.build()
    fq = tmp_path / "n.fq"
    fq.write_bytes("@café 中文 1\nACGT\n+\nIIII\n".encode("utf-8") + b"@raw\xe9\xff 2\nAC\n+\nII\n@plain 3\nA\n+\nI\n")
    for native in (True, False):
        names = [r["name"] for r in fastx.open(str(fq), native=native)]
        assert names[0] == "café 中文 1" and names[2] == "plain 3"
        assert names[1].encode(*fastx.TEXT_CODEC) == b"raw\xe9\xff 2"
    (names, flat, offsets), = list(fastx.read_batches(str(fq)))
    assert names.tolist() == [r["name"] for r in fastx.open(str(fq), native=False)]
    rec = np.zeros(3, dtype=_lib.READ_RESULT_DTYPE)
    rec["category"] = 1
    for ids in (names, names.tolist()):
        table = ScoreTable("protein")
        table.append(ids, rec, "n.fq")
        table.write_csv(str(tmp_path / "o.csv"))
        text = (tmp_path / "o.csv").read_bytes()
        assert "café 中文 1,nan".encode("utf-8") in text and b"raw\xe9\xff 2,nan" in text
        table.write_parquet(str(tmp_path / "o.parquet"))


def test_native_fasta_formatter_against_python(tmp_path):
    """orph_format_fasta (the text of the stdout / FASTA records, translate.py:244-298) against the reference's python
    formatting on random records: every output stream, ids with spaces and UTF-8, float repr of the jaccard."""
    from orpheum_b200 import _lib, build, gpu
    from orpheum_b200.constants_translate import FRAMES

    build.build()
    rng = np.random.default_rng(31)
    for with_lcp in (True, False):
        n = 500
        rec = np.zeros(n, dtype=_lib.READ_RESULT_DTYPE)
        cat = rng.choice(6, size=(n, 6), p=[.4, .05, .1, .25, .2, 0])
        rec["category"] = cat
        nk = rng.integers(1, 300, size=(n, 6))
        rec["n_kmers"] = np.where(np.isin(cat, (2, 3, 4)), nk, 0)
        rec["n_hits"] = np.where(np.isin(cat, (3, 4)), (rng.random((n, 6)) * (nk + 1)).astype(np.int64).clip(0, nk), 0)
        names = [f"read {i}  two  spaces/{i % 3}" if i % 5 else f"caf\xe9 {i}" for i in range(n)]
        reads = ["".join(rng.choice(list("ACGTN"), size=int(rng.integers(20, 90)))) for _ in range(n)]
        flat, offsets = gpu.concat_reads(reads)
        name_bytes = [s.encode("utf-8") for s in names]
        name_offsets = np.zeros(n + 1, dtype=np.uint64)
        name_offsets[1:] = np.cumsum([len(b) for b in name_bytes])
        name_blob = np.frombuffer(b"".join(name_bytes), dtype=np.uint8)
        peps, pep_of = [], {}
        for i in range(n):
            for f in range(6):
                if cat[i, f] == 3 or (with_lcp and cat[i, f] == 2):
                    pep_of[(i, f)] = "".join(rng.choice(list("ACDEFGHIKLMNPQRSTVWYX*"), size=int(rng.integers(0, 30))))
                    peps.append(pep_of[(i, f)])
        pep_offsets = np.zeros(len(peps) + 1, dtype=np.uint64)
        pep_offsets[1:] = np.cumsum([len(p) for p in peps])
        pep_blob = np.frombuffer("".join(peps).encode(), dtype=np.uint8) if pep_offsets[-1] else np.zeros(1, dtype=np.uint8)
        want = {0: [], 1: [], 2: [], 3: []}
        for i in range(n):
            for f, frame in enumerate(FRAMES):
                c = int(cat[i, f])
                if c == 2:
                    want[3].append(">{}\n{}\n".format(names[i] + " translation_frame: {}.format(frame)", pep_of.get((i, f), "")))
                if c in (3, 4):
                    jaccard = int(rec["n_hits"][i, f]) / int(rec["n_kmers"][i, f])
                    seqname = ("{}__translation_frame:{}".format(names[i], frame) + "__jaccard:{}".format(jaccard)).replace(" ", "__")
                    if c == 3:
                        want[0].append(">{}\n{}\n".format(seqname, pep_of[(i, f)]))
                        want[1].append(">{}\n{}\n".format(seqname, reads[i]))
                    else:
                        want[2].append(">{}\n{}\n".format(seqname, reads[i]))
        for kind in (0, 1, 2) + ((3,) if with_lcp else ()):
            got = gpu.format_fasta(kind, name_blob, name_offsets, rec, flat, offsets, pep_blob, pep_offsets, with_lcp)
            assert got.decode("utf-8") == "".join(want[kind]), (kind, with_lcp)


def test_native_csv_writer_many_blocks_and_threads(tmp_path, monkeypatch):
    """orph_csv_write_scores formats blocks of reads on several threads and writes them at their places in the file: a
    table of many blocks, appended in two calls (copied into a shared mapping of the file's tail, or pwritten), against
    csv.writer over the reference's rows."""
    import csv as pycsv

    from orpheum_b200 import _lib, build
    from orpheum_b200.columnar import ScoreTable
    from orpheum_b200.constants_translate import FRAMES, category_text

    build.build()
    rng = np.random.default_rng(99)
    n = 70_000
    rec = np.zeros(n, dtype=_lib.READ_RESULT_DTYPE)
    cat = rng.choice(6, size=(n, 6), p=[.4, .1, .1, .2, .2, 0])
    cat[rng.random(n) < 0.1] = 5
    nk = rng.integers(1, 300, size=(n, 6))
    nh = np.minimum((rng.random((n, 6)) * (nk + 1)).astype(np.int64), nk)
    rec["category"] = cat
    rec["n_kmers"] = np.where((cat >= 2) & (cat <= 4), nk, 0)
    rec["n_hits"] = np.where((cat == 3) | (cat == 4), nh, 0)
    names = [f"read{i}/{'x' * (i % 40)}" for i in range(n)]
    table = ScoreTable("protein")
    half = n // 2
    table.append(names[:half], rec[:half], "a.fq")
    table.append(names[half:], rec[half:], "b,with comma.fq")
    table.write_csv(str(tmp_path / "native.csv"))
    with open(tmp_path / "python.csv", "w", newline="") as f:
        w = pycsv.writer(f, lineterminator="\n")
        w.writerow(["read_id", "jaccard_in_peptide_db", "n_kmers", "category", "translation_frame", "filename"])
        for i in range(n):
            for fr in range(6):
                c = int(cat[i, fr])
                jac = int(rec["n_hits"][i, fr]) / int(rec["n_kmers"][i, fr]) if c in (3, 4) else np.nan
                w.writerow([names[i], jac, int(rec["n_kmers"][i, fr]) if c in (2, 3, 4) else np.nan, category_text(c, "protein"), FRAMES[fr],
                            "a.fq" if i < half else "b,with comma.fq"])
    assert (tmp_path / "native.csv").read_bytes() == (tmp_path / "python.csv").read_bytes()
    # the same table through the pwrite() route (taken where the file cannot be mapped)
    monkeypatch.setenv("ORPHEUM_B200_CSV_PWRITE", "1")
    table.write_csv(str(tmp_path / "native_pwrite.csv"))
    assert (tmp_path / "native_pwrite.csv").read_bytes() == (tmp_path / "python.csv").read_bytes()


def test_mod_prime_tiny_arithmetic():
    """The 32-bit Barrett pieces of device_common.cuh's mod_prime_tiny, restated on Python integers: exact for every prime the
    kernels may hand it (p <= (2^32 - 1) / 3), at the limit and at values chosen against each dropped term."""
    import random

    M32, M64 = 0xFFFFFFFF, 0xFFFFFFFFFFFFFFFF

    def tiny(h, p):
        magic = M64 // p
        hi, lo, mhi, mlo = h >> 32, h & M32, magic >> 32, magic & M32
        mid = (hi * mlo + lo * mhi) & M64
        q = (hi * mhi + (mid >> 32)) & M32
        r = (lo - q * p) & M32
        for _ in range(2):
            r = min(r, (r - p) & M32)
        return r

    rng = random.Random(4)
    primes = [2, 3, 251, 65521, 99999989, 99999971, 999999937, 999999883, 1073741789, 1431655751, M32 // 3]
    for p in primes:
        edge = [0, 1, p - 1, p, p + 1, M32, M32 + 1, M64, M64 - 1, (M64 // p) * p, (M64 // p) * p - 1, M64 - p, M32 << 32]
        for h in edge + [rng.getrandbits(64) for _ in range(20000)] + [rng.getrandbits(64) | (M32 << 32) for _ in range(2000)] \
                + [rng.randrange(M64 // p) * p + d for _ in range(2000) for d in (0, p - 1)]:
            assert tiny(h, p) == h % p, (h, p)


def test_native_summary_shares_match_the_numpy_formulation():
    """orph_summary_partial (classes, coding-frame histogram in first-appearance order, jaccard column) against the numpy
    functions it replaces on the CLI's background thread, and orph_f64_pairwise_sum against np.sum / np.nanmean / np.nanstd to
    the last bit, over sizes around every split of numpy's pairwise tree."""
    from orpheum_b200 import _lib, build, columnar

    build.build()
    rng = np.random.default_rng(2024)
    for n in (1, 7, 1000, 70_001):
        rec = np.zeros(n, dtype=_lib.READ_RESULT_DTYPE)
        cat = rng.choice(6, size=(n, 6), p=[.45, .1, .05, .15, .2, .05])
        cat[rng.random(n) < 0.1] = 5
        cat[rng.random(n) < 0.1] = 0
        cat[rng.random(n) < 0.05] = 4
        nk = rng.integers(1, 300, size=(n, 6))
        rec["category"] = cat
        rec["n_kmers"] = nk
        rec["n_hits"] = np.minimum((rng.random((n, 6)) * (nk + 1)).astype(np.int64), nk)
        got = columnar._partial_summary([f"r{i}" for i in range(n)], rec)
        present, coding_counts, jaccard, scored = columnar._record_columns(rec)
        want_classes = columnar._class_counts(present)
        assert got["classes"][:3] == want_classes[:3] and got["classes"][4] == want_classes[4]
        np.testing.assert_array_equal(got["classes"][3], want_classes[3])
        want_hist = columnar._coding_histogram(coding_counts)
        assert list(got["histogram"].items()) == list(want_hist.items())
        np.testing.assert_array_equal(got["scored"], scored)
        np.testing.assert_array_equal(got["jaccard"], jaccard)  # NaN == NaN here
    assert columnar._native_sums_match_numpy()
    for n in [1, 7, 8, 9, 127, 128, 129, 255, 256, 257, 1023, 4097, 131072, 131073, 262145, 1_500_003]:
        x = rng.integers(0, 50, n) / rng.integers(1, 50, n)
        ok = rng.random(n) < 0.5
        assert columnar._pairwise_sum(x, None) == float(np.sum(x))
        assert columnar._pairwise_sum(x, ok) == float(np.sum(np.where(ok, x, 0.0)))
        if ok.any():
            with_nan = np.where(ok, x, np.nan)
            mean = columnar._pairwise_sum(x, ok) / int(ok.sum())
            assert mean == float(np.nanmean(with_nan))
            assert float(np.sqrt(np.float64(columnar._pairwise_sum(x, ok, mean, True)) / int(ok.sum()))) == float(np.nanstd(with_nan))
    # the statistics of a large column: native sums and numpy's give the same dict
    x = rng.integers(0, 50, 1_200_000) / rng.integers(1, 50, 1_200_000)
    ok = rng.random(x.size) < 0.4
    fast = columnar.jaccard_info((x, ok))
    slow = columnar.jaccard_info(np.where(ok, x, np.nan))
    assert fast == slow
