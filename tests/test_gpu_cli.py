"""End-to-end parity of the host mirror + CUDA path against outputs of the UNMODIFIED reference CLI
(tests/golden/ref_cli, produced by oracle/gen_golden.py): CSV text, JSON summary, stdout FASTA and the four
optional FASTA files, byte for byte.  Mirrors tests/test_translate.py:210-445 and test_index.py:100-149."""
import json
import os
import shutil

import numpy as np
import pandas as pd
import pytest
from click.testing import CliRunner

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def workdir(tmp_path_factory, golden_dir):
    """A scratch copy of tests/golden so that the CLI's relative paths match the reference runs."""
    root = tmp_path_factory.mktemp("golden_copy")
    for sub in ("reference_data", "ref_cli"):
        shutil.copytree(os.path.join(golden_dir, sub), os.path.join(root, sub))
    out = os.path.join(root, "ours")
    os.makedirs(out)
    return str(root)


def manifest(golden_dir):
    return json.load(open(os.path.join(golden_dir, "ref_cli", "manifest.json")))


@pytest.mark.parametrize("tag", ["n22_protein7", "n22_dayhoff12", "synth_protein7", "synth_hp21"])
def test_translate_cli_matches_reference_outputs(workdir, golden_dir, tag, monkeypatch):
    from orpheum_b200 import build, translate

    build.build()
    run = next(m for m in manifest(golden_dir) if m["tag"] == tag)
    argv = [a.replace("ref_cli/" + tag, "ours/" + tag) for a in run["argv"]]
    monkeypatch.chdir(workdir)
    result = CliRunner().invoke(translate.cli, argv + ["--parquet", f"ours/{tag}.parquet"])
    assert result.exit_code == 0, result.exception
    ref = lambda ext: open(os.path.join("ref_cli", tag + ext)).read()
    ours = lambda ext: open(os.path.join("ours", tag + ext)).read()
    # CliRunner mixed tqdm's stderr progress lines ("274it [00:00, ...]") into the reference capture when the
    # filter was built from a FASTA; the contract is the FASTA text that follows
    ref_stdout = ref(".stdout.txt")
    ref_stdout = ref_stdout[ref_stdout.index(">"):] if ">" in ref_stdout else ""
    assert result.output == ref_stdout
    assert ours(".csv") == ref(".csv")
    for ext in (".coding_nt.fa", ".noncoding_nt.fa", ".lowcplx_nt.fa", ".lowcplx_pep.fa"):
        assert ours(ext) == ref(ext), ext
    assert json.loads(ours(".json")) == json.loads(ref(".json"))
    pd.testing.assert_frame_equal(pd.read_parquet(f"ours/{tag}.parquet"), pd.read_csv(os.path.join("ref_cli", tag + ".csv")))
    if not run["peptides_are_bloom_filter"]:
        saved = os.path.splitext(run["peptides"])[0] + f".alphabet-{run['alphabet']}_ksize-{run['ksize']}.bloomfilter.nodegraph"
        assert os.path.exists(saved)
        os.remove(saved)


def test_golden_fasta_and_csv_of_the_reference_suite(workdir, refdata, monkeypatch):
    """tests/test_translate.py:210-233,279-315 as written there: stdout contains true_protein_coding.fasta."""
    from orpheum_b200 import translate

    monkeypatch.chdir(workdir)
    reads = "reference_data/SRR306838_GSM752691_hsa_br_F_1_trimmed_subsampled_n22.fq"
    true_fasta = open("reference_data/translate/true_protein_coding.fasta").read()
    for alphabet, k in (("protein", 7), ("dayhoff", 12)):
        # from the peptide FASTA (builds the filter on the GPU with the default tablesize 1e8)
        result = CliRunner().invoke(translate.cli, ["--peptide-ksize", str(k), "--alphabet", alphabet,
                                                    "reference_data/index/Homo_sapiens.GRCh38.pep.subset.fa.gz", reads])
        assert result.exit_code == 0, result.exception
        assert true_fasta in result.output
        # from the golden Bloom filter
        csv = f"ours/golden_{alphabet}.csv"
        result = CliRunner().invoke(translate.cli, [
            "--peptide-ksize", str(k), "--csv", csv, "--peptides-are-bloom-filter", "--alphabet", alphabet,
            f"reference_data/index/Homo_sapiens.GRCh38.pep.subset.alphabet-{alphabet}_ksize-{k}.bloomfilter.nodegraph", reads])
        assert result.exit_code == 0 and true_fasta in result.output
        true = pd.read_csv(f"reference_data/translate/SRR306838_GSM752691_hsa_br_F_1_trimmed_subsampled_n22__alphabet-{alphabet}_ksize-{k}.csv")
        true["filename"] = reads
        pd.testing.assert_frame_equal(pd.read_csv(csv), true)
    # ksize mismatch with the filter -> ValueError (index.py:162-171)
    result = CliRunner().invoke(translate.cli, [
        "--peptide-ksize", "9", "--peptides-are-bloom-filter",
        "reference_data/index/Homo_sapiens.GRCh38.pep.subset.alphabet-protein_ksize-7.bloomfilter.nodegraph", reads])
    assert isinstance(result.exception, ValueError)
    # empty input still writes a summary (tests/test_translate.py:426-445)
    result = CliRunner().invoke(translate.cli, ["--json-summary", "ours/empty.json", "--tablesize", "1e6",
                                                "reference_data/index/Homo_sapiens.GRCh38.pep.first1000lines.fa",
                                                "reference_data/empty_fasta.fasta"])
    assert result.exit_code == 0 and os.path.exists("ours/empty.json")


def test_translate_class_interface(workdir, monkeypatch, capsys):
    """The reference's Translate unit tests (tests/test_translate.py:127-207)."""
    from orpheum_b200 import constants_index, translate

    monkeypatch.chdir(workdir)
    args = dict(reads=["reference_data/SRR306838_GSM752691_hsa_br_F_1_trimmed_subsampled_n22.fq"],
                peptides="reference_data/index/Homo_sapiens.GRCh38.pep.subset.fa.gz", peptide_ksize=None,
                save_peptide_bloom_filter=False, peptides_are_bloom_filter=False, jaccard_threshold=None, alphabet="protein",
                csv=False, parquet=False, json_summary=False, coding_nucleotide_fasta="ours/c.fa",
                noncoding_nucleotide_fasta="ours/n.fa", low_complexity_nucleotide_fasta="ours/ln.fa",
                low_complexity_peptide_fasta="ours/lp.fa", tablesize=constants_index.DEFAULT_MAX_TABLESIZE,
                n_tables=constants_index.DEFAULT_N_TABLES, long_reads=False, verbose=True)
    t = translate.Translate(args)
    assert t.peptide_ksize == 7 and t.nucleotide_ksize == 21 and t.get_jaccard_threshold() == 0.5
    frac, n = t.score_single_translation({2: "ACLILTSIILGKSQYNCKSCSV", -2: "TEQDLQLYCDFPNIIDVSIKQA", -3: "QNRIYSYIAIFLILLMSVLSK"})
    assert n == 82 and abs(frac - 0.19) < 0.05  # the reference's dict quirk (tests/test_translate.py:187-192)
    assert t.get_coding_score_line("description", 1.0, 40, None, -3) == ["description", 1.0, 40, "Coding", -3]
    assert t.get_coding_score_line("description", 0.5, 40, "test special case", -2) == ["description", 0.5, 40, "test special case", -2]
    t.set_coding_scores_all_files()
    assert len(t.fastas) == 4 and len(t.file_handles) == 4 and len(t.coding_scores) == 138
    for f in ("ours/c.fa", "ours/n.fa", "ours/ln.fa", "ours/lp.fa"):
        assert os.path.exists(f)
    assert open("ours/ln.fa").read() == ""  # the reference never writes this file (quirk)
    capsys.readouterr()
    t.file_handles = dict.fromkeys(t.file_handles)  # the files above are closed now
    scores = t.maybe_score_single_read("gold", "CGCTTGCTTAATACTGACATCAATAATATTAGGAAAATCGCAATATAACTGTAAATCCTGTTCTGTC")
    out = capsys.readouterr().out
    assert [s.translation_frame for s in scores] == [1, 2, 3, -1, -2, -3]
    assert scores[4] == (1.0, 16, "Coding", -2) and scores[1][:3] == (0.0, 16, "Non-coding")
    assert out == ">gold__translation_frame:-2__jaccard:1.0\nTEQDLQLYCDFPNIIDVSIKQA\n"
    with pytest.raises(ZeroDivisionError):
        t.maybe_score_single_read("empty", "")


def test_index_cli(workdir, refdata, monkeypatch):
    """tests/test_index.py:25-149: n_unique_kmers of built filters, CLI, --save-as, --index-from-dir, load."""
    from orpheum_b200 import index

    monkeypatch.chdir(workdir)
    first = "reference_data/index/Homo_sapiens.GRCh38.pep.first1000lines.fa"
    expected = {("protein", 7): 13966, ("dayhoff", 7): 10090, ("dayhoff", 11): 13605, ("dayhoff", 12): 13816,
                ("hydrophobic-polar", 31): 12888, ("hydrophobic-polar", 21): 13076, ("hydrophobic-polar", 7): 136}
    for (alphabet, k), n in expected.items():
        g = index.make_peptide_bloom_filter(first, k, alphabet, n_tables=4, tablesize=1e6)
        np.testing.assert_allclose(g.n_unique_kmers(), n, rtol=1e-3)
    r = CliRunner().invoke(index.cli, ["--tablesize", "1e6", first])
    assert r.exit_code == 0, r.exception
    default_name = "reference_data/index/Homo_sapiens.GRCh38.pep.first1000lines.alphabet-protein_ksize-7.bloomfilter.nodegraph"
    assert os.path.exists(default_name)
    g = index.load_nodegraph(default_name)
    assert g.ksize() == 7 and g.hashsizes() == [999983, 999979, 999961, 999959]
    os.remove(default_name)
    r = CliRunner().invoke(index.cli, ["--tablesize", "1e6", "--alphabet", "dayhoff", "--save-as", "ours/custom.nodegraph", first])
    assert r.exit_code == 0 and index.load_nodegraph("ours/custom.nodegraph").ksize() == 12
    r = CliRunner().invoke(index.cli, ["--alphabet", "not-a-molecule", first])
    assert isinstance(r.exception, ValueError)
    # a directory: non-sequence files (the .nodegraph goldens) are skipped silently
    g = index.make_peptide_bloom_filter((os.path.join("reference_data/index", p) for p in sorted(os.listdir("reference_data/index"))),
                                        7, "protein", n_tables=4, tablesize=1e6, index_dir="reference_data/index")
    np.testing.assert_allclose(g.n_unique_kmers(), 519609, rtol=1e-3)  # tests/test_index.py:71-82
    with pytest.raises(OSError):
        index.load_nodegraph(first)
    index.maybe_make_peptide_bloom_filter(
        "reference_data/index/Homo_sapiens.GRCh38.pep.subset.alphabet-protein_ksize-7.bloomfilter.nodegraph", 7, "protein",
        peptides_are_bloom_filter=True)


def test_columnar_outputs_equal_row_list_outputs(workdir, monkeypatch, tmp_path):
    """ScoreTable (arrays) vs CreateSaveSummary (python rows) on awkward read ids: commas, quotes, empty name,
    duplicated consecutive ids, duplicated non-consecutive ids."""
    import json as js

    from orpheum_b200 import constants_index, translate
    from orpheum_b200.create_save_summary import CreateSaveSummary

    monkeypatch.chdir(workdir)
    coding = "CGCTTGCTTAATACTGACATCAATAATATTAGGAAAATCGCAATATAACTGTAAATCCTGTTCTGTC"
    reads = [("plain", coding), ('with,comma "and quotes"', coding), ("dup", coding), ("dup", "ACGT" * 20), ("", "CAG" * 30),
             ("other", "C" * 50), ("dup", coding), ("tab\there", coding[:30])]
    fq = tmp_path / "awkward.fq"
    fq.write_text("".join(f"@{n}\n{s}\n+\n{'I' * len(s)}\n" for n, s in reads))
    args = dict(reads=[str(fq)], peptides="reference_data/index/Homo_sapiens.GRCh38.pep.subset.alphabet-protein_ksize-7.bloomfilter.nodegraph",
                peptide_ksize=7, save_peptide_bloom_filter=False, peptides_are_bloom_filter=True, jaccard_threshold=None,
                alphabet="protein", csv=False, parquet=False, json_summary=False, coding_nucleotide_fasta=None,
                noncoding_nucleotide_fasta=None, low_complexity_nucleotide_fasta=None, low_complexity_peptide_fasta=None,
                tablesize=constants_index.DEFAULT_MAX_TABLESIZE, n_tables=4, long_reads=False, verbose=False)
    t = translate.Translate(args)
    t.set_coding_scores_all_files()
    rows = t.coding_scores
    ref = CreateSaveSummary([str(fq)], str(tmp_path / "r.csv"), str(tmp_path / "r.parquet"), str(tmp_path / "r.json"), "f.nodegraph",
                            "protein", 7, 0.5, rows)
    ref.maybe_write_csv()
    ref.maybe_write_parquet()
    ref.maybe_write_json_summary()
    table = translate.Translate(args).score_all_files_columnar()
    assert len(table) == len(reads)
    table.write_csv(str(tmp_path / "c.csv"))
    table.write_parquet(str(tmp_path / "c.parquet"))
    table.write_json_summary(str(tmp_path / "c.json"), [str(fq)], "f.nodegraph", 7, 0.5)
    assert (tmp_path / "c.csv").read_text() == (tmp_path / "r.csv").read_text()
    assert js.load(open(tmp_path / "c.json")) == js.load(open(tmp_path / "r.json"))
    import pyarrow.parquet as pq

    a, b = pq.read_table(str(tmp_path / "c.parquet")), pq.read_table(str(tmp_path / "r.parquet"))
    assert a.schema.equals(b.schema) and a.to_pandas().equals(b.to_pandas())


def test_translate_cli_on_several_lanes_writes_the_same_files(workdir, golden_dir, monkeypatch):
    """The product multi-GPU path (Translate._replica_lanes -> orph_stream_open_multi): ORPHEUM_B200_DEVICES="0,0,0" runs
    three lanes (three private contexts, the filter replicated twice) on the one GPU of the test box; stdout, CSV, JSON summary
    and the FASTA files must be byte-identical to the single-lane run (which is pinned to the reference's golden outputs)."""
    from orpheum_b200 import build, translate

    build.build()
    run = next(m for m in manifest(golden_dir) if m["tag"] == "synth_protein7")
    monkeypatch.chdir(workdir)
    outputs = {}
    for label, devices in (("one", "0"), ("three", "0,0,0")):
        monkeypatch.setenv("ORPHEUM_B200_DEVICES", devices)
        argv = [a.replace("ref_cli/synth_protein7", f"ours/lanes_{label}") for a in run["argv"]]
        result = CliRunner().invoke(translate.cli, argv)
        assert result.exit_code == 0, result.exception
        files = {ext: open(os.path.join("ours", f"lanes_{label}" + ext), "rb").read()
                 for ext in (".csv", ".json", ".coding_nt.fa", ".noncoding_nt.fa", ".lowcplx_nt.fa", ".lowcplx_pep.fa")}
        outputs[label] = (result.output, files)
    assert outputs["one"][0] == outputs["three"][0] and len(outputs["one"][0]) > 0
    for ext, blob in outputs["one"][1].items():
        other = outputs["three"][1][ext]
        if ext == ".json":  # the summary names its own output-independent inputs only
            assert json.loads(blob) == json.loads(other)
        else:
            assert blob == other, ext
    assert outputs["three"][1][".csv"] == open(os.path.join("ref_cli", "synth_protein7.csv"), "rb").read()
    saved = os.path.splitext(run["peptides"])[0] + f".alphabet-{run['alphabet']}_ksize-{run['ksize']}.bloomfilter.nodegraph"
    if os.path.exists(saved):
        os.remove(saved)
