#!/usr/bin/env python
"""Generate the committed golden vectors under tests/golden/ by running the UNMODIFIED
reference (``/root/reference/orpheum``) through the stand-ins in oracle/shims.

Run in the build container only (the reference checkout does not exist on the GPU box):

    python oracle/gen_golden.py

Outputs
  tests/golden/reference_data/      verbatim copies of the reference's own test DATA fixtures
                                    (FASTQ/FASTA inputs, golden .nodegraph / CSV / FASTA outputs)
  tests/golden/ref_scores_*.json    per-read, per-frame rows from Translate.maybe_score_single_read
                                    (+ integer hit counts, stdout FASTA, the four optional FASTAs)
  tests/golden/ref_index_kats.json  n_unique_kmers / per-table popcounts / sha256 of .nodegraph files
                                    written by index.make_peptide_bloom_filter
  tests/golden/ref_cli/             CSV / JSON summary / stdout of the reference CLI on small inputs
  tests/golden/ref_misc_kats.json   hash, codon, complexity and alphabet known answers

Test infrastructure only.
"""
import contextlib
import hashlib
import io
import json
import math
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import refshim  # noqa: E402

refshim.install()

import orpheum.index as ref_index  # noqa: E402
import orpheum.sequence_encodings as ref_enc  # noqa: E402
import orpheum.translate as ref_translate  # noqa: E402
import orpheum.translate_single_seq as ref_tss  # noqa: E402
from click.testing import CliRunner  # noqa: E402
from sourmash.minhash import hash_murmur  # noqa: E402  (the pure-python stand-in)

from orpheum_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REFDATA = os.path.join(refshim.REFERENCE_ROOT, "tests", "data")
DATA_FILES = [
    "SRR306838_GSM752691_hsa_br_F_1_trimmed_subsampled_n22.fq",
    "SRR306838_GSM752691_hsa_br_F_1_trimmed_subsampled_n22.fq.gz",
    "SRR306838_GSM752691_hsa_br_F_1_trimmed_subsampled_first20.fq.gz",
    "empty_fasta.fasta",
    "low_complexity_nucleotides.fastq",
    "low_complexity_peptides.fastq",
    "index/Homo_sapiens.GRCh38.pep.first1000lines.fa",
    "index/Homo_sapiens.GRCh38.pep.subset.fa.gz",
    "index/Homo_sapiens.GRCh38.pep.subset.alphabet-protein_ksize-7.bloomfilter.nodegraph",
    "index/Homo_sapiens.GRCh38.pep.subset.alphabet-dayhoff_ksize-12.bloomfilter.nodegraph",
    "translate/SRR306838_GSM752691_hsa_br_F_1_trimmed_subsampled_n22__alphabet-protein_ksize-7.csv",
    "translate/SRR306838_GSM752691_hsa_br_F_1_trimmed_subsampled_n22__alphabet-dayhoff_ksize-12.csv",
    "translate/true_protein_coding.fasta",
]

SEQ = "CGCTTGCTTAATACTGACATCAATAATATTAGGAAAATCGCAATATAACTGTAAATCCTGTTCTGTC"
LOW_COMPLEXITY_SEQ = "CCCCCCCCCACCACCACCCCCCCCACCCCCCCCCCCCCCCCCCCCCCCCCCACCCCCCCACACACCCCCAACACCC"


def copy_reference_data():
    for rel in DATA_FILES:
        dst = os.path.join(GOLD, "reference_data", rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REFDATA, rel), dst)
        os.chmod(dst, 0o644)


def read_fasta(path):
    import screed

    with screed.open(path) as records:
        return [(r["name"], r["sequence"]) for r in records]


def proteome_arrays(fasta):
    seqs = [s for _, s in read_fasta(fasta) if "*" not in s and len(s) >= 120]
    residues = np.frombuffer("".join(seqs).encode(), dtype=np.uint8)
    offsets = np.zeros(len(seqs) + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum([len(s) for s in seqs])
    return residues, offsets


def edge_reads():
    """Hand-made reads for every branch and quirk."""
    n_read = SEQ[:40] + "N" + SEQ[41:]
    reads = [
        ("gold-seq", SEQ),
        ("gold-lowcomplexity", LOW_COMPLEXITY_SEQ),
        ("one-base", "A"),
        ("two-bases", "AC"),
        ("three-bases", "ACG"),
        ("all-lower", SEQ.lower() + "acgtacgta"),
        ("mixed-case", SEQ[:30] + SEQ[30:45].lower() + SEQ[45:]),
        ("xxN-codons", "GCNGCNACNCTNCGNGTNGGNTCNCCNNAA" + SEQ[:44]),
        ("gold-2_N wobble", n_read),
        ("len21", "GCTGCAGATGAAAAGCTGGTT"),
        ("len24", "GCTGCAGATGAAAAGCTGGTTCAG"),
        ("cag-repeat with spaces in name", "CAG" * 30),
        ("poly-c", "C" * 75),
        ("acgt-period", "ACGT" * 40),
        ("iupac", "ACGTRYKMSWBDHVN" * 6 + SEQ),
        ("n-first", "N" + SEQ),
        ("n-run", SEQ[:20] + "NNNNNN" + SEQ[20:]),
        ("at-repeat", "AT" * 60),
        ("gc-dinuc-long", "GCA" * 70),
    ]
    return reads


def synthetic_reads(residues, offsets, ksize, n=260, seed=77):
    buf, offs = synth.make_reads_mixed(n, residues, offsets, ksize=ksize, seed=seed, min_len=30, max_len=300)
    out = []
    for i in range(n):
        s = buf[int(offs[i]):int(offs[i + 1])].tobytes().decode()
        if len(s) == 0:
            continue
        if i % 37 == 5:
            s = s.lower()
        out.append((f"synth{i} len={len(s)}", s))
    return out


def nan_to_none(x):
    if isinstance(x, float) and math.isnan(x):
        return None
    if isinstance(x, (np.integer,)):
        return int(x)
    if isinstance(x, (np.floating,)):
        return None if math.isnan(float(x)) else float(x)
    return x


def score_vectors(alphabet, ksize, fasta, tablesize, reads, jaccard_threshold=None, tmp="/tmp/orpheum_golden"):
    """Run the unmodified Translate on `reads`; capture rows, integer hits, stdout and FASTAs."""
    os.makedirs(tmp, exist_ok=True)
    fas = {
        "coding_nucleotide_fasta": os.path.join(tmp, "coding_nt.fa"),
        "noncoding_nucleotide_fasta": os.path.join(tmp, "noncoding_nt.fa"),
        "low_complexity_nucleotide_fasta": os.path.join(tmp, "lowcplx_nt.fa"),
        "low_complexity_peptide_fasta": os.path.join(tmp, "lowcplx_pep.fa"),
    }
    args = dict(
        reads=[], peptides=fasta, peptide_ksize=ksize, save_peptide_bloom_filter=os.path.join(tmp, "f.nodegraph"),
        peptides_are_bloom_filter=False, jaccard_threshold=jaccard_threshold, alphabet=alphabet, csv=False,
        parquet=False, json_summary=False, tablesize=tablesize, n_tables=4, long_reads=False, verbose=False, **fas,
    )
    tr = ref_translate.Translate(args)
    hits_log = []
    orig = tr.score_single_translation

    def wrapped(translation):
        frac, n = orig(translation)
        # recompute the integer numerator exactly as translate.py:186-191 does
        encoded = ref_enc.encode_peptide(translation, tr.alphabet)
        kmers = set(encoded[i:i + tr.peptide_ksize] for i in range(len(encoded) - tr.peptide_ksize + 1))
        h = sum(1 for km in kmers if tr.peptide_bloom_filter.get(hash_murmur(km)) > 0)
        assert len(kmers) == n and h / n == frac
        hits_log.append(h)
        return frac, n

    tr.score_single_translation = wrapped
    tr.maybe_open_fastas()
    out_reads = []
    stdout = io.StringIO()
    with contextlib.redirect_stdout(stdout):
        for name, seq in reads:
            hits_log.clear()
            rows = tr.maybe_score_single_read(name, seq)
            hits = list(hits_log)
            rec = []
            hi = 0
            for j, n_kmers, cat, frame in rows:
                scored = not (isinstance(n_kmers, float) and math.isnan(n_kmers))
                h = None
                if scored:
                    h = hits[hi]
                    hi += 1
                rec.append([nan_to_none(j), nan_to_none(n_kmers), cat, frame, h])
            assert hi == len(hits)
            out_reads.append({"name": name, "seq": seq, "rows": rec})
    tr.maybe_close_fastas()
    g = tr.peptide_bloom_filter
    result = {
        "alphabet": alphabet,
        "ksize": ksize,
        "tablesize": tablesize,
        "n_tables": 4,
        "jaccard_threshold": tr.jaccard_threshold,
        "peptides": os.path.relpath(fasta, os.path.join(GOLD, "reference_data")),
        "filter_popcounts": [sum(bin(b).count("1") for b in t) for t in g._tables],
        "filter_sizes": g.hashsizes(),
        "stdout": stdout.getvalue(),
        "fastas": {k: open(v).read() for k, v in fas.items()},
        "reads": out_reads,
    }
    return result


def index_kats(fasta):
    out = []
    combos = [("protein", 7), ("dayhoff", 7), ("dayhoff", 11), ("dayhoff", 12), ("hydrophobic-polar", 31),
              ("hydrophobic-polar", 21), ("hydrophobic-polar", 7), ("hp", 31), ("botvinnik", 11), ("aa9", 10),
              ("gbmr4", 16), ("sdm12", 9), ("hsdm17", 8), ("protein", 13), ("dayhoff", 17), ("protein", 1)]
    for alphabet, k in combos:
        g = ref_index.make_peptide_bloom_filter(fasta, k, alphabet, n_tables=4, tablesize=1e6)
        path = "/tmp/orpheum_golden_idx.nodegraph"
        g.save(path)
        out.append({
            "alphabet": alphabet, "ksize": k, "tablesize": 1000000, "n_tables": 4,
            "n_unique_kmers": g.n_unique_kmers(),
            "popcounts": [sum(bin(b).count("1") for b in t) for t in g._tables],
            "sha256": hashlib.sha256(open(path, "rb").read()).hexdigest(),
        })
    # a different geometry
    g = ref_index.make_peptide_bloom_filter(fasta, 7, "protein", n_tables=2, tablesize=30011)
    path = "/tmp/orpheum_golden_idx.nodegraph"
    g.save(path)
    out.append({"alphabet": "protein", "ksize": 7, "tablesize": 30011, "n_tables": 2,
                "n_unique_kmers": g.n_unique_kmers(),
                "popcounts": [sum(bin(b).count("1") for b in t) for t in g._tables],
                "sizes": g.hashsizes(),
                "sha256": hashlib.sha256(open(path, "rb").read()).hexdigest()})
    return out


def misc_kats():
    keys = ["", "A", "ACG", "TEQDLQL", "QSSSPEF", "LDPPYSR", "bbbdbfecdacb", "phpphhhppppphpphhhppppphpphhhpp",
            "ABCDEFGHIJKLMNOP", "ABCDEFGHIJKLMNOPQ", "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghi", "XXXXXXX"]
    codons = {}
    for a in "ACGTNacgtRY":
        for b in "ACGTN":
            for c in "ACGTNn":
                cod = a + b + c
                codons[cod] = ref_tss.TranslateSingleSeq._translate_codon(cod)
    alph = {}
    for name in list(ref_enc.PEPTIDE_ENCODINGS) + list(ref_enc.PROTEIN_LIKE):
        alph[name] = ref_enc.encode_peptide("ACDEFGHIKLMNPQRSTVWYXUBZOJ*acdxy", name)
    six = {s: {str(k): v for k, v in ref_tss.TranslateSingleSeq(s, False).six_frame_translation().items()}
           for s in [SEQ, SEQ[:-1], SEQ[:-2], "ACGTNNACGATCGATNCGAT", "AC", "ACG", "ACGT", "ACGTA", "acgtacgtagc"]}
    return {
        "hash_murmur": {k: hash_murmur(k) for k in keys},
        "codons": codons,
        "alphabets": alph,
        "best_ksizes": dict(ref_enc.BEST_KSIZES),
        "six_frame": six,
        "fastp_complexity": {s: ref_translate.compute_fastp_complexity(s) for s in [SEQ, LOW_COMPLEXITY_SEQ, "A", "AC", "AAc"]},
        "kmer_complexity": {"seq7": ref_translate.compute_kmer_complexity(SEQ, 7),
                            "low7": ref_translate.compute_kmer_complexity(LOW_COMPLEXITY_SEQ, 7)},
        "thresholds": {a: ref_translate.get_jaccard_threshold(None, a)
                       for a in ["protein", "dayhoff", "dayhoff6", "hp", "hydrophobic-polar", "hp2", "botvinnik", ""]},
        "low_complexity_categories": dict(ref_translate.constants_translate.LOW_COMPLEXITY_CATEGORIES),
        "table_sizes": {str(int(x)): __import__("khmer").get_n_primes_near_x(4, x)
                        for x in (1e4, 1e6, 1e8, 1e9, 4e9, 30011, 100)},
        "fpr": {"per_translation_14_4e7": ref_index.per_translation_false_positive_rate(14, 4e7),
                "per_read_60_7": ref_index.per_read_false_positive_coding_rate(60, 7)},
    }


def cli_vectors():
    """Reference CLI end to end on (a) the 23 golden reads and (b) a synthetic FASTQ with N / short reads."""
    out_dir = os.path.join(GOLD, "ref_cli")
    os.makedirs(out_dir, exist_ok=True)
    data = os.path.join(GOLD, "reference_data")
    residues, offsets = proteome_arrays(os.path.join(data, "index", "Homo_sapiens.GRCh38.pep.first1000lines.fa"))
    fq = os.path.join(out_dir, "synth_mixed.fq")
    reads = synthetic_reads(residues, offsets, 7, n=150, seed=91) + edge_reads()[3:]
    with open(fq, "w") as f:
        for name, seq in reads:
            f.write(f"@{name}\n{seq}\n+\n{'I' * len(seq)}\n")
    runs = [
        ("n22_protein7", "protein", 7, "index/Homo_sapiens.GRCh38.pep.subset.alphabet-protein_ksize-7.bloomfilter.nodegraph",
         os.path.join(data, "SRR306838_GSM752691_hsa_br_F_1_trimmed_subsampled_n22.fq"), True),
        ("n22_dayhoff12", "dayhoff", 12, "index/Homo_sapiens.GRCh38.pep.subset.alphabet-dayhoff_ksize-12.bloomfilter.nodegraph",
         os.path.join(data, "SRR306838_GSM752691_hsa_br_F_1_trimmed_subsampled_n22.fq"), True),
        ("synth_protein7", "protein", 7, "index/Homo_sapiens.GRCh38.pep.first1000lines.fa", fq, False),
        ("synth_hp21", "hp", 21, "index/Homo_sapiens.GRCh38.pep.first1000lines.fa", fq, False),
    ]
    manifest = []
    cwd = os.getcwd()
    os.chdir(GOLD)  # so that the recorded filename column is a stable relative path
    try:
        for tag, alphabet, k, pep, reads_path, is_filter in runs:
            rel_reads = os.path.relpath(reads_path, GOLD)
            rel_pep = os.path.join("reference_data", pep)
            prefix = os.path.join("ref_cli", tag)
            argv = ["--peptide-ksize", str(k), "--alphabet", alphabet, "--tablesize", "1e6",
                    "--csv", prefix + ".csv", "--json-summary", prefix + ".json",
                    "--coding-nucleotide-fasta", prefix + ".coding_nt.fa",
                    "--noncoding-nucleotide-fasta", prefix + ".noncoding_nt.fa",
                    "--low-complexity-nucleotide-fasta", prefix + ".lowcplx_nt.fa",
                    "--low-complexity-peptide-fasta", prefix + ".lowcplx_pep.fa"]
            if is_filter:
                argv.append("--peptides-are-bloom-filter")
            else:
                argv += ["--save-peptide-bloom-filter"]
            argv += [rel_pep, rel_reads]
            result = CliRunner().invoke(ref_translate.cli, argv)
            assert result.exit_code == 0, (tag, result.exception)
            with open(prefix + ".stdout.txt", "w") as f:
                f.write(result.output)
            manifest.append({"tag": tag, "alphabet": alphabet, "ksize": k, "peptides": rel_pep, "reads": rel_reads,
                             "peptides_are_bloom_filter": is_filter, "argv": argv})
            # the filter the CLI saved next to the (copied) peptide fasta is not a fixture
            if not is_filter:
                saved = os.path.splitext(rel_pep)[0] + f".alphabet-{alphabet}_ksize-{k}.bloomfilter.nodegraph"
                if os.path.exists(saved):
                    os.remove(saved)
    finally:
        os.chdir(cwd)
    with open(os.path.join(out_dir, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)


def long_reads(residues, offsets, seed=123):
    """Reads beyond the 6144-nt bound of the warp-per-read byte path (block-per-read kernel): long open reading frames,
    random sequence, a tandem repeat, N bytes and a soft-masked stretch."""
    rng = np.random.default_rng(seed)
    codon = {"A": "GCT", "R": "CGT", "N": "AAC", "D": "GAC", "C": "TGC", "Q": "CAG", "E": "GAA", "G": "GGT", "H": "CAC",
             "I": "ATC", "L": "CTG", "K": "AAA", "M": "ATG", "F": "TTC", "P": "CCG", "S": "TCT", "T": "ACC", "W": "TGG",
             "Y": "TAC", "V": "GTT"}

    def orf(n_codons):
        out = []
        while len(out) < n_codons:
            p = int(rng.integers(0, len(offsets) - 1))
            seq = residues[int(offsets[p]):int(offsets[p + 1])].tobytes().decode()
            out.extend(codon.get(a, "GCT") for a in seq)
        return "".join(out[:n_codons])

    rand = lambda n: "".join(rng.choice(list("ACGT"), size=n))
    with_n = list(orf(2400))
    for pos in rng.integers(0, len(with_n), 25):
        with_n[int(pos)] = "N"
    masked = orf(2300)
    masked = masked[:3000] + masked[3000:3300].lower() + masked[3300:]
    return [
        ("long-orf 6150", orf(2050)),
        ("long-orf frame2 9001", "G" + orf(3000)),
        ("long-random 7000", rand(7000)),
        ("long-tandem 3x2100", orf(700) * 3),
        ("long-with-N 7200", "".join(with_n)),
        ("long-softmasked 6900", masked),
        ("long-lowcomplexity 8000", "A" * 4000 + "C" * 4000),
        ("long-repeat-peptide 6300", "GCTGAAAAGCTGGTT" * 420),
    ]


def long_vectors():
    data = os.path.join(GOLD, "reference_data")
    first1000 = os.path.join(data, "index", "Homo_sapiens.GRCh38.pep.first1000lines.fa")
    residues, offsets = proteome_arrays(first1000)
    vec = score_vectors("protein", 7, first1000, 1000000, long_reads(residues, offsets))
    with open(os.path.join(GOLD, "ref_scores_long_protein_k7.json"), "w") as f:
        json.dump(vec, f, separators=(",", ":"))
    cats = {}
    for r in vec["reads"]:
        for row in r["rows"]:
            cats[row[2]] = cats.get(row[2], 0) + 1
    print("long", len(vec["reads"]), "reads", cats, [max((row[1] or 0) for row in r["rows"]) for r in vec["reads"]])


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "long":  # only the long-read vector (added later; the others stay as they are)
        long_vectors()
        return
    os.makedirs(GOLD, exist_ok=True)
    copy_reference_data()
    data = os.path.join(GOLD, "reference_data")
    first1000 = os.path.join(data, "index", "Homo_sapiens.GRCh38.pep.first1000lines.fa")
    residues, offsets = proteome_arrays(first1000)
    for alphabet, k, thr in [("protein", 7, None), ("dayhoff", 12, None), ("hp", 31, None),
                             ("hydrophobic-polar", 21, None), ("botvinnik", 11, None), ("protein", 9, 0.3),
                             ("hsdm17", 8, None)]:
        reads = edge_reads() + synthetic_reads(residues, offsets, k)
        vec = score_vectors(alphabet, k, first1000, 1000000, reads, thr)
        with open(os.path.join(GOLD, f"ref_scores_{alphabet}_k{k}.json"), "w") as f:
            json.dump(vec, f, separators=(",", ":"))
        cats = {}
        for r in vec["reads"]:
            for row in r["rows"]:
                cats[row[2]] = cats.get(row[2], 0) + 1
        print(alphabet, k, len(vec["reads"]), "reads", cats)
    with open(os.path.join(GOLD, "ref_index_kats.json"), "w") as f:
        json.dump(index_kats(first1000), f, indent=1)
    with open(os.path.join(GOLD, "ref_misc_kats.json"), "w") as f:
        json.dump(misc_kats(), f, indent=1)
    cli_vectors()
    long_vectors()


if __name__ == "__main__":
    main()
