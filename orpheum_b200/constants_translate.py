"""Strings and thresholds of the translate path (mirror of ``orpheum/constants_translate.py``).

Category texts are part of the output contract (CSV / parquet / JSON summary); the numeric codes are
the ones the kernels write into ``orph_read_result.category``.
"""
from collections import namedtuple

from .sequence_encodings import ALPHABET_ALIASES

DEFAULT_JACCARD_THRESHOLD = 0.5
DEFAULT_HP_JACCARD_THRESHOLD = 0.8
COMPLEXITY_THRESHOLD = 0.3

SEQTYPE_TO_ANNOUNCEMENT = {
    "noncoding_nucleotide": "nucleotide sequence from reads WITHOUT matches to protein-coding peptides",
    "coding_nucleotide": "nucleotide sequence from reads WITH protein-coding translation frame nucleotides",
    "low_complexity_nucleotide": "nucleotide sequence from low complexity (low entropy) reads",
    "low_complexity_peptide": "peptide sequence from low complexity (low entropy) translated reads",
}

SCORING_DF_COLUMNS = ["read_id", "jaccard_in_peptide_db", "n_kmers", "category", "translation_frame", "filename"]

# alias -> "Low complexity peptide in {canonical} alphabet"   (constants_translate.py:27-34)
LOW_COMPLEXITY_CATEGORIES = {
    alias: f"Low complexity peptide in {canonical} alphabet"
    for canonical, aliases in ALPHABET_ALIASES.items()
    for alias in aliases
}

PROTEIN_CODING_CATEGORIES = {
    "too_short_peptide": "Translation is shorter than peptide k-mer size + 1",
    "stop_codons": "Translation frame has stop codon(s)",
    "coding": "Coding",
    "non_coding": "Non-coding",
    "low_complexity_nucleotide": "Low complexity nucleotide",
    "too_short_nucleotide": "Read length was shorter than 3 * peptide k-mer size",
}

SingleReadScore = namedtuple(
    "SingleReadScore", ["max_fraction_kmers_in_peptide_db", "max_n_kmers", "special_case", "translation_frame"]
)

FRAMES = (1, 2, 3, -1, -2, -3)


def category_text(code, alphabet):
    """orph_category code -> the reference's category string for this alphabet (KeyError for alphabets
    the reference has no low-complexity label for, exactly like translate.py:254-256)."""
    if code == 0:
        return PROTEIN_CODING_CATEGORIES["stop_codons"]
    if code == 1:
        return PROTEIN_CODING_CATEGORIES["too_short_peptide"]
    if code == 2:
        return LOW_COMPLEXITY_CATEGORIES[alphabet]
    if code == 3:
        return PROTEIN_CODING_CATEGORIES["coding"]
    if code == 4:
        return PROTEIN_CODING_CATEGORIES["non_coding"]
    if code == 5:
        # the reference labels fastp-low-complexity reads this way (translate.py:301-331)
        return PROTEIN_CODING_CATEGORIES["too_short_nucleotide"]
    raise ValueError(f"unknown category code {code}")
