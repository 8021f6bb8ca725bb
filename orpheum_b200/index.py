"""``orpheum index``: build the reference-proteome Bloom filter on the GPU.

Mirror of the reference's ``orpheum/index.py`` (same function names, arguments and error behaviour):
``make_peptide_bloom_filter`` (``:92-123``), ``maybe_make_peptide_bloom_filter`` (``:147-195``),
``maybe_save_peptide_bloom_filter`` (``:198-220``), ``load_nodegraph`` (``:70-77``),
``get_peptide_ksize`` (``:325-335``), the FPR helpers (``:18-67``) and the click command
(``:223-322``).  The per-record inner loop (skip records with ``*``, re-encode, hash every k-mer, set
the bits) is one C call per batch of records: ``orph_filter_add_peptides``.
"""
import math
import os

import click
import numpy as np

from . import constants_index, fastx
from .gpu import GpuNodegraph
from .sequence_encodings import ALPHABET_SIZES, BEST_KSIZES, encode_lut


def per_translation_false_positive_rate(
    n_kmers_in_translation,
    n_total_kmers=1e7,
    n_hash_functions=constants_index.DEFAULT_N_TABLES,
    tablesize=constants_index.DEFAULT_MAX_TABLESIZE,
):
    """P(every k-mer of one translation is a Bloom false positive) (index.py:18-37)."""
    per_kmer = math.pow(-math.expm1(-n_hash_functions * n_total_kmers / tablesize), n_hash_functions)
    return math.pow(per_kmer, n_kmers_in_translation)


def per_read_false_positive_coding_rate(
    read_length,
    peptide_ksize,
    n_total_kmers=4e7,
    alphabet="protein",
    n_hash_functions=constants_index.DEFAULT_N_TABLES,
    tablesize=constants_index.DEFAULT_MAX_TABLESIZE,
):
    """Sum over the six frames of the per-translation rate (index.py:40-67)."""
    n_kmers = min(n_total_kmers, ALPHABET_SIZES[alphabet] ** peptide_ksize)
    total = 0
    for frame in range(3):
        n_in_translation = math.floor((read_length - frame) / 3) - peptide_ksize + 1
        total += 2 * per_translation_false_positive_rate(
            n_in_translation, n_kmers, n_hash_functions=n_hash_functions, tablesize=tablesize
        )
    return total


def load_nodegraph(path, ctx=None):
    """khmer.load_nodegraph / khmer.Nodegraph.load: OSError unless `path` is an OXLI nodegraph (index.py:70-77)."""
    return GpuNodegraph.load(path, ctx=ctx)


def maybe_read_peptide_file(peptide_file):
    """Records of a sequence file, or [] for anything that is not one (index.py:80-89)."""
    try:
        return fastx.open(peptide_file)
    except (ValueError, UnicodeDecodeError):
        return []


def get_peptide_ksize(molecule, peptide_ksize=None):
    if peptide_ksize is None:
        try:
            peptide_ksize = BEST_KSIZES[molecule]
        except KeyError:
            raise ValueError(
                f"{molecule} does not have a default k-mer size! "
                f"Only 'protein', 'hydrophobic-polar', or 'dayhoff' have a default protein ksize"
            )
    return peptide_ksize


def _record_batches(records, max_residues=1 << 26):
    seqs, total = [], 0
    for record in records:
        s = record["sequence"].encode("utf-8", "surrogateescape")
        seqs.append(s)
        total += len(s)
        if total >= max_residues:
            yield seqs
            seqs, total = [], 0
    if seqs:
        yield seqs


def make_peptide_bloom_filter(
    peptides,
    peptide_ksize,
    molecule,
    n_tables=constants_index.DEFAULT_N_TABLES,
    tablesize=constants_index.DEFAULT_MAX_TABLESIZE,
    index_dir=None,
    ctx=None,
    count_unique=True,
):
    """Bloom filter of every peptide k-mer of `peptides` (a file, or an iterable of files with index_dir).
    count_unique=False skips the bookkeeping behind ``n_unique_kmers()`` (the saved file is identical)."""
    lut = encode_lut(molecule)  # ValueError for an unknown alphabet, like encode_peptide
    bloom = GpuNodegraph(peptide_ksize, int(tablesize), n_tables=n_tables, ctx=ctx)
    if not index_dir:
        peptides = [peptides]
    for peptide_fasta in peptides:
        records = maybe_read_peptide_file(peptide_fasta)
        try:
            for seqs in _record_batches(records):
                offsets = np.zeros(len(seqs) + 1, dtype=np.uint64)
                offsets[1:] = np.cumsum([len(s) for s in seqs])
                residues = np.frombuffer(b"".join(seqs), dtype=np.uint8) if offsets[-1] else np.zeros(1, np.uint8)
                bloom.add_peptides(residues, offsets, lut, count_unique=count_unique)
        finally:
            if hasattr(records, "close"):
                records.close()
    return bloom


def maybe_make_peptide_bloom_filter(
    peptides,
    peptide_ksize,
    molecule,
    peptides_are_bloom_filter,
    n_tables=constants_index.DEFAULT_N_TABLES,
    tablesize=constants_index.DEFAULT_MAX_TABLESIZE,
    index_dir=None,
    ctx=None,
):
    if peptides_are_bloom_filter:
        bloom = load_nodegraph(peptides, ctx=ctx)
        if peptide_ksize is not None and peptide_ksize != bloom.ksize():
            raise ValueError(
                f"Given peptide ksize ({peptide_ksize}) and ksize found in bloom filter ({bloom.ksize()}) are notequal"
            )
        return bloom
    peptide_ksize = get_peptide_ksize(molecule, peptide_ksize)
    return make_peptide_bloom_filter(
        peptides, peptide_ksize, molecule=molecule, n_tables=n_tables, tablesize=tablesize, index_dir=index_dir, ctx=ctx
    )


def maybe_save_peptide_bloom_filter(peptides, peptide_bloom_filter, molecule, save_peptide_bloom_filter, index_dir=None):
    """Save under the given name, or under the reference's default name (index.py:198-220)."""
    if not save_peptide_bloom_filter:
        return None
    if isinstance(save_peptide_bloom_filter, str):
        filename = save_peptide_bloom_filter
    else:
        suffix = f".alphabet-{molecule}_ksize-{peptide_bloom_filter.ksize()}.bloomfilter.nodegraph"
        if not index_dir:
            filename = os.path.splitext(peptides)[0] + suffix
        else:
            filename = os.path.join(index_dir, os.path.basename(index_dir) + suffix)
    peptide_bloom_filter.save(filename)
    return filename


@click.command()
@click.argument("peptides")
@click.option("--peptide-ksize", default=None, type=int,
              help="K-mer size of the peptide sequence to use. Defaults for different molecules are, "
                   f"protein: {constants_index.DEFAULT_PROTEIN_KSIZE}, dayhoff: {constants_index.DEFAULT_DAYHOFF_KSIZE},"
                   f" hydrophobic-polar: {constants_index.DEFAULT_HP_KSIZE}")
@click.option("--alphabet", "--molecule", default="protein",
              help="The type of amino acid alphabet/encoding to use. Default is 'protein', but 'dayhoff' or "
                   "'hydrophobic-polar' can be used")
@click.option("--save-as", default=None,
              help="If provided, save peptide bloom filter as this filename. Otherwise, add ksize and alphabet name "
                   "to input filename.")
@click.option("--tablesize", type=constants_index.BASED_INT, default="1e8", help="Size of the bloom filter table to use")
@click.option("--n-tables", type=int, default=constants_index.DEFAULT_N_TABLES, help="Size of the bloom filter table to use")
@click.option("--index-from-dir", is_flag=True, default=False,
              help="Build peptide bloom filter from a directory of peptide fasta files")
def cli(peptides, peptide_ksize=None, alphabet="protein", save_as=None, tablesize=constants_index.DEFAULT_MAX_TABLESIZE,
        n_tables=constants_index.DEFAULT_N_TABLES, index_from_dir=False):
    """Make a peptide bloom filter for your peptides"""
    index_dir = ""
    if index_from_dir:
        index_dir = peptides
        peptides = (os.path.join(index_dir, p) for p in os.listdir(index_dir))
    peptide_ksize = get_peptide_ksize(alphabet, peptide_ksize)
    bloom = make_peptide_bloom_filter(peptides, peptide_ksize, alphabet, n_tables=n_tables, tablesize=tablesize,
                                      index_dir=index_dir, count_unique=False)  # the CLI never reports the count
    maybe_save_peptide_bloom_filter(peptides, bloom, alphabet,
                                    save_peptide_bloom_filter=save_as if save_as is not None else True,
                                    index_dir=index_dir)
