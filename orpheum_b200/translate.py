"""``orpheum translate``: partition reads into coding / non-coding / low-complexity bins.

Host-side mirror of the reference's ``orpheum/translate.py``: same class, method names, arguments,
outputs and error behaviour, but ``score_reads_per_file`` (``translate.py:367-372``) parses a batch of
reads, makes ONE call into the CUDA library for the whole batch (``orph_score_reads``: complexity test,
six-frame translation, alphabet re-encoding, distinct k-mers, MurmurHash, Bloom probes, classification)
and only expands the integer results into the reference's rows and FASTA writes here.  There is no
host implementation of the scoring; without the CUDA library and a B200 the calls raise.
"""
import os
import sys

import click
import numpy as np

from . import _lib, constants_index, constants_translate, fastx
from .constants_translate import FRAMES, SingleReadScore, category_text
from .create_save_summary import CreateSaveSummary
from .gpu import ScoringStream, concat_reads, format_fasta, score_reads, translate_frames, translate_frames_raw
from .index import maybe_make_peptide_bloom_filter, maybe_save_peptide_bloom_filter
from .sequence_encodings import encode_lut, encode_peptide
from .translate_single_seq import TranslateSingleSeq  # noqa: F401  (re-exported: part of the reference's surface)

NAN = np.nan
_TRACE = bool(os.environ.get("ORPHEUM_B200_CLI_TRACE"))
_T0 = [0.0]


def _trace(what):
    """ORPHEUM_B200_CLI_TRACE=1: phase timestamps of the CLI on stderr (seconds since the command started)."""
    if _TRACE:
        import time

        now = time.perf_counter()
        if not _T0[0]:
            _T0[0] = now
        print("[orpheum_b200] %8.3f s  %s" % (now - _T0[0], what), file=sys.stderr)


def get_jaccard_threshold(jaccard_threshold, alphabet):
    """0.8 for the exact names 'hp' / 'hydrophobic-polar', else 0.5 (translate.py:31-37)."""
    if jaccard_threshold is None:
        if alphabet in ("hp", "hydrophobic-polar"):
            return constants_translate.DEFAULT_HP_JACCARD_THRESHOLD
        return constants_translate.DEFAULT_JACCARD_THRESHOLD
    return jaccard_threshold


def validate_jaccard(ctx, param, value):
    """click callback: --jaccard-threshold must lie in [0, 1] (translate.py:40-53)."""
    if value is None:
        return value
    try:
        jaccard = float(value)
        if not 0 <= jaccard <= 1:
            raise ValueError(value)
        return jaccard
    except ValueError:
        raise click.BadParameter(
            "--jaccard-threshold needs to be a number between 0 and 1, but {} was provided".format(value)
        )


# Single-sequence helpers of the reference's public surface (translate.py:56-111).  The batch path does
# not use them: the same arithmetic runs inside prefilter_kernel / score_*_kernel.
def compute_fastp_complexity(seq):
    return sum(1 for a, b in zip(seq, seq[1:]) if a != b) / len(seq)


def evaluate_is_fastp_low_complexity(seq):
    return compute_fastp_complexity(seq) < constants_translate.COMPLEXITY_THRESHOLD


def compute_kmer_complexity(seq, ksize):
    return (len(seq) - ksize + 1) / 2


def evaluate_is_kmer_low_complexity(seq, ksize):
    n_kmers = len(set(seq[i:i + ksize] for i in range(len(seq) - ksize + 1)))
    return n_kmers <= compute_kmer_complexity(seq, ksize)


def write_fasta(file_handle, description, sequence):
    file_handle.write(">{}\n{}\n".format(description, sequence))


class Translate:
    def __init__(self, args):
        self.args = args
        for key, value in args.items():
            setattr(self, key, value)
        if self.long_reads:
            raise NotImplementedError("Not implemented! ... yet :)")
        self.jaccard_threshold = get_jaccard_threshold(self.jaccard_threshold, self.alphabet)
        self.peptide_bloom_filter = maybe_make_peptide_bloom_filter(
            self.peptides, self.peptide_ksize, self.alphabet, self.peptides_are_bloom_filter,
            n_tables=self.n_tables, tablesize=self.tablesize,
        )
        self.verbose = True  # the reference forces this (translate.py:133); its logger is at ERROR level
        if not self.peptides_are_bloom_filter:
            self.peptide_bloom_filter_filename = maybe_save_peptide_bloom_filter(
                self.peptides, self.peptide_bloom_filter, self.alphabet, self.save_peptide_bloom_filter
            )
        else:
            self.peptide_bloom_filter_filename = self.peptides
        self.peptide_ksize = self.peptide_bloom_filter.ksize()
        self.nucleotide_ksize = 3 * self.peptide_ksize
        self._lut = encode_lut(self.alphabet)
        self.file_handles = {k: None for k in ("noncoding_nucleotide", "coding_nucleotide", "low_complexity_nucleotide",
                                               "low_complexity_peptide")}

    # -- FASTA outputs ---------------------------------------------------------------------------------
    def maybe_write_fasta(self, file_handle, description, sequence):
        if file_handle is not None:
            write_fasta(file_handle, description, sequence)

    def open_and_announce(self, filename, seqtype):
        # the reference announces through a logger that sits at ERROR level: nothing is printed
        return open(filename, "w", encoding=fastx.TEXT_CODEC[0], errors=fastx.TEXT_CODEC[1])

    def maybe_open_fastas(self):
        self.fastas = {
            "noncoding_nucleotide": self.noncoding_nucleotide_fasta,
            "coding_nucleotide": self.coding_nucleotide_fasta,
            "low_complexity_nucleotide": self.low_complexity_nucleotide_fasta,
            "low_complexity_peptide": self.low_complexity_peptide_fasta,
        }
        self.file_handles = {
            seqtype: (self.open_and_announce(fasta, seqtype) if fasta is not None else None)
            for seqtype, fasta in self.fastas.items()
        }

    def maybe_close_fastas(self):
        for handle in self.file_handles.values():
            if handle is not None:
                handle.close()

    def get_jaccard_threshold(self):
        return get_jaccard_threshold(self.jaccard_threshold, self.alphabet)

    # -- GPU scoring -------------------------------------------------------------------------------------
    def score_batch(self, sequences):
        """ONE C call for a list of read sequences -> structured array of ``orph_read_result``."""
        flat, offsets = concat_reads(sequences)
        return score_reads(self.peptide_bloom_filter, flat, offsets, self._lut, self.jaccard_threshold)

    def score_single_translation(self, translation):
        """(fraction of distinct k-mers in the filter, n distinct k-mers) of one peptide (translate.py:182-206);
        hashing and probing run on the GPU."""
        encoded = str(encode_peptide(translation, self.alphabet))
        k = self.peptide_ksize
        kmers = list(set(encoded[i:i + k] for i in range(len(encoded) - k + 1)))
        hashes = self.peptide_bloom_filter.ctx.hash_kmers(kmers)
        n_in_db = int(self.peptide_bloom_filter.get_many(hashes).sum())
        return n_in_db / len(kmers), len(kmers)

    def _emitted_peptides(self, flat, offsets, results):
        """Peptides of every frame the reference would print (coding -> stdout, low-complexity -> FASTA),
        translated on the GPU in one call: {(read index, frame key): peptide}."""
        cat = results["category"]
        reads_idx, frames_idx = np.nonzero((cat == _lib.CAT_CODING) | (cat == _lib.CAT_LOW_COMPLEXITY_PEPTIDE))
        if len(reads_idx) == 0:
            return {}
        keys = np.asarray(FRAMES, dtype=np.int8)[frames_idx]
        peptides = translate_frames(self.peptide_bloom_filter.ctx, flat, offsets, reads_idx, keys)
        return {(int(r), int(k)): p for r, k, p in zip(reads_idx, keys, peptides)}

    def _expand(self, description, sequence, result, peptides, index=0):
        """orph_read_result of one read -> the six SingleReadScore tuples, with the reference's side effects
        (stdout FASTA for coding frames, optional FASTA files), in frame order 1,2,3,-1,-2,-3."""
        scores = []
        for f, frame in enumerate(FRAMES):
            code = int(result["category"][f])
            text = category_text(code, self.alphabet)
            if code in (0, 1, 5):
                scores.append(SingleReadScore(NAN, NAN, text, frame))
                continue
            n_kmers = int(result["n_kmers"][f])
            if code == 2:
                # the reference's header literally contains the unformatted "{}.format(frame)" (translate.py:247)
                self.maybe_write_fasta(self.file_handles["low_complexity_peptide"],
                                       description + " translation_frame: {}.format(frame)",
                                       peptides[(index, frame)])
                scores.append(SingleReadScore(NAN, n_kmers, text, frame))
                continue
            jaccard = int(result["n_hits"][f]) / n_kmers  # the one IEEE division (translate.py:204)
            seqname = ("{}__translation_frame:{}".format(description, frame) + "__jaccard:{}".format(jaccard)).replace(" ", "__")
            if code == 3:
                write_fasta(sys.stdout, seqname, peptides[(index, frame)])
                sys.stdout.flush()
                self.maybe_write_fasta(self.file_handles["coding_nucleotide"], seqname, sequence)
            else:
                self.maybe_write_fasta(self.file_handles["noncoding_nucleotide"], seqname, sequence)
            scores.append(SingleReadScore(jaccard, n_kmers, text, frame))
        return scores

    def maybe_score_single_read(self, description, sequence):
        """Six per-frame scores of one read (translate.py:323-331) -- a batch of one through the GPU."""
        flat, offsets = concat_reads([sequence])
        results = score_reads(self.peptide_bloom_filter, flat, offsets, self._lut, self.jaccard_threshold)
        return self._expand(description, sequence, results[0], self._emitted_peptides(flat, offsets, results))

    # kept for interface parity; both are answered by the same GPU call
    def check_peptide_content(self, description, sequence):
        return self.maybe_score_single_read(description, sequence)

    def check_nucleotide_content(self, description, n_kmers, sequence):
        text = constants_translate.PROTEIN_CODING_CATEGORIES["too_short_nucleotide"]
        if n_kmers > 0:  # unreachable from the scoring path: it passes NaN (translate.py:328)
            text = constants_translate.PROTEIN_CODING_CATEGORIES["low_complexity_nucleotide"]
            self.maybe_write_fasta(self.file_handles["low_complexity_nucleotide"], description, sequence)
            return [SingleReadScore(NAN, n_kmers, text, frame) for frame in FRAMES]
        return [SingleReadScore(NAN, NAN, text, frame) for frame in FRAMES]

    def get_coding_score_line(self, description, jaccard, n_kmers, special_case, frame):
        if special_case is not None:
            return [description, jaccard, n_kmers, special_case, frame]
        if jaccard > self.jaccard_threshold:
            return [description, jaccard, n_kmers, "Coding", frame]
        return [description, jaccard, n_kmers, "Non-coding", frame]

    def _lines(self, reads_path, description, scores):
        lines = []
        for jaccard, n_kmers, special_case, frame in scores:
            line = self.get_coding_score_line(description, jaccard, n_kmers, special_case, frame)
            line.append(reads_path)
            lines.append(line)
        return lines

    def score_reads_per_record(self, reads_path, record):
        return self._lines(reads_path, record["name"],
                           self.maybe_score_single_read(record["name"], record["sequence"]))

    def score_reads_per_file(self, reads):
        """Score every read of one file: parse a batch -> one C call -> expand, preserving read and frame order."""
        scoring_lines = []
        for names, flat, offsets in fastx.read_batches(reads):
            results = score_reads(self.peptide_bloom_filter, flat, offsets, self._lut, self.jaccard_threshold)
            peptides = self._emitted_peptides(flat, offsets, results)
            data = flat.tobytes()
            for i, name in enumerate(names):
                sequence = data[int(offsets[i]):int(offsets[i + 1])].decode(*fastx.TEXT_CODEC)
                scoring_lines.extend(self._lines(reads, name, self._expand(name, sequence, results[i], peptides, i)))
        return scoring_lines

    # -- columnar path (what the CLI uses): no per-row python objects ---------------------------------------
    def _emit_batch(self, names, flat, offsets, results):
        """stdout / FASTA side effects of one batch, in read order then frame order (translate.py:244-298).  The records
        are formatted by the native library (``orph_format_fasta``) straight from the result array; the peptides come from
        one GPU call (``orph_translate_frames``)."""
        cat = results["category"]
        handles = self.file_handles
        with_lcp = handles["low_complexity_peptide"] is not None
        want = cat == _lib.CAT_CODING
        if with_lcp:
            want = want | (cat == _lib.CAT_LOW_COMPLEXITY_PEPTIDE)
        has_noncoding = handles["noncoding_nucleotide"] is not None and bool((cat == _lib.CAT_NON_CODING).any())
        reads_idx, frames_idx = np.nonzero(want)
        if len(reads_idx) == 0 and not has_noncoding:
            return
        if isinstance(names, fastx.NameList):
            name_blob, name_offsets = np.frombuffer(names.blob, dtype=np.uint8), names.offsets
        else:
            encoded = [n.encode(*fastx.TEXT_CODEC) for n in names]
            name_blob = np.frombuffer(b"".join(encoded), dtype=np.uint8)
            name_offsets = np.zeros(len(encoded) + 1, dtype=np.uint64)
            name_offsets[1:] = np.cumsum([len(e) for e in encoded])
        keys = np.asarray(FRAMES, dtype=np.int8)[frames_idx]
        peptides, pep_offsets = translate_frames_raw(self.peptide_bloom_filter.ctx, flat, offsets, reads_idx, keys)

        def emit(kind, handle):
            self._write_bytes(handle, format_fasta(kind, name_blob, name_offsets, results, flat, offsets, peptides, pep_offsets, with_lcp))

        if with_lcp:
            emit(3, handles["low_complexity_peptide"])
        if handles["coding_nucleotide"] is not None:
            emit(1, handles["coding_nucleotide"])
        if has_noncoding:
            emit(2, handles["noncoding_nucleotide"])
        emit(0, sys.stdout)
        sys.stdout.flush()

    def score_all_files_columnar(self, stream_csv=None):
        """Score every input file and return a columnar ``ScoreTable`` (same information as ``coding_scores``,
        6 rows per read, without building python rows); stdout / FASTA outputs are written on the way.  ``stream_csv``: the
        path the score CSV will be written to -- its rows are then formatted batch by batch while the next batches are
        scored (``table.write_csv(stream_csv)`` completes it)."""
        from .columnar import ScoreTable

        table = ScoreTable(self.alphabet, background_summary=bool(getattr(self, "json_summary", None)))
        if stream_csv:
            table.stream_csv_to(stream_csv)
        self.maybe_open_fastas()
        try:
            for reads in self.reads:
                if fastx._is_bz2(reads):  # bzip2 goes through the python reader
                    for names, flat, offsets in fastx.read_batches(reads, prefetch=1):
                        results = score_reads(self.peptide_bloom_filter, flat, offsets, self._lut, self.jaccard_threshold)
                        self._emit_batch(names, flat, offsets, results)
                        table.append(names, results, reads)
                    continue
                # file in, records out: parsed, packed, copied and scored by the native stream (orph_stream_*), the next
                # batch on its way while this one is expanded into the outputs
                lanes = self._replica_lanes()
                _trace("lanes decided")
                with ScoringStream(self.peptide_bloom_filter, reads, self._lut, self.jaccard_threshold, replicas=lanes) as stream:
                    _trace("stream open")
                    for batch in stream:
                        _trace("batch of %d reads scored" % len(batch.results))
                        if batch.status == _lib.ORPH_E_TOO_LONG:
                            bad = np.nonzero((batch.results["category"] == _lib.CAT_INVALID).all(axis=1))[0]
                            name = batch.names().tolist()[int(bad[0])] if len(bad) else "?"
                            raise ValueError(f"{reads}: read '{name}' is longer than {_lib.MAX_READ_LEN} nt and has a translation frame without a "
                                             "stop codon; such a frame (more than 65535 residues) cannot be scored")
                        _lib.check(batch.status)  # an empty read: ZeroDivisionError, as in the reference (translate.py:83)
                        results = batch.results.copy()
                        names = batch.names()
                        _trace("records and names copied")
                        self._emit_stream_batch_native(stream)
                        _trace("FASTA outputs written")
                        table.append(names, results, reads)
                        _trace("batch handed to the CSV / summary workers")
        except BaseException:
            table.discard_csv_stream()
            raise
        finally:
            self.maybe_close_fastas()
        return table

    def _replica_lanes(self):
        """Every GPU of the box takes part in scoring (the reference walks its files and reads serially,
        translate.py:367-379): the filter is replicated once over NVLink and batches are dealt round-robin to the devices
        by one native parser (orph_stream_open_multi); the outputs do not depend on the number of lanes.
        ORPHEUM_B200_DEVICES="0,1,..." names the lanes explicitly (repeats allowed); otherwise one lane per 2 GB of
        input, up to the visible GPUs -- a CUDA context per extra GPU costs a few tenths of a second, and one B200 already
        scores several times what the host's parser delivers, so small inputs stay on one GPU."""
        if getattr(self, "_lanes", None) is not None:
            return self._lanes
        from .gpu import device_count, replica_lanes

        graph = self.peptide_bloom_filter
        spec = os.environ.get("ORPHEUM_B200_DEVICES")
        if spec:
            devices = [int(d) for d in spec.split(",") if d.strip() != ""]
            self._lanes = replica_lanes(graph, len(devices), devices=devices[1:]) if len(devices) > 1 else []
            return self._lanes
        total = 0
        for path in self.reads:
            try:
                total += os.path.getsize(path)
            except OSError:
                pass
        n = max(1, min(device_count(), total // (2 << 30)))
        self._lanes = replica_lanes(graph, n) if n > 1 else []
        return self._lanes

    @staticmethod
    def _write_bytes(handle, blob):
        """FASTA bytes into a text handle: through its binary layer when that writes the same bytes (UTF-8, no newline
        translation), else decoded."""
        if not len(blob):
            return
        raw = getattr(handle, "buffer", None)
        if raw is not None and str(getattr(handle, "encoding", "")).lower().replace("-", "") == "utf8" \
                and getattr(handle, "errors", None) in ("strict", "surrogateescape", None) and os.linesep == "\n":
            handle.flush()
            raw.write(blob)
        else:
            handle.write(bytes(blob).decode(*fastx.TEXT_CODEC))

    def _emit_stream_batch_native(self, stream):
        """Side effects of the stream batch being held (translate.py:244-298), every byte of them produced by the library
        (orph_stream_emit): same files, same order of writes as _emit_batch."""
        handles = self.file_handles
        kinds = 1 | (8 if handles["low_complexity_peptide"] is not None else 0) | (2 if handles["coding_nucleotide"] is not None else 0) \
            | (4 if handles["noncoding_nucleotide"] is not None else 0)
        blobs = stream.emit(self.peptide_bloom_filter.ctx, kinds, copy=False)  # written out before the stream moves on
        if handles["low_complexity_peptide"] is not None:
            self._write_bytes(handles["low_complexity_peptide"], blobs[3])
        if handles["coding_nucleotide"] is not None:
            self._write_bytes(handles["coding_nucleotide"], blobs[1])
        if handles["noncoding_nucleotide"] is not None:
            self._write_bytes(handles["noncoding_nucleotide"], blobs[2])
        self._write_bytes(sys.stdout, blobs[0])
        sys.stdout.flush()

    def set_coding_scores_all_files(self):
        self.maybe_open_fastas()
        try:
            self.coding_scores = [line for reads in self.reads for line in self.score_reads_per_file(reads)]
        finally:
            self.maybe_close_fastas()

    def get_coding_scores_all_files(self):
        return self.coding_scores


@click.command()
@click.argument("peptides", nargs=1)
@click.argument("reads", nargs=-1)
@click.option("--peptide-ksize", default=None, type=int,
              help="K-mer size of the peptide sequence to use. Defaults for different alphabets are, "
                   "protein: {}, dayhoff {}, hydrophobic-polar {}".format(
                       constants_index.DEFAULT_PROTEIN_KSIZE, constants_index.DEFAULT_DAYHOFF_KSIZE,
                       constants_index.DEFAULT_HP_KSIZE))
@click.option("--save-peptide-bloom-filter", is_flag=True, default=False,
              help="If specified, save the peptide bloom filter. Default filename is the name of the fasta file plus a "
                   "suffix denoting the protein encoding and peptide ksize")
@click.option("--peptides-are-bloom-filter", is_flag=True, default=False, help="Peptide file is already a bloom filter")
@click.option("--jaccard-threshold", default=None, type=click.FLOAT, callback=validate_jaccard,
              help="Minimum fraction of peptide k-mers from read in the peptide database for this read to be called a "
                   "'coding read'. Default: {} for protein and dayhoff encodings, and {} for hydrophobic-polar (hp) "
                   "encoding".format(constants_translate.DEFAULT_JACCARD_THRESHOLD,
                                     constants_translate.DEFAULT_HP_JACCARD_THRESHOLD))
@click.option("--alphabet", "--encoding", "--molecule", default="protein",
              help="The type of amino acid encoding to use. Default is 'protein', but 'dayhoff' or 'hydrophobic-polar' "
                   "can be used")
@click.option("--csv", default=None, help="Name of csv file to write with all sequence reads and their coding scores")
@click.option("--parquet", default=None, help="Name of parquet file to write with all sequence reads and their coding scores")
@click.option("--json-summary", default=None,
              help="Name of json file to write summarization of coding/noncoding/other categorizations, the "
                   "min/max/mean/median/stddev of Jaccard scores, and other")
@click.option("--coding-nucleotide-fasta", help="If specified, save the coding nucleotides to this file")
@click.option("--noncoding-nucleotide-fasta", help="If specified, save the noncoding nucleotides to this file")
@click.option("--low-complexity-nucleotide-fasta", help="If specified, save the low-complexity nucleotides to this file")
@click.option("--low-complexity-peptide-fasta", help="If specified, save the low-complexity peptides to this file")
@click.option("--tablesize", type=constants_index.BASED_INT, default="1e8", help="Size of the bloom filter table to use")
@click.option("--n-tables", type=int, default=constants_index.DEFAULT_N_TABLES, help="Size of the bloom filter table to use")
@click.option("--long-reads", is_flag=True,
              help="If set, then only considers reading frames starting with start codon (ATG) and ending in a stop "
                   "codon (TAG, TAA, TGA)")
@click.option("--verbose", is_flag=True, help="Print more output")
def cli(peptides, reads, peptide_ksize=None, save_peptide_bloom_filter=True, peptides_are_bloom_filter=False,
        jaccard_threshold=None, alphabet="protein", csv=None, parquet=None, json_summary=None,
        coding_nucleotide_fasta=None, noncoding_nucleotide_fasta=None, low_complexity_nucleotide_fasta=None,
        low_complexity_peptide_fasta=None, tablesize=constants_index.DEFAULT_MAX_TABLESIZE,
        n_tables=constants_index.DEFAULT_N_TABLES, long_reads=False, verbose=False):
    """Writes coding peptides from reads to standard output"""
    _T0[0] = 0.0
    _trace("start")
    translate_obj = Translate(locals())
    _trace("filter loaded")
    table = translate_obj.score_all_files_columnar(stream_csv=csv)
    _trace("all batches scored")
    # the three outputs are independent: the CSV is formatted by the native writer (GIL released) while the parquet
    # table and the JSON summary are built here
    from concurrent.futures import ThreadPoolExecutor

    with ThreadPoolExecutor(max_workers=2) as pool:
        pending = []
        if csv:
            pending.append(pool.submit(table.write_csv, csv))
        if parquet and len(table):
            # (no read scored -- an empty input -- writes no parquet file; the reference builds a RecordBatch from six empty
            # python lists there, whose column types pyarrow cannot infer: nothing useful to be faithful to)
            pending.append(pool.submit(table.write_parquet, parquet))
        if json_summary:
            table.write_json_summary(json_summary, reads, translate_obj.peptide_bloom_filter_filename,
                                     translate_obj.peptide_ksize, translate_obj.jaccard_threshold)
            _trace("JSON summary written")
        for job in pending:
            job.result()
        _trace("CSV / parquet complete")


if __name__ == "__main__":
    cli()
