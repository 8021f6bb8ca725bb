"""Columnar outputs straight from the GPU result records.

The reference materialises six Python lists per read (``orpheum/translate.py:342-365``) and hands them to
``CreateSaveSummary`` (``orpheum/create_save_summary.py:42-240``); that cannot scale to the read counts the
GPU path scores.  ``ScoreTable`` keeps the per-read records as numpy arrays and writes the same files from
them -- the CSV byte for byte as ``csv.writer`` would (``nan``, ``repr(float)``, minimal quoting, ``\\n``), the
parquet with the same schema, the JSON summary with the same keys, values and key order.
"""
import json
import os

import numpy as np

from . import _lib
from .constants_translate import FRAMES, LOW_COMPLEXITY_CATEGORIES, PROTEIN_CODING_CATEGORIES, SCORING_DF_COLUMNS, category_text

_FRAMES = np.asarray(FRAMES, dtype=np.int64)


class ScoreTable:
    """Accumulates (names, orph_read_result records, file name) batches in read order."""

    def __init__(self, alphabet, background_summary=False):
        """background_summary: every appended batch's share of the JSON summary (read classes, coding-frame histogram,
        jaccard column, id hashes) is computed right away on a background thread -- numpy and the native calls release the
        GIL -- so that only the order statistics over the whole jaccard column are left when the last batch is in."""
        self.alphabet = alphabet
        self._names, self._records, self._files = [], [], []
        self._partials = [] if background_summary else None
        self._csv_path = self._csv_tmp = None
        self._csv_jobs = []

    def append(self, names, records, filename):
        """names: a list of str or a fastx.NameList (byte blob + offsets, kept as is for the CSV writer)."""
        if len(names):
            self._names.append(names)
            self._records.append(records)
            self._files.append((filename, len(names)))
            if self._partials is not None:
                self._partials.append(_background().submit(_partial_summary, names, records))
            if self._csv_tmp is not None:
                self._csv_jobs.append(_background("csv").submit(self.append_csv_batch, self._csv_tmp, names, records, filename))

    def stream_csv_to(self, path):
        """From now on every appended batch's rows are formatted and appended to the CSV right away, on a background thread
        (the native formatter releases the GIL); ``write_csv(path)`` then only waits for the last batch.  A run that fails
        leaves no CSV behind (``discard_csv_stream``), as with the reference, which writes it last."""
        self._csv_path = self._csv_tmp = path
        self.write_csv_header(self._csv_tmp)

    def discard_csv_stream(self):
        for job in self._csv_jobs:
            try:
                job.result()
            except Exception:
                pass
        if self._csv_tmp is not None and os.path.exists(self._csv_tmp):
            os.remove(self._csv_tmp)
        self._csv_path = self._csv_tmp = None
        self._csv_jobs = []

    def __len__(self):
        return sum(len(n) for n in self._names)

    # -- flat per-row columns (6 rows per read, frame order 1,2,3,-1,-2,-3) ------------------------------
    def _columns(self):
        if not self._names:
            return None
        names = np.concatenate([np.asarray(n.tolist() if hasattr(n, "tolist") else n, dtype=object) for n in self._names])
        rec = np.concatenate(self._records)
        cat = rec["category"].reshape(-1).astype(np.int64)
        n_kmers = rec["n_kmers"].reshape(-1).astype(np.int64)
        n_hits = rec["n_hits"].reshape(-1).astype(np.int64)
        scored = (cat == _lib.CAT_CODING) | (cat == _lib.CAT_NON_CODING)
        has_n = scored | (cat == _lib.CAT_LOW_COMPLEXITY_PEPTIDE)
        jaccard = np.full(cat.shape, np.nan)
        np.divide(n_hits, n_kmers, out=jaccard, where=scored)  # the one IEEE division of translate.py:204
        files = np.concatenate([np.full(6 * n, i, dtype=np.int64) for i, (_, n) in enumerate(self._files)])
        return {
            "names": names, "read_index": np.repeat(np.arange(len(names)), 6), "category": cat, "n_kmers": n_kmers,
            "n_hits": n_hits, "scored": scored, "has_n": has_n, "jaccard": jaccard,
            "frame": np.tile(_FRAMES, len(names)), "file_index": files, "file_names": [f for f, _ in self._files],
        }

    def _category_texts(self, present_codes):
        # low-complexity text only if such a frame exists (the reference raises KeyError lazily, translate.py:254-256)
        return {int(c): category_text(int(c), self.alphabet) for c in present_codes}

    # -- arrow / parquet ---------------------------------------------------------------------------------
    def to_arrow(self):
        import pyarrow as pa

        c = self._columns()
        if c is None:
            return None
        texts = self._category_texts(np.unique(c["category"]))
        codes = sorted(texts)
        remap = np.zeros(256, dtype=np.int32)
        for i, code in enumerate(codes):
            remap[code] = i
        category = pa.DictionaryArray.from_arrays(pa.array(remap[c["category"]], type=pa.int32()),
                                                  pa.array([texts[k] for k in codes])).cast(pa.string())
        try:
            name_values = pa.array(c["names"].tolist(), type=pa.string())
        except (UnicodeEncodeError, pa.ArrowInvalid):  # ids that are not valid UTF-8 (kept byte-exact in the CSV): U+FFFD here
            name_values = pa.array([n.encode("utf-8", "surrogateescape").decode("utf-8", "replace") for n in c["names"].tolist()],
                                   type=pa.string())
        read_id = pa.DictionaryArray.from_arrays(pa.array(c["read_index"], type=pa.int64()), name_values).cast(pa.string())
        filename = pa.DictionaryArray.from_arrays(pa.array(c["file_index"], type=pa.int64()),
                                                  pa.array(c["file_names"], type=pa.string())).cast(pa.string())
        if c["has_n"].all():
            n_kmers = pa.array(c["n_kmers"], type=pa.int64())  # python ints only -> int64, as pyarrow infers for the reference
        else:
            n_kmers = pa.array(np.where(c["has_n"], c["n_kmers"].astype(np.float64), np.nan), type=pa.float64())
        return pa.table([read_id, pa.array(c["jaccard"], type=pa.float64()), n_kmers, category,
                         pa.array(c["frame"], type=pa.int64()), filename], names=SCORING_DF_COLUMNS)

    def write_parquet(self, path):
        import pyarrow.parquet as pq

        table = self.to_arrow()
        if table is not None:
            pq.write_table(table, path)

    # -- CSV -----------------------------------------------------------------------------------------------
    @staticmethod
    def _csv_field(text):
        """One CSV field exactly as csv.writer renders it (minimal quoting)."""
        if any(ch in text for ch in ',"\r\n'):
            return '"' + text.replace('"', '""') + '"'
        return text

    def write_csv(self, path):
        """Byte-identical to ``csv.writer(lineterminator="\n")`` over the reference's row lists; the rows are
        formatted by the native writer (orph_csv_write_scores) batch by batch."""
        if self._csv_tmp is not None and path == self._csv_path:  # streamed while the batches came in
            for job in self._csv_jobs:
                job.result()
            self._csv_path = self._csv_tmp = None
            self._csv_jobs = []
            return
        self.write_csv_header(path)
        for names, records, (filename, _) in zip(self._names, self._records, self._files):
            self.append_csv_batch(path, names, records, filename)

    @staticmethod
    def write_csv_header(path):
        with open(path, "wb") as out:
            out.write((",".join(SCORING_DF_COLUMNS) + "\n").encode())

    def append_csv_batch(self, path, names, records, filename):
        """Append the 6 rows per read of one batch to the CSV (after write_csv_header).  The native formatter runs
        with the GIL released, so the CLI streams batches to it from a worker thread while later batches are scored."""
        import ctypes

        lib = _lib.load()
        present = np.unique(records["category"])
        texts = [None] * 6
        for code in present:
            texts[int(code)] = self._csv_field(category_text(int(code), self.alphabet)).encode()
        raw = getattr(names, "blob", None)
        if raw is not None and not any(ch in raw for ch in b',"\r\n'):
            # the usual case: no read id needs quoting, the parser's name blob IS the column of CSV fields
            blob = np.frombuffer(raw, dtype=np.uint8) if len(raw) else np.zeros(1, dtype=np.uint8)
            offsets = np.ascontiguousarray(names.offsets, dtype=np.uint64)
            n_fields = len(names)
        else:
            fields = [self._csv_field(n).encode("utf-8", "surrogateescape") for n in (names.tolist() if hasattr(names, "tolist") else names)]
            offsets = np.zeros(len(fields) + 1, dtype=np.uint64)
            offsets[1:] = np.cumsum([len(f) for f in fields])
            blob = np.frombuffer(b"".join(fields), dtype=np.uint8) if offsets[-1] else np.zeros(1, dtype=np.uint8)
            n_fields = len(fields)
        records = np.ascontiguousarray(records)
        arr = (ctypes.c_char_p * 6)(*texts)
        _lib.check(lib.orph_csv_write_scores(os.fsencode(path), 1, _lib.ptr(blob), _lib.ptr(offsets), n_fields,
                                             _lib.ptr(records), arr, self._csv_field(filename).encode(), 0))

    # -- JSON summary --------------------------------------------------------------------------------------
    def empty_categories(self):
        return empty_categories(self.alphabet)

    def _ids_all_distinct(self):
        """True when every read id of the table is provably different from every other one: 64-bit hashes of the ids (taken
        from the parser's name blobs, natively, no python strings) are pairwise distinct.  False = unknown (a repeated id,
        a rare hash collision, or names held as python lists): the caller compares the ids themselves."""
        import ctypes

        lib = _lib.load()
        hashes = []
        for names in self._names:
            h = _name_hashes(names)
            if h is None:
                return False
            hashes.append(h)
        allh = np.concatenate(hashes) if len(hashes) > 1 else hashes[0]
        flag = ctypes.c_int(0)
        _lib.check(lib.orph_u64_all_distinct(_lib.ptr(allh), len(allh), 0, ctypes.byref(flag)))
        return bool(flag.value)

    def _summary_from_records(self, filenames, peptide_bloom_filter_filename, peptide_ksize, jaccard_threshold):
        """The summary of a table whose read ids are all different (the usual case), straight from the record arrays: no
        per-row python objects, no 6-rows-per-read expansion of anything but the jaccard column."""
        rec = np.concatenate(self._records) if len(self._records) > 1 else self._records[0]
        present, coding_counts, jaccard, scored = _record_columns(rec)
        return _finish_summary(self.alphabet, _class_counts(present), _coding_histogram(coding_counts), (jaccard, scored), filenames,
                               peptide_bloom_filter_filename, peptide_ksize, jaccard_threshold)

    def _summary_from_partials(self, filenames, peptide_bloom_filter_filename, peptide_ksize, jaccard_threshold):
        """Combines the per-batch shares computed in the background; None if the read ids are not provably distinct."""
        import ctypes

        parts = [job.result() for job in self._partials]
        _trace("summary: per-batch shares ready")
        if any(p["hashes"] is None for p in parts):
            return None
        hashes = np.concatenate([p["hashes"] for p in parts]) if len(parts) > 1 else parts[0]["hashes"]
        flag = ctypes.c_int(0)
        _lib.check(_lib.load().orph_u64_all_distinct(_lib.ptr(hashes), len(hashes), 0, ctypes.byref(flag)))
        if not flag.value:
            return None
        classes = [sum(p["classes"][i] for p in parts) for i in range(3)] + [sum(p["classes"][3] for p in parts), sum(p["classes"][4] for p in parts)]
        histogram = {}
        for p in parts:  # keys in order of first appearance over the whole table
            for value, count in p["histogram"].items():
                histogram[value] = histogram.get(value, 0) + count
        _trace("summary: ids distinct, classes and histogram merged")
        jaccard = np.concatenate([p["jaccard"] for p in parts]) if len(parts) > 1 else parts[0]["jaccard"]
        scored = np.concatenate([p["scored"] for p in parts]) if len(parts) > 1 else parts[0]["scored"]
        _trace("summary: jaccard column concatenated")
        return _finish_summary(self.alphabet, classes, histogram, (jaccard, scored), filenames, peptide_bloom_filter_filename, peptide_ksize,
                               jaccard_threshold)

    def summary(self, filenames, peptide_bloom_filter_filename, peptide_ksize, jaccard_threshold):
        """The dict ``CreateSaveSummary.maybe_write_json_summary`` dumps (create_save_summary.py:91-240)."""
        if self._names and self._partials is not None and len(self._partials) == len(self._names):
            done = self._summary_from_partials(filenames, peptide_bloom_filter_filename, peptide_ksize, jaccard_threshold)
            if done is not None:
                return done
        if self._names and self._ids_all_distinct():
            return self._summary_from_records(filenames, peptide_bloom_filter_filename, peptide_ksize, jaccard_threshold)
        c = self._columns()
        if c is None:
            return summarize(self.alphabet, None, filenames, peptide_bloom_filter_filename, peptide_ksize, jaccard_threshold)
        names = c["names"]
        coding = c["category"] == _lib.CAT_CODING
        if len(set(names.tolist())) == len(names):
            # the usual case, every read id occurs once: groups are the reads themselves and the coding frames per id are
            # the coding frames per read (no per-id python bookkeeping)
            per_read = coding.reshape(-1, 6).sum(axis=1)
            rows = (np.repeat(np.arange(len(names)), 6), c["category"], per_read[per_read > 0], c["jaccard"])
        else:
            group_of_read = consecutive_groups(names)
            coding_rows = np.nonzero(coding)[0]
            rows = (np.repeat(group_of_read, 6), c["category"], names[c["read_index"][coding_rows]].tolist(), c["jaccard"])
        return summarize(self.alphabet, rows, filenames, peptide_bloom_filter_filename, peptide_ksize, jaccard_threshold)

    def write_json_summary(self, path, filenames, peptide_bloom_filter_filename, peptide_ksize, jaccard_threshold):
        summary = self.summary(filenames, peptide_bloom_filter_filename, peptide_ksize, jaccard_threshold)
        _trace("summary: statistics done")
        with open(path, "w") as handle:
            json.dump(summary, fp=handle)
        return summary


_bg_executors = {}


def _trace(what):
    from .translate import _trace as t  # (translate imports this module lazily)

    t(what)


def _background(which="summary"):
    """One background thread per kind of work ("summary", "csv"), shared by every table of the process, created on first
    use: jobs of one kind run in submission order."""
    if which not in _bg_executors:
        from concurrent.futures import ThreadPoolExecutor

        _bg_executors[which] = ThreadPoolExecutor(1, thread_name_prefix="orph-" + which)
    return _bg_executors[which]


def _name_hashes(names):
    """64-bit hashes of a batch's read ids, taken natively from the parser's name blob; None for python lists."""
    blob, offsets = getattr(names, "blob", None), getattr(names, "offsets", None)
    if blob is None or offsets is None:
        return None
    raw = np.frombuffer(blob, dtype=np.uint8) if len(blob) else np.zeros(1, dtype=np.uint8)
    offs = np.ascontiguousarray(offsets, dtype=np.uint64)
    out = np.empty(len(names), dtype=np.uint64)
    _lib.check(_lib.load().orph_hash_names(_lib.ptr(raw), _lib.ptr(offs), len(names), _lib.ptr(out), 0))
    return out


def _record_columns(rec):
    """(presence matrix [n, 6], coding frames per read that has any, jaccard per row, scored mask per row) of a record array
    whose read ids are all different."""
    cat = rec["category"]                                     # [n, 6] uint8
    bits = np.bitwise_or.reduce(np.left_shift(np.uint8(1), cat), axis=1)  # categories present per read, one bit each
    present = (bits[:, None] >> np.arange(6, dtype=np.uint8)) & 1 != 0
    per_read = (cat == _lib.CAT_CODING).sum(axis=1)
    scored = ((cat == _lib.CAT_CODING) | (cat == _lib.CAT_NON_CODING)).reshape(-1)
    jaccard = np.full(scored.shape, np.nan)
    np.divide(rec["n_hits"].reshape(-1), rec["n_kmers"].reshape(-1), out=jaccard, where=scored)  # translate.py:204
    return present, per_read[per_read > 0], jaccard, scored


def _class_counts(present):
    """One category per read: Coding > too-short peptide > too-short read > the only category > Non-coding.
    -> (coding, too-short peptide, too-short read, reads whose only category is code c [6], non-coding)."""
    coding = present[:, _lib.CAT_CODING]
    short_pep = ~coding & present[:, _lib.CAT_TOO_SHORT_PEPTIDE]
    short_read = ~coding & ~short_pep & present[:, _lib.CAT_LOW_COMPLEXITY_NUCLEOTIDE]
    undecided = ~coding & ~short_pep & ~short_read
    single = undecided & (present.sum(axis=1) == 1)
    only = np.bincount(np.argmax(present[single], axis=1), minlength=6).astype(np.int64)
    return [int(coding.sum()), int(short_pep.sum()), int(short_read.sum()), only, int((undecided & ~single).sum())]


def _coding_histogram(coding_counts):
    """{coding frames per id: number of ids}, keys in order of first appearance (as Counter / dict insertion gives them)."""
    coding_counts = np.asarray(coding_counts)
    if coding_counts.size == 0:
        return {}
    if coding_counts.max() < 64:  # the usual case (at most six frames per read): no sort
        n_ids = np.bincount(coding_counts)
        values = np.nonzero(n_ids)[0]
        first = [int(np.argmax(coding_counts == v)) for v in values]
    else:
        values, first, n_ids_ = np.unique(coding_counts, return_index=True, return_counts=True)
        n_ids = dict(zip(values.tolist(), n_ids_.tolist()))
    return {int(values[at]): int(n_ids[values[at]]) for at in np.argsort(first)}


def _partial_summary(names, records):
    """One batch's share of the summary, from ONE native call over its records (orph_summary_partial: no interpreter
    time while the main thread expands the next batch); same values as _record_columns / _class_counts / _coding_histogram."""
    import ctypes

    records = np.ascontiguousarray(records)
    n = len(records)
    jaccard = np.empty(6 * n, dtype=np.float64)
    scored = np.empty(6 * n, dtype=np.uint8)
    classes = (ctypes.c_uint64 * 10)()
    hist = (ctypes.c_uint64 * 7)()
    first = (ctypes.c_uint64 * 7)()
    _lib.check(_lib.load().orph_summary_partial(_lib.ptr(records), n, _lib.ptr(jaccard), _lib.ptr(scored), classes, hist, first, 0))
    order = sorted((first[c], c) for c in range(1, 7) if hist[c])
    return {"classes": [int(classes[0]), int(classes[1]), int(classes[2]), np.array(classes[3:9], dtype=np.int64), int(classes[9])],
            "histogram": {c: int(hist[c]) for _, c in order}, "jaccard": jaccard, "scored": scored.view(np.bool_),
            "hashes": _name_hashes(names)}


_native_sums_ok = None


def _native_sums_match_numpy():
    """orph_f64_pairwise_sum restates numpy's float64 summation order; should a numpy build ever sum differently, this
    one-time comparison on a column shaped like the real one notices and the statistics stay with numpy."""
    global _native_sums_ok
    if _native_sums_ok is None:
        rng = np.random.default_rng(12345)
        n = 300_007
        x = rng.integers(0, 60, n) / rng.integers(1, 60, n)
        ok = rng.random(n) < 0.4
        try:
            arr = np.where(ok, x, 0.0)
            cnt = int(ok.sum())
            mean = np.sum(arr) / cnt
            dev = arr - mean
            dev[~ok] = 0.0
            dev *= dev
            _native_sums_ok = (_pairwise_sum(x, ok) == float(np.sum(arr)) and _pairwise_sum(x, ok, float(mean), True) == float(np.sum(dev))
                               and _pairwise_sum(x[:1000], None) == float(np.sum(x[:1000])))
        except Exception:
            _native_sums_ok = False
    return _native_sums_ok


def _pairwise_sum(x, ok, shift=0.0, square=False):
    import ctypes

    out = ctypes.c_double()
    x = np.ascontiguousarray(x, dtype=np.float64)
    mask = None if ok is None else np.ascontiguousarray(ok).view(np.uint8)
    _lib.check(_lib.load().orph_f64_pairwise_sum(_lib.ptr(x), _lib.ptr(mask) if mask is not None else None, len(x), float(shift), int(bool(square)), 0,
                                                 ctypes.byref(out)))
    return out.value


def _finish_summary(alphabet, classes, histogram, jaccard, filenames, peptide_bloom_filter_filename, peptide_ksize, jaccard_threshold):
    input_files = [os.path.basename(f) for f in ([filenames] if isinstance(filenames, str) else filenames)]
    counts = empty_categories(alphabet)
    counts["Coding"] += classes[0]
    counts[PROTEIN_CODING_CATEGORIES["too_short_peptide"]] += classes[1]
    counts[PROTEIN_CODING_CATEGORIES["too_short_nucleotide"]] += classes[2]
    for code in np.nonzero(classes[3])[0]:
        counts[category_text(int(code), alphabet)] += int(classes[3][code])
    counts["Non-coding"] += classes[4]
    total = sum(counts.values())
    n_coding_ids = sum(histogram.values())
    label = "Number of reads with {} putative protein-coding translations".format
    return {
        "input_files": input_files,
        "jaccard_info": jaccard_info(jaccard),
        "categorization_counts": counts,
        "categorization_percentages": {k: 100 * v / total for k, v in counts.items()},
        "histogram_n_coding_frames_per_read": {label(k): v for k, v in histogram.items()},
        "histogram_n_coding_frames_per_read_percentages": {label(k): 100 * v / n_coding_ids for k, v in histogram.items()},
        "peptide_bloom_filter": peptide_bloom_filter_filename, "peptide_alphabet": alphabet, "peptide_ksize": peptide_ksize,
        "jaccard_threshold": jaccard_threshold,
    }


def empty_categories(alphabet):
    counts = dict.fromkeys(PROTEIN_CODING_CATEGORIES.values(), 0)
    counts[LOW_COMPLEXITY_CATEGORIES[alphabet]] = 0
    return counts


def consecutive_groups(names):
    """Group index per element, a new group wherever the value differs from its predecessor: what
    ``itertools.groupby`` does to the read ids in the reference (create_save_summary.py:180-182)."""
    names = np.asarray(names, dtype=object)
    start = np.ones(len(names), dtype=bool)
    if len(names) > 1:
        start[1:] = names[1:] != names[:-1]
    return np.cumsum(start) - 1


def jaccard_info(jaccard):
    """count / mean / std / min / quartiles / max of the non-NaN jaccards.  Mean and std keep numpy's nan-functions (the
    summation order over the full column is part of the value, to the last bit); the order statistics are taken once
    from the compacted column instead of five separate NaN-stripping passes -- same elements, same linear
    interpolation, same values.  A (column, validity mask) pair takes the same arithmetic without numpy's temporaries."""
    if isinstance(jaccard, tuple):
        jaccard, ok = jaccard
        valid = jaccard[ok]
        if valid.size:
            # np.nanmean / np.nanstd, step by step (numpy/lib/_nanfunctions_impl.py): NaN -> 0, pairwise sum over the FULL
            # column, divide by the count; deviations with the NaN positions reset to 0, squared, summed, divided
            cnt = int(valid.size)
            if jaccard.size > 1_000_000 and _native_sums_match_numpy():
                # the same two sums in numpy's summation order, natively and on every core (orph_f64_pairwise_sum), while
                # another thread takes the order statistics (numpy's partition drops the GIL)
                order_stats = _background("quantiles").submit(lambda: (np.percentile(valid, [25, 50, 75]), valid.min(), valid.max()))
                mean = np.float64(_pairwise_sum(jaccard, ok)) / cnt
                std = np.sqrt(np.float64(_pairwise_sum(jaccard, ok, float(mean), True)) / cnt)
                (q25, q50, q75), lowest, highest = order_stats.result()
                return {"count": cnt, "mean": mean, "std": std, "min": lowest, "25%": q25, "50%": q50, "75%": q75, "max": highest}
            else:
                arr = np.where(ok, jaccard, 0.0)
                mean = np.sum(arr) / cnt
                np.subtract(arr, mean, out=arr)
                arr[~ok] = 0.0
                np.multiply(arr, arr, out=arr)
                std = np.sqrt(np.sum(arr) / cnt)
            q25, q50, q75 = np.percentile(valid, [25, 50, 75])
            return {"count": int(valid.size), "mean": mean, "std": std, "min": valid.min(), "25%": q25, "50%": q50, "75%": q75,
                    "max": valid.max()}
    valid = jaccard[~np.isnan(jaccard)]
    if valid.size == 0:
        nan = float("nan")
        return {"count": 0, "mean": np.nanmean(jaccard) if jaccard.size else nan, "std": np.nanstd(jaccard) if jaccard.size else nan,
                "min": nan, "25%": nan, "50%": nan, "75%": nan, "max": nan}
    q25, q50, q75 = np.percentile(valid, [25, 50, 75])
    return {"count": int(valid.size), "mean": np.nanmean(jaccard), "std": np.nanstd(jaccard), "min": valid.min(), "25%": q25, "50%": q50,
            "75%": q75, "max": valid.max()}


def summarize(alphabet, rows, filenames, peptide_bloom_filter_filename, peptide_ksize, jaccard_threshold):
    """The JSON-summary dict.  rows = (group index per row, category code per row, read ids of the Coding rows in
    row order -- or, when every id is unique, an int array with the number of Coding rows of each id that has any, in
    read order --, jaccard per row), or None when no read was scored."""
    input_files = [os.path.basename(f) for f in ([filenames] if isinstance(filenames, str) else filenames)]
    if rows is None:
        empty = empty_categories(alphabet)
        zero_hist = {str(i): 0 for i in range(len(empty))}
        return {
            "input_files": input_files,
            "jaccard_info": {"count": 0, "mean": None, "std": None, "min": None, "25%": None, "50%": None, "75%": None,
                             "max": None},
            "categorization_counts": empty, "categorization_percentages": empty,
            "histogram_n_coding_frames_per_read": zero_hist,
            "histogram_n_coding_frames_per_read_percentages": dict(zero_hist),
        }
    group, codes, coding_ids, jaccard = rows
    # one category per group of rows: Coding > too-short peptide > too-short read > the only category > Non-coding
    if codes is None:
        present = group  # the fast form: already the [n_groups, 6] presence matrix
    else:
        present = np.zeros((int(group[-1]) + 1, 6), dtype=bool)
        present[group, codes] = True
    classes = _class_counts(present)
    # coding frames per read id; ids are merged by value here, as collections.Counter does (:216-222)
    if isinstance(coding_ids, np.ndarray):  # already the number of coding frames of each id that has any (unique ids)
        histogram = _coding_histogram(coding_ids)
    else:
        histogram, per_id = {}, {}
        for read_id in coding_ids:
            per_id[read_id] = per_id.get(read_id, 0) + 1
        for n in per_id.values():
            histogram[n] = histogram.get(n, 0) + 1
    return _finish_summary(alphabet, classes, histogram, jaccard, filenames, peptide_bloom_filter_filename, peptide_ksize, jaccard_threshold)
