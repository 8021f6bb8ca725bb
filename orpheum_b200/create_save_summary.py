"""CSV / parquet / JSON-summary writers for the scoring rows (mirror of
``orpheum/create_save_summary.py``: ``maybe_write_csv`` ``:52-64``, ``maybe_write_parquet`` ``:66-83``,
``maybe_write_json_summary`` ``:91-134``, the per-read category roll-up ``:172-212`` and the
coding-frame histogram ``:214-240``).  File formats are part of the parity contract.
"""
import csv as csv_module
import json
import os
from collections import Counter
from itertools import groupby

import numpy as np

from .constants_translate import LOW_COMPLEXITY_CATEGORIES, PROTEIN_CODING_CATEGORIES, SCORING_DF_COLUMNS


class CreateSaveSummary:
    def __init__(self, filenames, csv, parquet, json_summary, peptide_bloom_filter_filename, alphabet, peptide_ksize,
                 jaccard_threshold, coding_scores):
        if type(filenames) is str:
            filenames = [filenames]
        self.unique_filenames = [os.path.basename(f) for f in filenames]
        self.csv = csv
        self.parquet = parquet
        self.json_summary = json_summary
        self.peptide_bloom_filter_filename = peptide_bloom_filter_filename
        self.alphabet = alphabet
        self.peptide_ksize = peptide_ksize
        self.jaccard_threshold = jaccard_threshold
        self.coding_scores = coding_scores
        if self.coding_scores != []:
            columns = list(zip(*self.coding_scores))
            (self.read_ids, self.jaccard_in_peptide_dbs, self.n_kmers, self.categories, self.translation_frames,
             self.filenames) = (list(c) for c in columns)

    # -- writers -------------------------------------------------------------------------------------
    def maybe_write_csv(self):
        if not self.csv:
            return
        with open(self.csv, "w") as handle:
            writer = csv_module.writer(handle, lineterminator="\n")
            writer.writerow(SCORING_DF_COLUMNS)
            writer.writerows(self.coding_scores)

    def maybe_write_parquet(self):
        if not self.parquet:
            return
        import pyarrow as pa
        import pyarrow.parquet as pq

        batch = pa.RecordBatch.from_arrays(
            [self.read_ids, self.jaccard_in_peptide_dbs, self.n_kmers, self.categories, self.translation_frames,
             self.filenames],
            names=SCORING_DF_COLUMNS,
        )
        pq.write_table(pa.Table.from_batches([batch]), self.parquet)

    def make_empty_coding_categories(self):
        counts = dict.fromkeys(PROTEIN_CODING_CATEGORIES.values(), 0)
        counts[LOW_COMPLEXITY_CATEGORIES[self.alphabet]] = 0
        return counts

    def maybe_write_json_summary(self):
        if not self.json_summary:
            return None
        if self.coding_scores == []:
            empty = self.make_empty_coding_categories()
            zero_hist = {str(i): 0 for i in range(len(empty))}
            summary = {
                "input_files": self.unique_filenames,
                "jaccard_info": None,
                "categorization_counts": empty,
                "categorization_percentages": empty,
                "histogram_n_coding_frames_per_read": zero_hist,
                "histogram_n_coding_frames_per_read_percentages": dict(zero_hist),
            }
            # the reference lists "count" first
            summary["jaccard_info"] = {"count": 0, "mean": None, "std": None, "min": None, "25%": None, "50%": None,
                                       "75%": None, "max": None}
        else:
            summary = self.generate_coding_summary()
        with open(self.json_summary, "w") as handle:
            json.dump(summary, fp=handle)
        if self.coding_scores != []:
            for attr in ("read_ids", "jaccard_in_peptide_dbs", "n_kmers", "categories", "translation_frames", "filenames",
                         "coding_scores"):
                delattr(self, attr)
        return summary

    # -- summary -------------------------------------------------------------------------------------
    def generate_coding_summary(self):
        frame_percentages, frame_counts = self.get_n_translated_frames_per_read()
        category_percentages, category_counts = self.get_n_per_coding_category()
        j = self.jaccard_in_peptide_dbs
        jaccard_info = {
            "count": int(np.count_nonzero(~np.isnan(j))),
            "mean": np.nanmean(j),
            "std": np.nanstd(j),
            "min": np.nanmin(j),
            "25%": np.nanpercentile(j, 25),
            "50%": np.nanpercentile(j, 50),
            "75%": np.nanpercentile(j, 75),
            "max": np.nanmax(j),
        }
        return {
            "input_files": self.unique_filenames,
            "jaccard_info": jaccard_info,
            "categorization_counts": category_counts,
            "categorization_percentages": category_percentages,
            "histogram_n_coding_frames_per_read": frame_counts,
            "histogram_n_coding_frames_per_read_percentages": frame_percentages,
            "peptide_bloom_filter": self.peptide_bloom_filter_filename,
            "peptide_alphabet": self.alphabet,
            "peptide_ksize": self.peptide_ksize,
            "jaccard_threshold": self.jaccard_threshold,
        }

    def get_n_per_coding_category(self):
        """One category per read: Coding > too-short peptide > too-short read > the single category > Non-coding."""
        counts = self.make_empty_coding_categories()
        too_short_peptide = PROTEIN_CODING_CATEGORIES["too_short_peptide"]
        too_short_read = PROTEIN_CODING_CATEGORIES["too_short_nucleotide"]
        for _read_id, group in groupby(zip(self.read_ids, self.categories), key=lambda pair: pair[0]):
            present = sorted(set(category for _, category in group))
            if "Coding" in present:
                counts["Coding"] += 1
            elif too_short_peptide in present:
                counts[too_short_peptide] += 1
            elif too_short_read in present:
                counts[too_short_read] += 1
            elif len(present) == 1:
                counts[present[0]] += 1
            else:
                counts["Non-coding"] += 1
        total = sum(counts.values())
        return {category: 100 * n / total for category, n in counts.items()}, counts

    def get_n_translated_frames_per_read(self):
        coding_per_read = Counter(r for r, c in zip(self.read_ids, self.categories) if c == "Coding")
        histogram = Counter(coding_per_read.values())
        total = sum(histogram.values())
        label = "Number of reads with {} putative protein-coding translations".format
        return ({label(k): 100 * v / total for k, v in histogram.items()}, {label(k): v for k, v in histogram.items()})
