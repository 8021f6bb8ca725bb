"""Row-list front end of the output writers, kept for users of the reference's object interface
(``CreateSaveSummary(filenames, csv, parquet, json_summary, filter_name, alphabet, k, threshold, rows)``,
``orpheum/create_save_summary.py:19-134``).  The CLI itself writes from arrays (``columnar.ScoreTable``); this
adapter turns the six-element rows ``[read_id, jaccard, n_kmers, category, frame, filename]`` into the same
column form and shares the summary arithmetic with it.
"""
import csv as _csv
import json
import os

import numpy as np

from . import columnar
from .constants_translate import SCORING_DF_COLUMNS, category_text

_COLUMN_ATTRS = ("read_ids", "jaccard_in_peptide_dbs", "n_kmers", "categories", "translation_frames", "filenames")


class CreateSaveSummary:
    def __init__(self, filenames, csv, parquet, json_summary, peptide_bloom_filter_filename, alphabet, peptide_ksize,
                 jaccard_threshold, coding_scores):
        self.input_filenames = [filenames] if type(filenames) is str else filenames
        self.unique_filenames = [os.path.basename(f) for f in self.input_filenames]
        self.csv, self.parquet, self.json_summary = csv, parquet, json_summary
        self.peptide_bloom_filter_filename = peptide_bloom_filter_filename
        self.alphabet, self.peptide_ksize, self.jaccard_threshold = alphabet, peptide_ksize, jaccard_threshold
        self.coding_scores = coding_scores
        if coding_scores != []:
            for attr, column in zip(_COLUMN_ATTRS, zip(*coding_scores)):
                setattr(self, attr, list(column))

    def maybe_write_csv(self):
        if self.csv:
            with open(self.csv, "w", encoding="utf-8", errors="surrogateescape") as handle:
                writer = _csv.writer(handle, lineterminator="\n")
                writer.writerow(SCORING_DF_COLUMNS)
                writer.writerows(self.coding_scores)

    def maybe_write_parquet(self):
        if self.parquet:
            import pyarrow as pa
            import pyarrow.parquet as pq

            arrays = [pa.array(getattr(self, attr)) for attr in _COLUMN_ATTRS]
            pq.write_table(pa.Table.from_arrays(arrays, names=SCORING_DF_COLUMNS), self.parquet)

    def make_empty_coding_categories(self):
        return columnar.empty_categories(self.alphabet)

    def generate_coding_summary(self):
        code_of = {}
        for code in range(6):
            try:
                code_of[category_text(code, self.alphabet)] = code
            except KeyError:  # alphabets the reference has no low-complexity label for
                pass
        codes = np.fromiter((code_of[c] for c in self.categories), dtype=np.int64, count=len(self.categories))
        rows = (columnar.consecutive_groups(self.read_ids), codes,
                [r for r, c in zip(self.read_ids, self.categories) if c == "Coding"],
                np.asarray(self.jaccard_in_peptide_dbs, dtype=np.float64))
        return columnar.summarize(self.alphabet, rows, self.input_filenames, self.peptide_bloom_filter_filename,
                                  self.peptide_ksize, self.jaccard_threshold)

    def maybe_write_json_summary(self):
        if not self.json_summary:
            return None
        if self.coding_scores == []:
            summary = columnar.summarize(self.alphabet, None, self.input_filenames, self.peptide_bloom_filter_filename,
                                         self.peptide_ksize, self.jaccard_threshold)
        else:
            summary = self.generate_coding_summary()
        with open(self.json_summary, "w") as handle:
            json.dump(summary, fp=handle)
        if self.coding_scores != []:  # the reference frees the columns once the summary is written
            for attr in _COLUMN_ATTRS + ("coding_scores",):
                delattr(self, attr)
        return summary
