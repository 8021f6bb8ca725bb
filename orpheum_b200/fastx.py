"""FASTA / FASTQ(.gz/.bz2) reading with screed's record semantics (the reference calls ``screed.open``
at translate.py:370 and index.py:83): ``name`` is the full header line, ``sequence`` is returned
untouched (no upper-casing), an empty file yields nothing, anything else raises ValueError.
"""
import bz2
import gzip
import io

import numpy as np


def _open_text(path):
    with io.open(path, "rb") as f:
        magic = f.read(3)
    if magic[:2] == b"\x1f\x8b":
        return gzip.open(path, "rt", encoding="latin-1")
    if magic == b"BZh":
        return bz2.open(path, "rt", encoding="latin-1")
    return io.open(path, "rt", encoding="latin-1")


class Record(dict):
    __getattr__ = dict.__getitem__


class _Reader:
    def __init__(self, path):
        self._f = _open_text(path)
        first = self._f.read(1)
        if first not in (">", "@", ""):
            self._f.close()
            raise ValueError(f"unknown file format for '{path}'")
        self._kind = first
        self._f.seek(0)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def close(self):
        self._f.close()

    def __iter__(self):
        f = self._f
        if self._kind == ">":
            name, parts = None, []
            for line in f:
                line = line.rstrip("\r\n")
                if line[:1] == ">":
                    if name is not None:
                        yield Record(name=name, sequence="".join(parts))
                    name, parts = line[1:], []
                else:
                    parts.append(line)
            if name is not None:
                yield Record(name=name, sequence="".join(parts))
        elif self._kind == "@":
            while True:
                header = f.readline()
                if not header:
                    return
                seq = f.readline().rstrip("\r\n")
                f.readline()
                qual = f.readline().rstrip("\r\n")
                yield Record(name=header.rstrip("\r\n")[1:], sequence=seq, quality=qual)


def open(path):  # noqa: A001 - same public name as screed.open
    return _Reader(path)


def read_batches(path, max_records=1 << 20, max_bases=1 << 28):
    """Yield (names, flat uint8 bytes, uint64 offsets[n+1]) batches in file order, in the layout the C ABI takes."""
    names, seqs, n_bases = [], [], 0
    with open(path) as records:
        for rec in records:
            names.append(rec["name"])
            s = rec["sequence"].encode("latin-1")
            seqs.append(s)
            n_bases += len(s)
            if len(names) >= max_records or n_bases >= max_bases:
                yield _pack(names, seqs)
                names, seqs, n_bases = [], [], 0
    if names:
        yield _pack(names, seqs)


def _pack(names, seqs):
    offsets = np.zeros(len(seqs) + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum([len(s) for s in seqs])
    flat = np.frombuffer(b"".join(seqs), dtype=np.uint8) if offsets[-1] else np.zeros(0, dtype=np.uint8)
    return names, flat, offsets
