"""screed.open stand-in (pinned screed==1.0.4 / 1.0): FASTA/FASTQ(.gz) record iterator.

``name`` is the full header line (no split at whitespace), ``sequence`` is returned as is.
A file that is neither FASTA nor FASTQ raises ValueError; an empty file yields nothing.
"""
import gzip
import io


class _Record(dict):
    def __getattr__(self, k):
        return self[k]


class _Reader:
    def __init__(self, path):
        with io.open(path, "rb") as f:
            magic = f.read(2)
        self._f = gzip.open(path, "rt") if magic == b"\x1f\x8b" else io.open(path, "rt", errors="strict")
        try:
            first = self._f.read(1)
        except UnicodeDecodeError:
            self._f.close()
            raise ValueError(f"unknown file format for '{path}'")
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
        if self._kind == ">":
            name, chunks = None, []
            for line in self._f:
                line = line.rstrip("\r\n")
                if line.startswith(">"):
                    if name is not None:
                        yield _Record(name=name, sequence="".join(chunks))
                    name, chunks = line[1:], []
                else:
                    chunks.append(line)
            if name is not None:
                yield _Record(name=name, sequence="".join(chunks))
        elif self._kind == "@":
            while True:
                header = self._f.readline()
                if not header:
                    return
                seq = self._f.readline().rstrip("\r\n")
                self._f.readline()
                qual = self._f.readline().rstrip("\r\n")
                yield _Record(name=header.rstrip("\r\n")[1:], sequence=seq, quality=qual)


def open(path):  # noqa: A001 - mirrors screed's public name
    return _Reader(path)
