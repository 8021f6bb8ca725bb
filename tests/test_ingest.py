"""CPU tests of the fused host ingest (orph_ingest_*: parallel FASTQ parse + 2-bit pack in one pass): the batches must
hold exactly the records the general reader (fastx.py / fastx.cpp, screed's semantics) yields, packed exactly as the
SIMD packer packs them, for every file layout and for every way the text can be cut into windows and segments."""
import ctypes
import gzip
import os

import numpy as np
import pytest

from orpheum_b200 import _lib, build, fastx, gpu


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def ingest_all(lib, path, max_reads=0, threads=4):
    """-> (names, sequences, flagged read indices), decoded from the packed stream where the read is packable."""
    h = ctypes.c_void_p()
    rc = lib.orph_ingest_open(os.fsencode(path), threads, ctypes.byref(h))
    if rc == _lib.ORPH_E_INVALID:
        raise ValueError(_lib.last_error())
    if rc != 0:
        raise OSError(_lib.last_error())
    names, seqs, flagged = [], [], []
    raw = _lib.IngestBatch()
    try:
        while True:
            rc = lib.orph_ingest_next(h, max_reads, ctypes.byref(raw))
            if rc != 0:
                raise OSError(_lib.last_error())
            m = int(raw.n_reads)
            if m == 0:
                break
            arr = lambda p, t, k: np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(t)), shape=(k,)) if k else np.zeros(0, dtype=t)
            off = arr(raw.offsets32, ctypes.c_uint32, m + 1).astype(np.int64)
            total = int(raw.total_bases)
            assert off[0] == 0 and off[-1] == total
            packed = arr(raw.packed, ctypes.c_uint64, (total + 31) // 32)
            noff = arr(raw.name_offsets, ctypes.c_uint64, m + 1).astype(np.int64)
            blob = ctypes.string_at(raw.names, int(noff[-1]))
            start = arr(raw.seq_start, ctypes.c_uint64, m)
            length = arr(raw.seq_len, ctypes.c_uint32, m)
            assert (np.diff(off) == length).all() and int(raw.max_read_len) == int(length.max())
            fl = set(arr(raw.flagged, ctypes.c_uint32, int(raw.n_flagged)).tolist())
            idx = np.arange(total, dtype=np.int64)
            codes = ((packed[idx // 32] >> (2 * (idx % 32)).astype(np.uint64)) & np.uint64(3)).astype(np.uint8) if total else np.zeros(0, np.uint8)
            decoded = np.frombuffer(b"ACTG", dtype=np.uint8)[codes].tobytes()
            for i in range(m):
                text_bytes = ctypes.string_at(raw.text + int(start[i]), int(length[i]))
                if i in fl:
                    assert any(ch not in b"ACGT" for ch in text_bytes)
                    flagged.append(len(seqs))
                else:
                    assert decoded[off[i]:off[i + 1]] == text_bytes
                seqs.append(text_bytes.decode("latin-1"))
                names.append(blob[noff[i]:noff[i + 1]].decode("latin-1"))
    finally:
        lib.orph_ingest_close(h)
    return names, seqs, flagged


def reference_records(path):
    recs = list(fastx._Reader(path))  # the pure-Python reader
    return [r["name"] for r in recs], [r["sequence"] for r in recs]


def random_reads(rng, n, lo=1, hi=400, dirty=0.01):
    seqs = []
    for _ in range(n):
        s = bytearray(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), int(rng.integers(lo, hi))).tobytes())
        if rng.random() < dirty:
            s[int(rng.integers(0, len(s)))] = int(rng.choice(np.frombuffer(b"Nacgtnry", dtype=np.uint8)))
        seqs.append(bytes(s))
    return seqs


def write_fastq(path, seqs, rng, eol=b"\n", final_newline=True, opener=open, plus_repeats_name=False):
    parts = []
    for i, s in enumerate(seqs):
        name = b"read_%d len=%d  x" % (i, len(s))
        q = rng.integers(33, 74, len(s)).astype(np.uint8).tobytes()  # any quality byte, '@' and '+' included
        parts.append(b"@" + name + eol + s + eol + b"+" + (name if plus_repeats_name else b"") + eol + q + eol)
    blob = b"".join(parts)
    if not final_newline:
        blob = blob[: -len(eol)]
    with opener(path, "wb") as f:
        f.write(blob)


def bgzf_opener(path, mode):
    """file-like whose content lands in `path` as BGZF when closed"""
    import io

    from orpheum_b200.synth import write_bgzf

    class _Buf(io.BytesIO):
        def close(self):
            write_bgzf(path, self.getvalue())
            super().close()

    return _Buf()


@pytest.mark.parametrize("layout", ["plain", "crlf", "no_final_newline", "gz", "bgzf", "plus_repeats_name", "tiny_reads", "long_reads"])
def test_fused_ingest_matches_reader(lib, tmp_path, layout):
    rng = np.random.default_rng(hash(layout) % 1000)
    path = str(tmp_path / ("r.fq.gz" if layout in ("gz", "bgzf") else "r.fq"))
    if layout == "tiny_reads":
        seqs = random_reads(rng, 300_000, 1, 6)  # hundreds of thousands of records per 1 MiB segment
    elif layout == "long_reads":
        seqs = random_reads(rng, 40, 200_000, 3_000_000, dirty=0.2)  # records that span several segments
    else:
        seqs = random_reads(rng, 60_000)
    write_fastq(path, seqs, rng, eol=b"\r\n" if layout == "crlf" else b"\n", final_newline=layout != "no_final_newline",
                opener=gzip.open if layout == "gz" else bgzf_opener if layout == "bgzf" else open, plus_repeats_name=layout == "plus_repeats_name")
    want_names, want_seqs = reference_records(path)
    assert want_seqs == [s.decode() for s in seqs]
    for max_reads, threads in ((0, 4), (5000, 1), (20_000, 3)):
        names, got, flagged = ingest_all(lib, path, max_reads, threads)
        assert got == want_seqs and names == want_names, (layout, max_reads)
        assert flagged == [i for i, s in enumerate(seqs) if any(ch not in b"ACGT" for ch in s)]


def test_general_layouts_and_errors(lib, tmp_path):
    rng = np.random.default_rng(5)
    seqs = random_reads(rng, 3000, 30, 300)
    # wrapped FASTQ, a strict head followed by wrapped records, FASTA
    wrapped = str(tmp_path / "wrapped.fq")
    with open(wrapped, "wb") as f:
        for i, s in enumerate(seqs):
            w = 1000 if i < 1500 else 50
            q = b"I" * len(s)
            f.write(b"@w%d\n" % i + b"\n".join(s[j:j + w] for j in range(0, len(s), w)) + b"\n+\n" + b"\n".join(q[j:j + w] for j in range(0, len(q), w)) + b"\n")
    fasta = str(tmp_path / "p.fa")
    with open(fasta, "wb") as f:
        for i, s in enumerate(seqs):
            f.write(b">p%d desc\n" % i + b"\n".join(s[j:j + 60] for j in range(0, len(s), 60)) + b"\n")
    for path in (wrapped, fasta):
        want_names, want_seqs = reference_records(path)
        names, got, _ = ingest_all(lib, path, 700, 2)
        assert got == want_seqs and names == want_names
    empty = str(tmp_path / "empty.fq")
    open(empty, "wb").close()
    assert ingest_all(lib, empty) == ([], [], [])
    junk = str(tmp_path / "junk.txt")
    with open(junk, "w") as f:
        f.write("hello\n")
    with pytest.raises(ValueError):
        ingest_all(lib, junk)
    with pytest.raises(OSError):
        ingest_all(lib, str(tmp_path / "missing.fq"))
    truncated = str(tmp_path / "trunc.fq")
    with open(truncated, "w") as f:
        f.write("@a\nACGT\n+\nIIII\n@b\nACGTACGT\n+\nIII")
    with pytest.raises(OSError):
        ingest_all(lib, truncated)


def test_bgzf_damage_and_odd_files(lib, tmp_path):
    """Blocked gzip is inflated member by member in parallel: a flipped byte inside a member is an I/O error (CRC), a file
    without the empty end-of-file member and a file of one tiny member still parse, long reads span many members."""
    from orpheum_b200.synth import write_bgzf

    rng = np.random.default_rng(8)
    seqs = random_reads(rng, 20_000, 50, 300) + random_reads(rng, 3, 150_000, 400_000, dirty=0)
    plain = str(tmp_path / "r.fq")
    write_fastq(plain, seqs, rng)
    data = open(plain, "rb").read()
    want_names, want_seqs = reference_records(plain)
    for name, kw in (("full.fq.gz", {}), ("no_eof.fq.gz", {"eof_marker": False}), ("small_blocks.fq.gz", {"block": 997})):
        path = str(tmp_path / name)
        write_bgzf(path, data, **kw)
        assert gzip.open(path, "rb").read() == data  # a valid multi-member gzip file
        for max_reads, threads in ((0, 4), (3000, 2)):
            names, got, _ = ingest_all(lib, path, max_reads, threads)
            assert got == want_seqs and names == want_names, name
    tiny = str(tmp_path / "tiny.fq.gz")
    write_bgzf(tiny, b"@a\nACGT\n+\nIIII\n")
    assert ingest_all(lib, tiny)[:2] == (["a"], ["ACGT"])
    damaged = str(tmp_path / "damaged.fq.gz")
    blob = bytearray(open(str(tmp_path / "full.fq.gz"), "rb").read())
    blob[len(blob) // 2] ^= 0x5A
    open(damaged, "wb").write(bytes(blob))
    with pytest.raises(OSError):
        ingest_all(lib, damaged)


def test_gather_reads(lib, tmp_path):
    rng = np.random.default_rng(11)
    seqs = random_reads(rng, 5000, 20, 200)
    path = str(tmp_path / "g.fq")
    write_fastq(path, seqs, rng)
    h = ctypes.c_void_p()
    assert lib.orph_ingest_open(os.fsencode(path), 2, ctypes.byref(h)) == 0
    raw = _lib.IngestBatch()
    assert lib.orph_ingest_next(h, 0, ctypes.byref(raw)) == 0 and raw.n_reads == 5000
    index = np.array([4999, 0, 17, 17, 2500], dtype=np.uint64)
    offsets = np.zeros(len(index) + 1, dtype=np.uint64)
    assert lib.orph_gather_reads(raw.text, raw.seq_start, raw.seq_len, _lib.ptr(index), len(index), None, _lib.ptr(offsets)) == 0
    out = np.zeros(int(offsets[-1]), dtype=np.uint8)
    assert lib.orph_gather_reads(raw.text, raw.seq_start, raw.seq_len, _lib.ptr(index), len(index), _lib.ptr(out), _lib.ptr(offsets)) == 0
    assert out.tobytes() == b"".join(seqs[int(i)] for i in index)
    lib.orph_ingest_close(h)
