"""Thin object layer over the C ABI: a device context, the khmer-compatible Bloom filter
(``GpuNodegraph``) and batch scoring.  Everything here calls ``liborpheum_b200.so``; nothing
computes on the host.

``GpuNodegraph`` quacks like ``khmer.Nodegraph`` for the calls the reference makes
(``ksize / get / add / save / n_unique_kmers``, classmethod ``load``; index.py:70-77,101,119,
202-218; translate.py:145,190,198).
"""
import ctypes
import os
import struct
import threading

import numpy as np

from . import _lib
from ._lib import READ_RESULT_DTYPE, READS_ASCII, READS_PACKED2, check, ptr

_default_ctx = None
_default_lock = threading.Lock()


class Context:
    """One GPU + one CUDA stream (``orph_ctx``)."""

    def __init__(self, device=0, stream=None):
        self._lib = _lib.load()
        h = ctypes.c_void_p()
        check(self._lib.orph_ctx_create(int(device), ctypes.c_void_p(stream) if stream else None, ctypes.byref(h)))
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.orph_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        check(self._lib.orph_ctx_synchronize(self._h))

    def check(self):
        check(self._lib.orph_ctx_check(self._h))

    def set_force_exact_dedup(self, on):
        check(self._lib.orph_ctx_set_force_exact_dedup(self._h, int(bool(on))))

    KERNEL_NAMES = ("pack_ascii", "prefilter", "score_frames", "score_bytes", "score_tasks", "finalize_best", "task_sort", "score_long")

    def set_thread_path(self, on):
        """Test hook: False sends short reads through the warp-per-read kernel."""
        check(self._lib.orph_ctx_set_thread_path(self._h, int(bool(on))))

    def set_host_threads(self, n):
        """Host threads that 2-bit pack ASCII batches inside score_reads (<0 auto, 0 never pack on the host)."""
        check(self._lib.orph_ctx_set_host_threads(self._h, int(n)))

    def ingest_stats(self):
        packed, ascii_, threads, rate, h2d = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_int(), ctypes.c_double(), ctypes.c_double()
        check(self._lib.orph_ctx_ingest_stats(self._h, ctypes.byref(packed), ctypes.byref(ascii_), ctypes.byref(threads),
                                              ctypes.byref(rate), ctypes.byref(h2d)))
        return {"chunks_packed_on_host": int(packed.value), "chunks_sent_ascii": int(ascii_.value),
                "host_threads": int(threads.value), "pack_bytes_per_s": float(rate.value), "h2d_bytes_per_s": float(h2d.value)}

    def set_kernel_timing(self, on):
        check(self._lib.orph_ctx_set_kernel_timing(self._h, int(bool(on))))

    def kernel_times(self):
        """{kernel: (total_ms, launches)} since the last call; synchronizes."""
        ms = np.zeros(len(self.KERNEL_NAMES), dtype=np.float64)
        n = np.zeros(len(self.KERNEL_NAMES), dtype=np.uint64)
        check(self._lib.orph_ctx_kernel_times(self._h, ptr(ms), ptr(n)))
        return {name: (float(ms[i]), int(n[i])) for i, name in enumerate(self.KERNEL_NAMES)}

    @property
    def launch_count(self):
        return int(self._lib.orph_ctx_launch_count(self._h))

    def hash_kmers(self, kmers):
        """hash_murmur of equal-length byte strings, on the GPU."""
        kmers = [k.encode("latin-1") if isinstance(k, str) else bytes(k) for k in kmers]
        n = len(kmers)
        out = np.zeros(n, dtype=np.uint64)
        if n == 0:
            return out
        length = len(kmers[0])
        assert all(len(k) == length for k in kmers)
        flat = np.frombuffer(b"".join(kmers), dtype=np.uint8) if length else np.zeros(1, dtype=np.uint8)
        check(self._lib.orph_hash_kmers(self._h, ptr(flat), length, n, ptr(out)))
        return out

    def gather_peak(self, footprint_bytes, loads_per_thread=256, persist_l2=False):
        v = ctypes.c_double()
        check(self._lib.orph_bench_gather_peak(self._h, int(footprint_bytes), int(loads_per_thread), int(persist_l2),
                                               ctypes.byref(v)))
        return v.value


def default_context():
    global _default_ctx
    with _default_lock:
        if _default_ctx is None:
            _default_ctx = Context(int(os.environ.get("LOCAL_RANK", "0")) if os.environ.get("ORPHEUM_B200_USE_LOCAL_RANK") else 0)
        return _default_ctx


def table_sizes(tablesize, n_tables):
    out = np.zeros(int(n_tables), dtype=np.uint64)
    check(_lib.load().orph_table_sizes(int(tablesize), int(n_tables), ptr(out)))
    return [int(x) for x in out]


class GpuNodegraph:
    """khmer.Nodegraph on the GPU: n_tables prime-sized bit tables resident in HBM."""

    def __init__(self, ksize, starting_size, n_tables, ctx=None, _sizes=None, _handle=None):
        self._lib = _lib.load()
        self.ctx = ctx or default_context()
        if _handle is not None:
            self._h = _handle
        else:
            sizes = np.asarray(_sizes if _sizes is not None else table_sizes(int(starting_size), int(n_tables)), dtype=np.uint64)
            h = ctypes.c_void_p()
            check(self._lib.orph_filter_create(self.ctx._h, int(ksize), len(sizes), ptr(sizes), ctypes.byref(h)))
            self._h = h

    def __del__(self):
        try:
            if getattr(self, "_h", None) and getattr(self.ctx, "_h", None):
                self._lib.orph_filter_destroy(self._h)
            self._h = None
        except Exception:
            pass

    # -- khmer API -----------------------------------------------------------------------------
    def ksize(self):
        k = ctypes.c_uint32()
        check(self._lib.orph_filter_info(self._h, ctypes.byref(k), None, None))
        return int(k.value)

    def n_tables(self):
        n = ctypes.c_uint32()
        check(self._lib.orph_filter_info(self._h, None, ctypes.byref(n), None))
        return int(n.value)

    def hashsizes(self):
        sizes = np.zeros(_lib.MAX_TABLES, dtype=np.uint64)
        n = ctypes.c_uint32()
        check(self._lib.orph_filter_info(self._h, None, ctypes.byref(n), ptr(sizes)))
        return [int(x) for x in sizes[: n.value]]

    def n_unique_kmers(self):
        return int(self._lib.orph_filter_n_unique_kmers(self._h))

    def n_occupied(self, table=0):
        v = ctypes.c_uint64()
        check(self._lib.orph_filter_popcount(self._h, int(table), ctypes.byref(v)))
        return int(v.value)

    def add(self, hashed):
        """Nodegraph.add(int): True if at least one bit was newly set."""
        h = np.array([int(hashed) & 0xFFFFFFFFFFFFFFFF], dtype=np.uint64)
        new = np.zeros(1, dtype=np.uint8)
        check(self._lib.orph_filter_add_hashes(self._h, ptr(h), 1, ptr(new)))
        return bool(new[0])

    def add_many(self, hashes):
        h = np.ascontiguousarray(hashes, dtype=np.uint64)
        new = np.zeros(len(h), dtype=np.uint8)
        check(self._lib.orph_filter_add_hashes(self._h, ptr(h), len(h), ptr(new)))
        return new.astype(bool)

    def get(self, hashed):
        """Nodegraph.get(int) -> 0 / 1."""
        return int(self.get_many([int(hashed) & 0xFFFFFFFFFFFFFFFF])[0])

    def get_many(self, hashes):
        h = np.ascontiguousarray(hashes, dtype=np.uint64)
        out = np.zeros(len(h), dtype=np.uint8)
        check(self._lib.orph_filter_query_hashes(self._h, ptr(h), len(h), ptr(out)))
        return out

    def add_peptides(self, residues, offsets, lut, count_unique=True):
        """The inner loop of index.make_peptide_bloom_filter for a batch of records (index.py:107-122).
        count_unique=False skips the bookkeeping behind n_unique_kmers() (same bits, faster build)."""
        residues = np.ascontiguousarray(residues, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        lut = np.ascontiguousarray(lut, dtype=np.uint8)
        assert lut.shape == (256,)
        inc = ctypes.c_uint64()
        check(self._lib.orph_filter_add_peptides(self._h, ptr(residues), ptr(offsets), len(offsets) - 1, ptr(lut),
                                                 ctypes.byref(inc) if count_unique else None))
        return int(inc.value)

    def tables(self):
        """Host copies of the tables, byte for byte as in the .nodegraph file."""
        sizes = self.hashsizes()
        tabs = [np.zeros(s // 8 + 1, dtype=np.uint8) for s in sizes]
        arr = (ctypes.c_void_p * len(tabs))(*[t.ctypes.data for t in tabs])
        occ = ctypes.c_uint64()
        check(self._lib.orph_filter_to_host(self._h, arr, ctypes.byref(occ)))
        return tabs, int(occ.value)

    def save(self, path):
        """Write an OXLI v4 nodegraph file (khmer format), byte-identical to khmer's."""
        tabs, occupied = self.tables()
        sizes = self.hashsizes()
        with open(path, "wb") as f:
            f.write(b"OXLI" + bytes([4, 2]) + struct.pack("<IBQ", self.ksize(), len(sizes), occupied))
            for size, tab in zip(sizes, tabs):
                f.write(struct.pack("<Q", size))
                f.write(tab.tobytes())

    @classmethod
    def load(cls, path, ctx=None):
        """khmer.Nodegraph.load / khmer.load_nodegraph: OSError on anything that is not a nodegraph file."""
        with open(path, "rb") as f:
            head = f.read(19)
            if len(head) != 19 or head[:4] != b"OXLI" or head[4] != 4 or head[5] != 2:
                raise OSError(f"{path} is not a khmer nodegraph (OXLI v4, type 2) file")
            ksize, n_tables, _occupied = struct.unpack("<IBQ", head[6:])
            if not 1 <= n_tables <= _lib.MAX_TABLES:
                raise OSError(f"{path}: unsupported number of tables {n_tables}")
            sizes, tabs = [], []
            for _ in range(n_tables):
                raw = f.read(8)
                if len(raw) != 8:
                    raise OSError(f"{path} is truncated")
                (size,) = struct.unpack("<Q", raw)
                data = np.frombuffer(f.read(size // 8 + 1), dtype=np.uint8)
                if len(data) != size // 8 + 1:
                    raise OSError(f"{path} is truncated")
                sizes.append(size)
                tabs.append(np.ascontiguousarray(data))
        ctx = ctx or default_context()
        lib = _lib.load()
        sizes_a = np.asarray(sizes, dtype=np.uint64)
        arr = (ctypes.c_void_p * n_tables)(*[t.ctypes.data for t in tabs])
        h = ctypes.c_void_p()
        check(lib.orph_filter_from_host(ctx._h, ksize, n_tables, ptr(sizes_a), arr, ctypes.byref(h)))
        return cls(None, None, None, ctx=ctx, _handle=h)

    # -- replication ---------------------------------------------------------------------------
    def device_buffer(self):
        """(device pointer, n_bytes, table byte offsets) of the one contiguous allocation."""
        p = ctypes.c_void_p()
        n = ctypes.c_uint64()
        offs = np.zeros(_lib.MAX_TABLES, dtype=np.uint64)
        check(self._lib.orph_filter_device_buffer(self._h, ctypes.byref(p), ctypes.byref(n), ptr(offs)))
        return int(p.value), int(n.value), [int(x) for x in offs[: self.n_tables()]]

    def replicate(self, dst_ctx):
        h = ctypes.c_void_p()
        check(self._lib.orph_filter_replicate(self._h, dst_ctx._h, ctypes.byref(h)))
        return GpuNodegraph(None, None, None, ctx=dst_ctx, _handle=h)


# -- scoring ---------------------------------------------------------------------------------------
def concat_reads(sequences):
    """list of str/bytes -> (uint8 flat, uint64 offsets[n+1])."""
    seqs = [s.encode("utf-8", "surrogateescape") if isinstance(s, str) else bytes(s) for s in sequences]
    offsets = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if seqs:
        offsets[1:] = np.cumsum([len(s) for s in seqs])
    flat = np.frombuffer(b"".join(seqs), dtype=np.uint8) if seqs else np.zeros(0, dtype=np.uint8)
    return np.ascontiguousarray(flat), offsets


def pack_reads(flat, offsets, n_threads=0):
    """Host 2-bit packer: (uint64 words, uint8 needs_ascii[n])."""
    flat = np.ascontiguousarray(flat, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    total = int(offsets[-1] - offsets[0]) if n > 0 else 0
    packed = np.zeros((total + 31) // 32 + 1, dtype=np.uint64)
    needs = np.zeros(max(n, 1), dtype=np.uint8)
    check(_lib.load().orph_pack_reads(ptr(flat), ptr(offsets), n, ptr(packed), ptr(needs), int(n_threads)))
    return packed, needs[:n]


def score_reads(graph, reads, offsets, lut, jaccard_threshold, reads_format=READS_ASCII, ctx=None):
    """One C call for a batch held in HOST memory -> structured array of orph_read_result.

    ``reads`` is the concatenated ASCII byte stream (READS_ASCII) or the 2-bit stream (READS_PACKED2)."""
    ctx = ctx or graph.ctx
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    lut = np.ascontiguousarray(lut, dtype=np.uint8)
    n = len(offsets) - 1
    out = np.zeros(n, dtype=READ_RESULT_DTYPE)
    if reads_format == READS_ASCII:
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
    else:
        reads = np.ascontiguousarray(reads, dtype=np.uint64)
    rc = _lib.load().orph_score_reads(ctx._h, graph._h, ptr(reads) if reads.size else ptr(np.zeros(1, np.uint8)),
                                      int(reads_format), ptr(offsets), n, ptr(lut), float(jaccard_threshold), ptr(out))
    check(rc)
    return out


def score_reads_device(graph, d_reads_ptr, d_offsets_ptr, n_reads, max_read_len, lut, jaccard_threshold, d_out_ptr,
                       reads_format=READS_PACKED2, ctx=None):
    """Enqueue scoring of a batch already resident in HBM (raw device pointers, e.g. tensor.data_ptr())."""
    ctx = ctx or graph.ctx
    lut = np.ascontiguousarray(lut, dtype=np.uint8)
    check(_lib.load().orph_score_reads_device(ctx._h, graph._h, ctypes.c_void_p(d_reads_ptr), int(reads_format),
                                              ctypes.c_void_p(d_offsets_ptr), int(n_reads), int(max_read_len), ptr(lut),
                                              float(jaccard_threshold), ctypes.c_void_p(d_out_ptr)))


def pack_reads_device(ctx, d_ascii_ptr, d_offsets_ptr, n_reads, total_bases, d_packed_ptr, d_nmask_ptr, d_has_n_ptr, d_needs_bytes_ptr):
    """ASCII reads resident in HBM -> 2-bit words + N plane + per-read "holds N" / "needs its bytes" flags (raw device
    pointers; orph_pack_reads_device).  Only enqueues."""
    check(_lib.load().orph_pack_reads_device(ctx._h, ctypes.c_void_p(d_ascii_ptr), ctypes.c_void_p(d_offsets_ptr), int(n_reads), int(total_bases),
                                             ctypes.c_void_p(d_packed_ptr), ctypes.c_void_p(d_nmask_ptr), ctypes.c_void_p(d_has_n_ptr),
                                             ctypes.c_void_p(d_needs_bytes_ptr)))


def score_reads_device_n(graph, d_packed_ptr, d_nmask_ptr, d_has_n_ptr, d_offsets_ptr, n_reads, max_read_len, lut, jaccard_threshold, d_out_ptr, ctx=None):
    """score_reads_device for 2-bit reads with an N plane (orph_score_reads_device_n)."""
    ctx = ctx or graph.ctx
    lut = np.ascontiguousarray(lut, dtype=np.uint8)
    check(_lib.load().orph_score_reads_device_n(ctx._h, graph._h, ctypes.c_void_p(d_packed_ptr), ctypes.c_void_p(d_nmask_ptr), ctypes.c_void_p(d_has_n_ptr),
                                                ctypes.c_void_p(d_offsets_ptr), int(n_reads), int(max_read_len), ptr(lut), float(jaccard_threshold),
                                                ctypes.c_void_p(d_out_ptr)))


class StreamBatch:
    """One scored batch of a ScoringStream.  ``results`` is a view into the library's pinned memory and, like every other
    pointer here, only valid until the stream moves on: copy what must outlive the iteration."""

    def __init__(self, raw):
        self._raw = raw
        self.n_reads = int(raw.n_reads)
        self.n_byte_path = int(raw.n_byte_path)
        self.status = int(raw.status)
        self.results = np.ctypeslib.as_array(ctypes.cast(raw.results, ctypes.POINTER(ctypes.c_uint8)),
                                             shape=(self.n_reads * READ_RESULT_DTYPE.itemsize,)).view(READ_RESULT_DTYPE)

    def copy_results_into(self, out, at=0):
        """Copies this batch's records to ``out[at : at + n_reads]`` (a READ_RESULT_DTYPE array) as one memmove -- numpy's
        element-wise structured assignment is ten times slower than the bytes take to move."""
        if out.dtype != READ_RESULT_DTYPE or not out.flags["C_CONTIGUOUS"] or at < 0 or at + self.n_reads > len(out):
            raise ValueError("copy_results_into: out must be a contiguous READ_RESULT_DTYPE array with room for the batch")
        ctypes.memmove(out.ctypes.data + at * READ_RESULT_DTYPE.itemsize, self._raw.results, self.n_reads * READ_RESULT_DTYPE.itemsize)
        return at + self.n_reads

    def names(self):
        """fastx.NameList (owns copies of the name bytes and offsets)."""
        from .fastx import NameList

        n = self.n_reads
        offsets = np.ctypeslib.as_array(ctypes.cast(self._raw.name_offsets, ctypes.POINTER(ctypes.c_uint64)), shape=(n + 1,)).copy()
        return NameList(ctypes.string_at(self._raw.names, int(offsets[-1])) if offsets[-1] else b"", offsets)

    def reads(self, index=None):
        """(flat uint8 bytes, uint64 offsets) of the reads ``index`` (default: all of them), as the file has them."""
        index = np.arange(self.n_reads, dtype=np.uint64) if index is None else np.ascontiguousarray(index, dtype=np.uint64)
        offsets = np.zeros(len(index) + 1, dtype=np.uint64)
        lib = _lib.load()
        raw = self._raw
        check(lib.orph_gather_reads(raw.text, raw.seq_start, raw.seq_len, ptr(index), len(index), None, ptr(offsets)))
        flat = np.empty(max(int(offsets[-1]), 1), dtype=np.uint8)
        check(lib.orph_gather_reads(raw.text, raw.seq_start, raw.seq_len, ptr(index), len(index), ptr(flat), ptr(offsets)))
        return flat[: int(offsets[-1])], offsets


class ScoringStream:
    """``for batch in ScoringStream(graph, path, lut, threshold)``: the file is parsed, 2-bit packed, copied and scored by
    the native library (orph_stream_*); Python sees one finished StreamBatch per iteration.  Stands in for the loop over
    ``screed.open(reads)`` in Translate.score_reads_per_file (translate.py:367-372)."""

    def __init__(self, graph, path, lut, jaccard_threshold, batch_reads=1 << 20, n_threads=0, ctx=None, replicas=None):
        """``replicas``: extra (Context, GpuNodegraph) lanes -- replicas of ``graph`` on other GPUs (``graph.replicate``) or
        on further contexts of the same GPU; batch k is scored on lane k % n_lanes, batches still arrive in file order."""
        self._lib = _lib.load()
        self.ctx = ctx or graph.ctx
        lut = np.ascontiguousarray(lut, dtype=np.uint8)
        h = ctypes.c_void_p()
        lanes = [(self.ctx, graph)] + list(replicas or [])
        ctxs = (ctypes.c_void_p * len(lanes))(*[c._h for c, _ in lanes])
        filters = (ctypes.c_void_p * len(lanes))(*[g._h for _, g in lanes])
        self._lanes = lanes  # contexts and filters must outlive the stream
        rc = self._lib.orph_stream_open_multi(ctxs, filters, len(lanes), os.fsencode(path), ptr(lut), float(jaccard_threshold), int(batch_reads),
                                              int(n_threads), ctypes.byref(h))
        if rc == _lib.ORPH_E_INVALID and "unknown file format" in _lib.last_error():
            raise ValueError(f"unknown file format for '{path}'")
        if rc == _lib.ORPH_E_IO:
            raise OSError(_lib.last_error())
        check(rc)
        self._h = h
        self._graph = graph  # the filter must outlive the stream

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def close(self):
        if getattr(self, "_h", None):
            self._lib.orph_stream_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def emit(self, ctx, kinds_mask, copy=True):
        """FASTA text of the batch the iteration currently holds, formatted by the library (orph_stream_emit): bit k of
        ``kinds_mask`` asks for kind k of ``format_fasta``.  -> list of 4 ``bytes`` (empty where nothing was asked or emitted).
        ``ctx`` runs the peptide translation: a context of the caller's, e.g. the filter's.  ``copy=False``: memoryviews of the
        library's own buffers instead (valid until the next call on this stream) -- what a caller that only writes them out
        wants; the coding peptides of a batch are a hundred megabytes."""
        ptrs = (ctypes.c_void_p * 4)()
        lens = (ctypes.c_uint64 * 4)()
        check(self._lib.orph_stream_emit(self._h, ctx._h, int(kinds_mask), ctypes.byref(ptrs), ctypes.byref(lens)))
        if copy:
            return [ctypes.string_at(ptrs[k], int(lens[k])) if lens[k] else b"" for k in range(4)]
        return [memoryview((ctypes.c_char * int(lens[k])).from_address(ptrs[k])).cast("B") if lens[k] else b"" for k in range(4)]

    def __iter__(self):
        raw = _lib.StreamBatch()
        while True:
            rc = self._lib.orph_stream_next(self._h, ctypes.byref(raw))
            if rc == _lib.ORPH_E_IO:
                raise OSError(_lib.last_error())
            check(rc)
            if raw.n_reads == 0:
                return
            yield StreamBatch(raw)


def device_count():
    """GPUs visible to this process (cudaGetDeviceCount through the library: no torch import, no context created)."""
    return int(_lib.load().orph_device_count())


def replica_lanes(graph, n_lanes, devices=None):
    """(Context, GpuNodegraph) lanes for ScoringStream(replicas=...): the filter replicated over NVLink onto the other GPUs
    of the box (``devices``: explicit device indices, repeats allowed -- two contexts on one GPU)."""
    if devices is None:
        n_dev = max(1, device_count())
        devices = [(graph.ctx.device + i) % n_dev for i in range(1, n_lanes)]
    lanes = []
    for d in devices:
        ctx = Context(int(d))
        lanes.append((ctx, graph.replicate(ctx)))
    return lanes


def translate_frames(ctx, reads, offsets, item_read, item_frame):
    """Un-encoded peptides of the listed (read index, frame key) pairs, translated on the GPU with the
    reference's codon semantics -> list of str in item order."""
    out, out_offsets = translate_frames_raw(ctx, reads, offsets, item_read, item_frame)
    blob = out.tobytes()
    return [blob[int(out_offsets[i]):int(out_offsets[i + 1])].decode("latin-1") for i in range(len(out_offsets) - 1)]


def format_fasta(kind, names_blob, name_offsets, results, reads, read_offsets, peptides, pep_offsets, peptides_include_low_complexity):
    """orph_format_fasta -> bytes (see include/orpheum_b200.h)."""
    lib = _lib.load()
    names_blob = np.ascontiguousarray(names_blob, dtype=np.uint8)
    name_offsets = np.ascontiguousarray(name_offsets, dtype=np.uint64)
    results = np.ascontiguousarray(results)
    n = len(results)
    need = ctypes.c_uint64()
    args = [int(kind), ptr(names_blob), ptr(name_offsets), n, ptr(results), ptr(reads) if reads is not None else None,
            ptr(read_offsets) if read_offsets is not None else None, ptr(peptides) if peptides is not None else None,
            ptr(pep_offsets) if pep_offsets is not None else None, int(bool(peptides_include_low_complexity))]
    # one formatting pass in the usual case: a buffer sized from the inputs (header = name + ~60 bytes of decoration per
    # emitted frame), a second pass only if that guess was too small
    n_items = len(pep_offsets) - 1 if pep_offsets is not None else 0
    per_name = (len(names_blob) // max(n, 1)) + 64
    guess = n_items * per_name + (len(peptides) if peptides is not None else 0) + 4096
    if kind == 1:
        guess += 2 * int(len(reads))
    elif kind == 2:  # up to six records per read, each with the whole read (untouched pages of np.empty cost nothing)
        guess += 6 * (int(len(reads)) + n * per_name)
    out = np.empty(max(guess, 1), dtype=np.uint8)
    check(lib.orph_format_fasta(*args, ptr(out), len(out), ctypes.byref(need)))
    if int(need.value) > len(out):
        out = np.empty(int(need.value), dtype=np.uint8)
        check(lib.orph_format_fasta(*args, ptr(out), len(out), ctypes.byref(need)))
    return out[: int(need.value)].tobytes()


def translate_frames_raw(ctx, reads, offsets, item_read, item_frame):
    """Like translate_frames, but returns (uint8 blob, uint64 offsets[n_items + 1]) without building python strings."""
    reads = np.ascontiguousarray(reads, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    item_read = np.ascontiguousarray(item_read, dtype=np.uint64)
    item_frame = np.ascontiguousarray(item_frame, dtype=np.int8)
    n = len(item_read)
    if n == 0:
        return np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.uint64)
    lengths = (offsets[item_read + 1] - offsets[item_read]).astype(np.int64)
    phase = np.abs(item_frame.astype(np.int64)) - 1
    plen = np.maximum(lengths - phase, 0) // 3
    out_offsets = np.zeros(n + 1, dtype=np.uint64)
    out_offsets[1:] = np.cumsum(plen)
    out = np.zeros(max(int(out_offsets[-1]), 1), dtype=np.uint8)
    check(_lib.load().orph_translate_frames(ctx._h, ptr(reads), ptr(offsets), len(offsets) - 1, ptr(item_read),
                                            ptr(item_frame), n, ptr(out_offsets), ptr(out)))
    return out[: int(out_offsets[-1])], out_offsets


def jaccard_column(results):
    """float64 [n, 6] jaccard exactly as the reference computes it (NaN where it reports NaN)."""
    cat = results["category"]
    scored = (cat == _lib.CAT_CODING) | (cat == _lib.CAT_NON_CODING)
    out = np.full(cat.shape, np.nan)
    np.divide(results["n_hits"], results["n_kmers"], out=out, where=scored)
    return out
