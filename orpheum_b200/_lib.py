"""ctypes binding of ``csrc/liborpheum_b200.so`` (the C ABI declared in include/orpheum_b200.h).

There is no CPU fallback: if the library has not been built, or no B200 is present, the
calls raise.  ``load()`` only dlopens the library (works without a GPU, which is what the
CPU-side symbol test uses); anything that touches the device goes through ``orph_ctx_create``.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ORPHEUM_B200_LIB") or os.path.join(_HERE, "csrc", "liborpheum_b200.so")  # (override: kernel A/B builds)

ORPH_OK = 0
ORPH_E_INVALID, ORPH_E_CUDA, ORPH_E_NOMEM, ORPH_E_TOO_LONG, ORPH_E_EMPTY_READ, ORPH_E_IO = -1, -2, -3, -4, -5, -6
READS_ASCII, READS_PACKED2 = 0, 1
MAX_TABLES = 16
MAX_READ_LEN = 196607

CAT_STOP_CODONS, CAT_TOO_SHORT_PEPTIDE, CAT_LOW_COMPLEXITY_PEPTIDE, CAT_CODING, CAT_NON_CODING, \
    CAT_LOW_COMPLEXITY_NUCLEOTIDE = range(6)
CAT_INVALID = 255
FLAG_OPEN_MASK, FLAG_LOW_COMPLEXITY, FLAG_BYTE_PATH = 0x3F, 0x40, 0x80

FRAMES = (1, 2, 3, -1, -2, -3)

# struct orph_read_result (32 bytes)
READ_RESULT_DTYPE = np.dtype(
    [("n_kmers", "<u2", (6,)), ("n_hits", "<u2", (6,)), ("category", "u1", (6,)), ("best_frame", "i1"), ("flags", "u1")]
)
assert READ_RESULT_DTYPE.itemsize == 32



class IngestBatch(ctypes.Structure):
    """struct orph_ingest_batch"""
    _fields_ = [("n_reads", ctypes.c_uint64), ("packed", ctypes.c_void_p), ("offsets32", ctypes.c_void_p),
                ("total_bases", ctypes.c_uint64), ("max_read_len", ctypes.c_uint32), ("reserved", ctypes.c_uint32),
                ("names", ctypes.c_void_p), ("name_offsets", ctypes.c_void_p), ("text", ctypes.c_void_p),
                ("seq_start", ctypes.c_void_p), ("seq_len", ctypes.c_void_p), ("flagged", ctypes.c_void_p),
                ("n_flagged", ctypes.c_uint64)]


class StreamBatch(ctypes.Structure):
    """struct orph_stream_batch"""
    _fields_ = [("n_reads", ctypes.c_uint64), ("results", ctypes.c_void_p), ("names", ctypes.c_void_p),
                ("name_offsets", ctypes.c_void_p), ("text", ctypes.c_void_p), ("seq_start", ctypes.c_void_p),
                ("seq_len", ctypes.c_void_p), ("n_byte_path", ctypes.c_uint64), ("status", ctypes.c_int)]


# every symbol include/orpheum_b200.h declares: name -> (restype, argtypes)
_vp, _u8p, _u64p = ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p
_pp = ctypes.POINTER(ctypes.c_void_p)
PROTOTYPES = {
    "orph_device_count": (ctypes.c_int, []),
    "orph_ctx_create": (ctypes.c_int, [ctypes.c_int, _vp, _pp]),
    "orph_ctx_destroy": (None, [_vp]),
    "orph_ctx_synchronize": (ctypes.c_int, [_vp]),
    "orph_ctx_check": (ctypes.c_int, [_vp]),
    "orph_ctx_set_force_exact_dedup": (ctypes.c_int, [_vp, ctypes.c_int]),
    "orph_ctx_set_thread_path": (ctypes.c_int, [_vp, ctypes.c_int]),
    "orph_ctx_launch_count": (ctypes.c_uint64, [_vp]),
    "orph_ctx_set_kernel_timing": (ctypes.c_int, [_vp, ctypes.c_int]),
    "orph_ctx_kernel_times": (ctypes.c_int, [_vp, _vp, _u64p]),
    "orph_ctx_set_host_threads": (ctypes.c_int, [_vp, ctypes.c_int]),
    "orph_ctx_ingest_stats": (ctypes.c_int, [_vp, _u64p, _u64p, _vp, _vp, _vp]),
    "orph_last_error": (ctypes.c_char_p, []),
    "orph_abi_version": (ctypes.c_int, []),
    "orph_table_sizes": (ctypes.c_int, [ctypes.c_uint64, ctypes.c_uint32, _u64p]),
    "orph_filter_create": (ctypes.c_int, [_vp, ctypes.c_uint32, ctypes.c_uint32, _u64p, _pp]),
    "orph_filter_from_host": (ctypes.c_int, [_vp, ctypes.c_uint32, ctypes.c_uint32, _u64p, _vp, _pp]),
    "orph_filter_to_host": (ctypes.c_int, [_vp, _vp, _u64p]),
    "orph_filter_destroy": (None, [_vp]),
    "orph_filter_info": (ctypes.c_int, [_vp, _vp, _vp, _u64p]),
    "orph_filter_n_unique_kmers": (ctypes.c_uint64, [_vp]),
    "orph_filter_popcount": (ctypes.c_int, [_vp, ctypes.c_uint32, _u64p]),
    "orph_filter_device_buffer": (ctypes.c_int, [_vp, _pp, _u64p, _u64p]),
    "orph_filter_replicate": (ctypes.c_int, [_vp, _vp, _pp]),
    "orph_filter_add_peptides": (ctypes.c_int, [_vp, _u8p, _u64p, ctypes.c_uint64, _u8p, _u64p]),
    "orph_filter_add_hashes": (ctypes.c_int, [_vp, _u64p, ctypes.c_uint64, _u8p]),
    "orph_filter_query_hashes": (ctypes.c_int, [_vp, _u64p, ctypes.c_uint64, _u8p]),
    "orph_hash_kmers": (ctypes.c_int, [_vp, _u8p, ctypes.c_uint32, ctypes.c_uint64, _u64p]),
    "orph_score_reads": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int, _u64p, ctypes.c_uint64, _u8p, ctypes.c_double, _vp]),
    "orph_score_reads_device": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int, _u64p, ctypes.c_uint64, ctypes.c_uint32, _u8p,
                                               ctypes.c_double, _vp]),
    "orph_pack_reads": (ctypes.c_int, [_u8p, _u64p, ctypes.c_uint64, _u64p, _u8p, ctypes.c_int]),
    "orph_translate_frames": (ctypes.c_int, [_vp, _u8p, _u64p, ctypes.c_uint64, _u64p, _vp, ctypes.c_uint64, _u64p, _u8p]),
    "orph_jaccard": (ctypes.c_double, [_vp, ctypes.c_int]),
    "orph_fastx_open": (ctypes.c_int, [ctypes.c_char_p, _pp]),
    "orph_fastx_next_batch": (ctypes.c_int, [_vp, ctypes.c_uint64, ctypes.c_uint64, _u64p, _pp, _pp, _pp, _pp]),
    "orph_fastx_copy_batch": (ctypes.c_int, [_vp, _u8p, _u64p]),
    "orph_fastx_close": (None, [_vp]),
    "orph_ingest_open": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int, _pp]),
    "orph_ingest_next": (ctypes.c_int, [_vp, ctypes.c_uint64, ctypes.POINTER(IngestBatch)]),
    "orph_ingest_close": (None, [_vp]),
    "orph_gather_reads": (ctypes.c_int, [_vp, _vp, _vp, _u64p, ctypes.c_uint64, _u8p, _u64p]),
    "orph_stream_open": (ctypes.c_int, [_vp, _vp, ctypes.c_char_p, _u8p, ctypes.c_double, ctypes.c_uint64, ctypes.c_int, _pp]),
    "orph_score_reads_device_n": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _u64p, ctypes.c_uint64, ctypes.c_uint32, _u8p, ctypes.c_double, _vp]),
    "orph_pack_reads_device": (ctypes.c_int, [_vp, _vp, _u64p, ctypes.c_uint64, ctypes.c_uint64, _vp, _vp, _vp, _vp]),
    "orph_stream_emit": (ctypes.c_int, [_vp, _vp, ctypes.c_uint32, ctypes.POINTER(ctypes.c_void_p * 4), ctypes.POINTER(ctypes.c_uint64 * 4)]),
    "orph_stream_open_multi": (ctypes.c_int, [_pp, _pp, ctypes.c_int, ctypes.c_char_p, _u8p, ctypes.c_double, ctypes.c_uint64, ctypes.c_int, _pp]),
    "orph_stream_next": (ctypes.c_int, [_vp, ctypes.POINTER(StreamBatch)]),
    "orph_stream_close": (None, [_vp]),
    "orph_csv_write_scores": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int, _u8p, _u64p, ctypes.c_uint64, _vp, _vp, ctypes.c_char_p, ctypes.c_int]),
    "orph_hash_names": (ctypes.c_int, [_u8p, _u64p, ctypes.c_uint64, _u64p, ctypes.c_int]),
    "orph_u64_all_distinct": (ctypes.c_int, [_u64p, ctypes.c_uint64, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]),
    "orph_summary_partial": (ctypes.c_int, [_vp, ctypes.c_uint64, _vp, _u8p, _u64p, _u64p, _u64p, ctypes.c_int]),
    "orph_f64_pairwise_sum": (ctypes.c_int, [_vp, _u8p, ctypes.c_uint64, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]),
    "orph_format_fasta": (ctypes.c_int, [ctypes.c_int, _u8p, _u64p, ctypes.c_uint64, _vp, _u8p, _u64p, _u8p, _u64p, ctypes.c_int, _u8p,
                                         ctypes.c_uint64, _u64p]),
    "orph_format_float_repr": (ctypes.c_int, [ctypes.c_double, ctypes.c_char_p]),
    "orph_synth_reads_device": (ctypes.c_int, [_vp, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32, _vp, _vp,
                                               ctypes.c_uint64, _vp, _vp]),
    "orph_bench_gather_peak": (ctypes.c_int, [_vp, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int,
                                              ctypes.POINTER(ctypes.c_double)]),
}


class OrpheumB200Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"[orph_status {code}] {message}")
        self.code = code


_lib = None


def load():
    """dlopen the in-tree CUDA library.  Raises (no fallback) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OrpheumB200Error(
            ORPH_E_IO,
            f"{LIB_PATH} is missing: build it with `python -m orpheum_b200.build` (or __graft_entry__.build()); "
            "there is no CPU fallback for the scoring path",
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error():
    return load().orph_last_error().decode("utf-8", "replace")


def check(rc):
    """Map an orph_status to the exception the reference raises in the same situation."""
    if rc == ORPH_OK:
        return
    msg = last_error()
    if rc == ORPH_E_EMPTY_READ:
        raise ZeroDivisionError("division by zero")  # orpheum/translate.py:83 on an empty read
    if rc == ORPH_E_NOMEM:
        raise MemoryError(msg)
    if rc == ORPH_E_INVALID or rc == ORPH_E_TOO_LONG:
        raise ValueError(msg)
    raise OrpheumB200Error(rc, msg)


def ptr(a):
    """Address of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    return a.ctypes.data
