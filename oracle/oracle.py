"""ctypes front-end of the CPU oracle (oracle/orpheum_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of orpheum_oracle.c.  Nothing under
``orpheum_b200/`` imports this module; only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu_baseline / ``--impl reference`` legs do.

The alphabet tables here are an independent restatement of
``orpheum/sequence_encodings.py:32-58`` (dayhoff), ``:95-118`` (hp) and the other
reduced alphabets (``:62-277``); ``encode_lut`` follows ``encode_peptide``
(``sequence_encodings.py:457-468``): identity for protein-like names, unmapped bytes
pass through unchanged, unknown alphabet -> ValueError.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liborpheum_oracle.so")

CAT_STOP, CAT_SHORT_PEPTIDE, CAT_LOWCPLX_PEPTIDE, CAT_CODING, CAT_NONCODING, CAT_LOWCPLX_NT = range(6)
FRAMES = (1, 2, 3, -1, -2, -3)

# group string -> letter, per alphabet (sequence_encodings.py)
_GROUPS = {
    "dayhoff": {"C": "a", "AGPST": "b", "DENQ": "c", "HKR": "d", "ILMV": "e", "FWY": "f"},
    "dayhoff_v2": {"C": "a", "AGP": "b", "ST": "B", "DENQ": "c", "HKR": "d", "ILMV": "e", "FWY": "f"},
    "hp": {"AFGILMPVWY": "h", "CDEHKNQRST": "p"},
    "botvinnik": {"AG": "a", "LIV": "b", "FY": "c", "ST": "d", "NQ": "e", "DE": "f", "RK": "g", "C": "h", "M": "i",
                  "W": "j", "H": "k", "P": "m"},
    "aa9": {"AST": "a", "CFILMVY": "b", "DN": "c", "EQ": "d", "G": "e", "H": "f", "KR": "g", "P": "h", "W": "i"},
    "gbmr4": {"ADKERNTSQ": "a", "YFLIVMCWH": "b", "G": "c", "P": "d"},
    "sdm12": {"A": "a", "D": "b", "KER": "c", "N": "d", "TSQ": "e", "YF": "f", "LIVM": "g", "C": "h", "W": "i",
              "H": "j", "G": "k", "P": "l"},
    "hsdm17": {"A": "a", "D": "b", "KE": "c", "R": "d", "N": "e", "T": "f", "S": "g", "Q": "h", "Y": "i", "F": "j",
               "LIV": "k", "M": "l", "C": "m", "G": "n", "P": "o", "W": "p", "H": "q"},
}
_ENCODED = {
    "hp": "hp", "hp2": "hp", "hydrophobic-polar": "hp",
    "dayhoff": "dayhoff", "dayhoff6": "dayhoff", "dayhoff_v2": "dayhoff_v2",
    "botvinnik": "botvinnik", "botvinnik8": "botvinnik",
    "aa9": "aa9", "gbmr4": "gbmr4", "sdm12": "sdm12", "hsdm17": "hsdm17",
}
_IDENTITY = ("protein", "peptide", "protein20", "peptide20", "aa20")


def encode_lut(alphabet):
    """256-byte table equivalent to ``encode_peptide(seq, alphabet)``."""
    lut = np.arange(256, dtype=np.uint8)
    if alphabet in _ENCODED:
        for group, letter in _GROUPS[_ENCODED[alphabet]].items():
            for aa in group:
                lut[ord(aa)] = ord(letter)
    elif alphabet not in _IDENTITY:
        raise ValueError(f"{alphabet} is not a valid amino acid encoding")
    return lut


def build():
    """Compile the oracle library in place (gcc only)."""
    subprocess.run(["make", "-C", _HERE, "liborpheum_oracle.so"], check=True, stdout=subprocess.DEVNULL)


class FrameScore(ctypes.Structure):
    _fields_ = [
        ("jaccard", ctypes.c_double),
        ("n_kmers", ctypes.c_int64),
        ("n_hits", ctypes.c_int64),
        ("category", ctypes.c_int32),
        ("frame", ctypes.c_int32),
        ("probes_min", ctypes.c_uint64),
        ("probes_max", ctypes.c_uint64),
    ]


FRAME_SCORE_DTYPE = np.dtype(
    [("jaccard", "<f8"), ("n_kmers", "<i8"), ("n_hits", "<i8"), ("category", "<i4"), ("frame", "<i4"),
     ("probes_min", "<u8"), ("probes_max", "<u8")]
)

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    src = os.path.join(_HERE, "orpheum_oracle.c")
    if not os.path.exists(_LIB_PATH) or (
        os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_LIB_PATH)
    ):
        build()
    L = ctypes.CDLL(_LIB_PATH)
    u8p, u64p, vp = ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_uint64), ctypes.c_void_p
    L.orc_hash_murmur.restype = ctypes.c_uint64
    L.orc_hash_murmur.argtypes = [ctypes.c_char_p, ctypes.c_uint64]
    L.orc_table_sizes.restype = ctypes.c_int
    L.orc_table_sizes.argtypes = [ctypes.c_uint64, ctypes.c_uint32, u64p]
    L.orc_ng_new.restype = vp
    L.orc_ng_new.argtypes = [ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint32]
    L.orc_ng_new_sizes.restype = vp
    L.orc_ng_new_sizes.argtypes = [ctypes.c_uint32, ctypes.c_uint32, u64p]
    L.orc_ng_free.restype = None
    L.orc_ng_free.argtypes = [vp]
    L.orc_ng_load.restype = vp
    L.orc_ng_load.argtypes = [ctypes.c_char_p]
    L.orc_ng_save.restype = ctypes.c_int
    L.orc_ng_save.argtypes = [vp, ctypes.c_char_p]
    for name in ("orc_ng_ksize", "orc_ng_n_tables"):
        getattr(L, name).restype = ctypes.c_uint32
        getattr(L, name).argtypes = [vp]
    L.orc_ng_table_size.restype = ctypes.c_uint64
    L.orc_ng_table_size.argtypes = [vp, ctypes.c_uint32]
    L.orc_ng_table.restype = u8p
    L.orc_ng_table.argtypes = [vp, ctypes.c_uint32]
    L.orc_ng_n_unique_kmers.restype = ctypes.c_uint64
    L.orc_ng_n_unique_kmers.argtypes = [vp]
    L.orc_ng_popcount.restype = ctypes.c_uint64
    L.orc_ng_popcount.argtypes = [vp, ctypes.c_uint32]
    L.orc_ng_add.restype = ctypes.c_int
    L.orc_ng_add.argtypes = [vp, ctypes.c_uint64]
    L.orc_ng_get.restype = ctypes.c_int
    L.orc_ng_get.argtypes = [vp, ctypes.c_uint64]
    L.orc_ng_add_peptide.restype = ctypes.c_uint64
    L.orc_ng_add_peptide.argtypes = [vp, ctypes.c_char_p, ctypes.c_uint64, vp]
    L.orc_ng_add_peptides.restype = None
    L.orc_ng_add_peptides.argtypes = [vp, vp, vp, ctypes.c_uint64, vp]
    L.orc_count_distinct_kmers.restype = ctypes.c_uint64
    L.orc_count_distinct_kmers.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_uint32]
    L.orc_translate_codon.restype = ctypes.c_uint8
    L.orc_translate_codon.argtypes = [ctypes.c_char_p]
    L.orc_reverse_complement.restype = None
    L.orc_reverse_complement.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_char_p]
    L.orc_six_frame_translation.restype = None
    L.orc_six_frame_translation.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_char_p, u64p]
    L.orc_fastp_ndiff.restype = ctypes.c_uint64
    L.orc_fastp_ndiff.argtypes = [ctypes.c_char_p, ctypes.c_uint64]
    L.orc_score_read.restype = ctypes.c_int
    L.orc_score_read.argtypes = [vp, ctypes.c_char_p, ctypes.c_uint64, vp, ctypes.c_double, vp]
    L.orc_score_reads.restype = ctypes.c_int64
    L.orc_score_reads.argtypes = [vp, vp, vp, ctypes.c_uint64, vp, ctypes.c_double, vp, ctypes.c_int]
    L.orc_max_threads.restype = ctypes.c_int
    L.orc_max_threads.argtypes = []
    _lib = L
    return L


def _b(s):
    return s.encode("latin-1") if isinstance(s, str) else bytes(s)


def hash_murmur(kmer):
    """sourmash.minhash.hash_murmur(kmer) (seed 42)."""
    b = _b(kmer)
    return int(lib().orc_hash_murmur(b, len(b)))


def table_sizes(tablesize, n_tables):
    out = (ctypes.c_uint64 * n_tables)()
    if lib().orc_table_sizes(int(tablesize), n_tables, out) != 0:
        raise ValueError("could not find table sizes")
    return [int(x) for x in out]


def translate_codon(codon):
    b = _b(codon)
    if len(b) != 3:
        return ""
    return chr(lib().orc_translate_codon(b))


def reverse_complement(seq):
    b = _b(seq)
    out = ctypes.create_string_buffer(len(b) + 1)
    lib().orc_reverse_complement(b, len(b), out)
    return out.raw[: len(b)].decode("latin-1")


def six_frame_translation(seq):
    """dict {1,2,3,-1,-2,-3: peptide} like TranslateSingleSeq.six_frame_translation()."""
    b = _b(seq)
    stride = len(b) // 3 + 1
    out = ctypes.create_string_buffer(6 * stride)
    lens = (ctypes.c_uint64 * 6)()
    lib().orc_six_frame_translation(b, len(b), out, lens)
    return {fr: out.raw[i * stride : i * stride + lens[i]].decode("latin-1") for i, fr in enumerate(FRAMES)}


def fastp_complexity(seq):
    b = _b(seq)
    return lib().orc_fastp_ndiff(b, len(b)) / len(b)


def count_distinct_kmers(seq, k):
    b = _b(seq)
    return int(lib().orc_count_distinct_kmers(b, len(b), k))


class Nodegraph:
    """khmer.Nodegraph look-alike backed by the C oracle."""

    def __init__(self, ksize=None, tablesize=None, n_tables=4, _handle=None):
        self._L = lib()
        if _handle is not None:
            self._h = _handle
        else:
            self._h = self._L.orc_ng_new(int(ksize), int(tablesize), int(n_tables))
            if not self._h:
                raise ValueError("bad Nodegraph parameters")

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_ng_free(self._h)
            self._h = None

    @classmethod
    def load(cls, path):
        h = lib().orc_ng_load(os.fsencode(path))
        if not h:
            raise OSError(f"{path} is not a readable OXLI nodegraph file")
        return cls(_handle=h)

    def save(self, path):
        if self._L.orc_ng_save(self._h, os.fsencode(path)) != 0:
            raise OSError(f"cannot write {path}")

    def ksize(self):
        return int(self._L.orc_ng_ksize(self._h))

    def n_tables(self):
        return int(self._L.orc_ng_n_tables(self._h))

    def hashsizes(self):
        return [int(self._L.orc_ng_table_size(self._h, i)) for i in range(self.n_tables())]

    def n_unique_kmers(self):
        return int(self._L.orc_ng_n_unique_kmers(self._h))

    def n_occupied(self, table=0):
        return int(self._L.orc_ng_popcount(self._h, table))

    def add(self, h):
        return bool(self._L.orc_ng_add(self._h, int(h) & 0xFFFFFFFFFFFFFFFF))

    def get(self, h):
        return int(self._L.orc_ng_get(self._h, int(h) & 0xFFFFFFFFFFFFFFFF))

    def table(self, i):
        """numpy view (copy) of table i's bytes, exactly as stored in the OXLI file."""
        n = self.hashsizes()[i] // 8 + 1
        p = self._L.orc_ng_table(self._h, i)
        return np.ctypeslib.as_array(p, shape=(n,)).copy()

    def add_peptide(self, seq, lut):
        b = _b(seq)
        lut = np.ascontiguousarray(lut, dtype=np.uint8)
        return int(self._L.orc_ng_add_peptide(self._h, b, len(b), lut.ctypes.data))

    def add_peptides(self, residues, offsets, lut):
        residues = np.ascontiguousarray(residues, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        lut = np.ascontiguousarray(lut, dtype=np.uint8)
        self._L.orc_ng_add_peptides(self._h, residues.ctypes.data, offsets.ctypes.data, len(offsets) - 1, lut.ctypes.data)


def score_read(graph, seq, lut, jaccard_threshold):
    """Six per-frame results for one read as a numpy structured array (FRAME_SCORE_DTYPE).
    Raises ZeroDivisionError for an empty read like translate.py:83."""
    b = _b(seq)
    lut = np.ascontiguousarray(lut, dtype=np.uint8)
    out = np.zeros(6, dtype=FRAME_SCORE_DTYPE)
    rc = lib().orc_score_read(graph._h, b, len(b), lut.ctypes.data, float(jaccard_threshold), out.ctypes.data)
    if rc != 0:
        raise ZeroDivisionError("division by zero")
    return out


def score_reads(graph, reads, offsets, lut, jaccard_threshold, n_threads=1):
    """Batch scoring over concatenated reads -> array [n_reads, 6] of FRAME_SCORE_DTYPE."""
    reads = np.ascontiguousarray(reads, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    lut = np.ascontiguousarray(lut, dtype=np.uint8)
    n = len(offsets) - 1
    out = np.zeros((n, 6), dtype=FRAME_SCORE_DTYPE)
    n_empty = lib().orc_score_reads(
        graph._h, reads.ctypes.data, offsets.ctypes.data, n, lut.ctypes.data, float(jaccard_threshold),
        out.ctypes.data, int(n_threads),
    )
    return out, int(n_empty)


def max_threads():
    return int(lib().orc_max_threads())
