"""CPU-side checks of the drop-in boundary: the C-ABI library builds/loads and exports exactly the
symbols include/orpheum_b200.h declares (no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

from orpheum_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def declared_functions():
    text = open(os.path.join(ROOT, "include", "orpheum_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(orph_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(lib):
    names = declared_functions()
    assert len(names) >= 25
    assert sorted(_lib.PROTOTYPES) == names
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"


def test_abi_version_and_struct_size(lib):
    assert lib.orph_abi_version() == 2
    assert _lib.READ_RESULT_DTYPE.itemsize == 32
    assert _lib.READ_RESULT_DTYPE.fields["category"][1] == 24
    assert _lib.READ_RESULT_DTYPE.fields["best_frame"][1] == 30


def test_table_sizes_match_khmer(lib, misc_kats):
    from orpheum_b200 import gpu

    for x, primes in misc_kats["table_sizes"].items():
        assert gpu.table_sizes(int(x), 4) == primes
    with pytest.raises(ValueError):
        gpu.table_sizes(2, 4)


def test_host_packer(lib):
    from orpheum_b200 import gpu

    rng = np.random.default_rng(0)
    seqs = ["".join(rng.choice(list("ACGT"), size=int(n))) for n in rng.integers(1, 200, 300)]
    seqs[5] = seqs[5][:3] + "N" + seqs[5][4:]
    seqs[17] = seqs[17].lower()
    seqs[299] = seqs[299] + "R"
    flat, offsets = gpu.concat_reads(seqs)
    for threads in (1, 4):
        packed, needs = gpu.pack_reads(flat, offsets, n_threads=threads)
        assert sorted(np.nonzero(needs)[0].tolist()) == [5, 17, 299]
        code = {"A": 0, "C": 1, "T": 2, "G": 3}
        for r in (0, 1, 100, 298):
            for i, ch in enumerate(seqs[r]):
                g = int(offsets[r]) + i
                assert (int(packed[g // 32]) >> (2 * (g % 32))) & 3 == code[ch]


def test_no_gpu_fails_loudly(lib):
    """The product path has no CPU fallback: without a device, creating a context raises."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from orpheum_b200 import gpu

    with pytest.raises(Exception):
        gpu.Context(0)


def test_host_packer_simd_against_numpy(lib):
    """The SIMD packer (SSSE3 / AVX2, thread pool) against a numpy restatement: every byte value, reads that start and
    end inside 32-base words, a non-zero first offset, a partial last word, one and many threads."""
    from orpheum_b200 import gpu

    rng = np.random.default_rng(7)
    n = 40_000
    lengths = rng.integers(1, 400, n)
    offsets = np.zeros(n + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lengths)
    offsets += np.uint64(13)  # the stream does not start at base 0 of the buffer
    total = int(offsets[-1])
    flat = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, total)].copy()
    dirty = rng.choice(n, 400, replace=False)
    for r in dirty:  # arbitrary bytes (lower case, N, IUPAC, control characters, >= 0x80) somewhere in the read
        a, b = int(offsets[r]), int(offsets[r + 1])
        pos = rng.integers(a, b, rng.integers(1, 4))
        flat[pos] = rng.integers(0, 256, len(pos), dtype=np.uint8)
    valid = np.zeros(256, dtype=bool)
    valid[np.frombuffer(b"ACGT", dtype=np.uint8)] = True
    body = flat[13:]
    codes = ((body >> 1) & 3).astype(np.uint64)
    pad = (-len(codes)) % 32
    words = np.concatenate([codes, np.zeros(pad, dtype=np.uint64)]).reshape(-1, 32)
    want = np.zeros(len(words), dtype=np.uint64)
    for i in range(32):
        want |= words[:, i] << np.uint64(2 * i)
    bad_pos = np.nonzero(~valid[body])[0] + 13
    want_needs = np.zeros(n, dtype=np.uint8)
    want_needs[np.searchsorted(offsets, bad_pos, side="right") - 1] = 1
    for threads in (1, 3, 8):
        packed, needs = gpu.pack_reads(flat, offsets, n_threads=threads)
        np.testing.assert_array_equal(packed[: len(want)], want)
        np.testing.assert_array_equal(needs, want_needs)
