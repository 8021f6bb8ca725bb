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
    assert lib.orph_abi_version() == 1
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
