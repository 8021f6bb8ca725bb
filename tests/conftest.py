import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFDATA = os.path.join(GOLDEN, "reference_data")

SCORE_VECTORS = [
    ("protein", 7), ("dayhoff", 12), ("hp", 31), ("hydrophobic-polar", 21), ("botvinnik", 11), ("protein", 9),
    ("hsdm17", 8),
]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_score_vector(alphabet, ksize, tag=""):
    """tag "long_": the vector of reads beyond 6144 nt (oracle/gen_golden.py long)."""
    with open(os.path.join(GOLDEN, f"ref_scores_{tag}{alphabet}_k{ksize}.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def refdata():
    return REFDATA


@pytest.fixture(scope="session")
def misc_kats():
    with open(os.path.join(GOLDEN, "ref_misc_kats.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def index_kats():
    with open(os.path.join(GOLDEN, "ref_index_kats.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def reads_n22(refdata):
    return os.path.join(refdata, "SRR306838_GSM752691_hsa_br_F_1_trimmed_subsampled_n22.fq")


def read_fastx(path):
    """Minimal FASTA/FASTQ(.gz) reader for the tests (name = full header line)."""
    import gzip

    with open(path, "rb") as f:
        magic = f.read(2)
    opener = gzip.open if magic == b"\x1f\x8b" else open
    with opener(path, "rt") as f:
        lines = [ln.rstrip("\r\n") for ln in f]
    out = []
    if not lines:
        return out
    if lines[0].startswith("@"):
        for i in range(0, len(lines) - 1, 4):
            out.append((lines[i][1:], lines[i + 1]))
    else:
        name, chunks = None, []
        for ln in lines:
            if ln.startswith(">"):
                if name is not None:
                    out.append((name, "".join(chunks)))
                name, chunks = ln[1:], []
            else:
                chunks.append(ln)
        if name is not None:
            out.append((name, "".join(chunks)))
    return out

