"""Builds the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "liborpheum_b200.so")
SOURCES = ["orpheum_b200.cu", "fastx.cpp", "hostpack.cpp", "fused_ingest.cpp", "ingest.cpp", "summary.cpp"]
HEADERS = ["device_common.cuh", "score_kernels.cuh", "task_kernels.cuh", "bloom_kernels.cuh", "long_kernels.cuh", "ingest_pipeline.inl", "synth_kernels.cuh", "hostpack.h", "fused_ingest.h", "ingest_internal.h", "stream.inl", "../../include/orpheum_b200.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-pthread,-Wall",
    "-shared",
]


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; the CUDA library cannot be built")
    return exe


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.exists(d) and os.path.getmtime(d) > built for d in deps)


def build(force=False, verbose=False, extra_flags=(), out=None):
    """Compile csrc/*.cu into csrc/liborpheum_b200.so (or `out`: a development variant built with extra -D flags)."""
    if out is None and not force and not is_stale():
        return LIB_PATH
    cmd = [nvcc(), *NVCC_FLAGS, *extra_flags, "-o", out or LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES] + ["-lz"]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return out or LIB_PATH


if __name__ == "__main__":
    import sys

    build(force=True, verbose=True, extra_flags=tuple(sys.argv[1:]))
