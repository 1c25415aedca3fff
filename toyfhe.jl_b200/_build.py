"""Builds libtoyfhe_b200.so in-tree with nvcc for sm_100a (no torch dependency:
the library is a plain C-ABI shared object, see include/toyfhe_b200.h)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libtoyfhe_b200.so")
SOURCES = ["api.cu", "ntt_kernels.cu", "rns_kernels.cu", "rns_fast.cu", "ntt_kernels2.cu", "ntt_kernels3.cu"]
HEADERS = ["engine.h", "modarith.cuh", "ntt_core.cuh", "ntt_core2.cuh", "tables.h", os.path.join("..", "..", "include", "toyfhe_b200.h")]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = [nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-shared", "-Xcompiler", "-fPIC", "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd)
    return LIB
