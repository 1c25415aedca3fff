"""Builds libtoyfhe_b200.so in-tree with nvcc for sm_100a (no torch dependency:
the library is a plain C-ABI shared object, see include/toyfhe_b200.h)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libtoyfhe_b200.so")
SOURCES = ["api.cu", "ntt_kernels.cu", "rns_kernels.cu", "rns_fast.cu", "ntt_kernels3.cu", "ntt_kernels4.cu", "sample_kernels.cu", "ckks_kernels.cu"]
HEADERS = ["engine.h", "modarith.cuh", "ntt_core.cuh", "ntt_core3.cuh", "ntt_v3_kernels.cuh", "tables.h", os.path.join("..", "..", "include", "toyfhe_b200.h")]


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
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    flags = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC"]
    if verbose:
        flags.append("-Xptxas=-v")
    # one nvcc process per translation unit, in parallel (the unrolled kernels take ~1 min of ptxas in total)
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        procs.append((src, obj, subprocess.Popen([nvcc] + flags + ["-c", os.path.join(CSRC, src), "-o", obj],
                                                 stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs, failed = [], []
    for src, obj, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            failed.append((src, out))
        objs.append(obj)
    if failed:
        raise RuntimeError("nvcc failed:\n" + "\n".join(f"--- {s}\n{o}" for s, o in failed))
    subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs)
    return LIB
