"""CPU-only: run the per-thread bodies of the CUDA NTT kernels (ntt_core.cuh /
compiled as host code in tests/emu) thread by thread and compare
with the oracle.  This checks thread mappings, swizzles, twiddle indices, the
natural-order stores and the lazy-reduction bounds without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import toyfhe_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "emu_ntt.cpp")
LIB = os.path.join(HERE, "emu", "libemu.so")
_u64p = C.POINTER(C.c_uint64)


@pytest.fixture(scope="module")
def emu():
    csrc = os.path.join(os.path.dirname(HERE), "toyfhe.jl_b200", "csrc")
    deps = [SRC] + [os.path.join(csrc, f) for f in ("ntt_core.cuh", "ntt_core3.cuh", "modarith.cuh", "tables.h")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", LIB, SRC])
    return C.CDLL(LIB)


def P(a):
    return a.ctypes.data_as(_u64p)


def brev(x, bits):
    return int(format(x, "0%db" % bits)[::-1], 2) if bits else 0


def fwd_global_stages(a, q, psi, s0):
    """levels 1..s0 of the merged CT ladder over the whole row (what ntt_fwd_stage_kernel does)"""
    N = len(a)
    lg = N.bit_length() - 1
    a = [int(v) for v in a]
    for s in range(1, s0 + 1):
        half = N >> s
        for j in range(1 << (s - 1)):
            w = pow(psi, brev((1 << (s - 1)) + j, lg), q)
            base = j * 2 * half
            for k in range(half):
                u, v = a[base + k], a[base + k + half] * w % q
                a[base + k], a[base + k + half] = (u + v) % q, (u - v) % q
    return np.array(a, dtype=np.uint64)


def inv_global_stages(a, q, psi, s0):
    N = len(a)
    lg = N.bit_length() - 1
    ipsi = pow(psi, q - 2, q)
    a = [int(v) for v in a]
    for s in range(s0, 0, -1):
        half = N >> s
        for j in range(1 << (s - 1)):
            w = pow(ipsi, brev((1 << (s - 1)) + j, lg), q)
            base = j * 2 * half
            for k in range(half):
                u, v = a[base + k], a[base + k + half]
                a[base + k], a[base + k + half] = (u + v) % q, (u - v) * w % q
    ninv = pow(N, q - 2, q)
    return np.array([v * ninv % q for v in a], dtype=np.uint64)


PRIMES = {60: None, 40: None}


def ring(N, logq):
    qs, psis = O.prime_chain(N, (logq,))
    return qs[0], psis[0], CO.Rns(N, qs, psis)


@pytest.mark.parametrize("R", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("mode", [0, 1])
def test_v1_row_kernels(emu, R, mode):
    N = 1 << (10 + R)
    for logq in (60, 40):
        q, psi, orc = ring(N, logq)
        rng = np.random.default_rng(R * 10 + mode)
        a = rng.integers(0, q, size=(1, N), dtype=np.uint64)
        a[0, :3] = [q - 1, 0, q - 1]
        want = orc.nntt(a)
        got = np.zeros_like(a)
        assert emu.emu_ntt(R, mode, 0, C.c_uint64(q), C.c_uint64(psi), C.c_uint32(0), P(a), P(got)) == 0
        assert np.array_equal(got, want)
        back = np.zeros_like(a)
        emu.emu_ntt(R, mode, 1, C.c_uint64(q), C.c_uint64(psi), C.c_uint32(0), P(want), P(back))
        assert np.array_equal(back, a)


@pytest.mark.parametrize("R", [0, 2, 4])
def test_v1_row_kernels_mode2_approximate_quotient(emu, R):
    """MODE 2 ladder (shoup_lazy4, T in [0,4q), table reduction of X at fixed levels) on 2^60 + e primes:
    bit-exact against the oracle and no lazy-range violation on random, all-(q-1) and alternating rows."""
    N = 1 << (10 + R)
    emu.emu_overflow_count.restype = C.c_ulonglong
    qs, psis = O.prime_chain(N, (60,) * 17)
    for i in (0, 16):
        q, psi = qs[i], psis[i]
        assert q >> 60 == 1 and q - (1 << 60) < 1 << 28
        orc = CO.Rns(N, [q], [psi])
        rng = np.random.default_rng(R + i)
        rows = [rng.integers(0, q, size=N, dtype=np.uint64), np.full(N, q - 1, dtype=np.uint64),
                np.where(np.arange(N) % 2 == 0, q - 1, 0).astype(np.uint64)]
        for a in rows:
            a = np.ascontiguousarray(a.reshape(1, N))
            got = np.zeros_like(a)
            emu.emu_overflow_count()
            assert emu.emu_ntt(R, 2, 0, C.c_uint64(q), C.c_uint64(psi), C.c_uint32(0), P(a), P(got)) == 0
            assert emu.emu_overflow_count() == 0
            assert np.array_equal(got, orc.nntt(a))


@pytest.mark.parametrize("s0", [0, 1])
def test_v3_forward_kernel(emu, s0):
    """third-generation forward kernel (skewed row buffer, approximate quotient, reductions fused into the
    butterfly adds): bit-exact and inside its lazy ranges on random and extreme rows, first and last chain prime"""
    N = 1 << (14 + s0)
    emu.emu_ntt3_fwd.restype = C.c_longlong
    qs, psis = O.prime_chain(N, (60,) * 17)
    for i in (0, 16):
        q, psi = qs[i], psis[i]
        orc = CO.Rns(N, [q], [psi])
        rng = np.random.default_rng(s0 * 7 + i)
        rows = [rng.integers(0, q, size=N, dtype=np.uint64), np.full(N, q - 1, dtype=np.uint64),
                np.where(np.arange(N) % 2 == 0, q - 1, 0).astype(np.uint64),
                np.where(np.arange(N) < N // 2, q - 1, 1).astype(np.uint64)]
        for a in rows[: 4 if s0 == 0 else 2]:
            a = np.ascontiguousarray(a.reshape(1, N))
            staged = np.ascontiguousarray(fwd_global_stages(a[0], q, psi, s0)).reshape(1, N) if s0 else a
            got = np.zeros_like(a)
            assert emu.emu_ntt3_fwd(4, C.c_uint64(q), C.c_uint64(psi), C.c_uint32(s0), P(staged), P(got)) == 0
            assert np.array_equal(got, orc.nntt(a))
            if s0 == 0:
                emu.emu_ntt3_inv.restype = C.c_longlong
                back = np.zeros_like(a)
                assert emu.emu_ntt3_inv(4, C.c_uint64(q), C.c_uint64(psi), P(got), P(back)) == 0
                assert np.array_equal(back, a)
                # the inverse must also hold its ranges on arbitrary (not transform-image) canonical input
                assert emu.emu_ntt3_inv(4, C.c_uint64(q), C.c_uint64(psi), P(a), P(back)) == 0
                assert np.array_equal(back, orc.inntt(a))
    bad = (1 << 60) + (1 << 28) + 1    # e too large for the approximate-quotient tail: rejected before any arithmetic
    assert emu.emu_ntt3_fwd(4, C.c_uint64(bad), C.c_uint64(3), C.c_uint32(0), P(got), P(got)) == -1


@pytest.mark.parametrize("R", [2, 3, 4])
@pytest.mark.parametrize("logq", [32, 40, 50, 60])
def test_v3_kernels_all_sizes_and_prime_widths(emu, R, logq):
    """the same ladder at N = 2^12, 2^13, 2^14 on primes 2^b + e of every width the reference's chains use
    (60/40-bit CKKS chains of examples/encrypted_mnist, 50-bit primes of test/bfv_crt.jl): forward and inverse
    bit-exact against the oracle, lazy ranges held on random and extreme rows"""
    N = 1 << (10 + R)
    emu.emu_ntt3_fwd.restype = C.c_longlong
    emu.emu_ntt3_inv.restype = C.c_longlong
    qs, psis = O.prime_chain(N, (logq,) * 3)
    for i in (0, 2):
        q, psi = qs[i], psis[i]
        orc = CO.Rns(N, [q], [psi])
        rng = np.random.default_rng(R * 100 + logq + i)
        rows = [rng.integers(0, q, size=N, dtype=np.uint64), np.full(N, q - 1, dtype=np.uint64),
                np.where(np.arange(N) % 2 == 0, q - 1, 0).astype(np.uint64)]
        for a in rows:
            a = np.ascontiguousarray(a.reshape(1, N))
            got, back = np.zeros_like(a), np.zeros_like(a)
            assert emu.emu_ntt3_fwd(R, C.c_uint64(q), C.c_uint64(psi), C.c_uint32(0), P(a), P(got)) == 0
            assert np.array_equal(got, orc.nntt(a))
            assert emu.emu_ntt3_inv(R, C.c_uint64(q), C.c_uint64(psi), P(got), P(back)) == 0
            assert np.array_equal(back, a)
            assert emu.emu_ntt3_inv(R, C.c_uint64(q), C.c_uint64(psi), P(a), P(back)) == 0
            assert np.array_equal(back, orc.inntt(a))


def test_worst_case_inputs_lazy_bounds(emu):
    """all-(q-1) rows maximise every lazy intermediate: the lazy ladder must not wrap 2^64"""
    N = 1 << 14
    q, psi, orc = ring(N, 60)
    a = np.full((1, N), q - 1, dtype=np.uint64)
    want = orc.nntt(a)
    for fn, args in ((emu.emu_ntt, (4, 1, 0)),):
        got = np.zeros_like(a)
        fn(*args, C.c_uint64(q), C.c_uint64(psi), C.c_uint32(0), P(a), P(got))
        assert np.array_equal(got, want)


@pytest.mark.parametrize("s0", [1, 2])
def test_long_rows_as_sub_blocks(emu, s0):
    N = 1 << (14 + s0)
    q, psi, orc = ring(N, 60)
    rng = np.random.default_rng(s0)
    a = rng.integers(0, q, size=(1, N), dtype=np.uint64)
    want = orc.nntt(a)
    staged = np.ascontiguousarray(fwd_global_stages(a[0], q, psi, s0)).reshape(1, N)
    got = np.zeros_like(a)
    emu.emu_ntt(4, 1, 0, C.c_uint64(q), C.c_uint64(psi), C.c_uint32(s0), P(staged), P(got))
    assert np.array_equal(got, want)
    part = np.zeros_like(a)
    emu.emu_ntt(4, 1, 1, C.c_uint64(q), C.c_uint64(psi), C.c_uint32(s0), P(want), P(part))
    assert np.array_equal(inv_global_stages(part[0], q, psi, s0), a[0])


def test_shared_memory_layouts_are_conflict_free(emu):
    assert emu.emu_bank_conflicts(4) == 1      # 512x32 kernel at N = 2^14
    for R in (2, 3, 4):                        # skewed layout of the third-generation kernels
        assert emu.emu_bank_conflicts3(R) == 1
