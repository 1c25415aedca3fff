"""ctypes bindings for oracle/oracle.c (TEST INFRASTRUCTURE ONLY -- see the
header of oracle.c).  Arrays are numpy uint64, C-contiguous, residue-major
``[..][L][N]`` exactly like the engine's C-ABI buffers."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "build", "liboracle.so")
_lib = None

_u64p = C.POINTER(C.c_uint64)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "build/liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_rns_create.restype = C.c_void_p
        _lib.orc_rns_create.argtypes = [C.c_uint64, C.c_int, _u64p, _u64p]
        _lib.orc_rns_destroy.argtypes = [C.c_void_p]
        _lib.orc_num_threads.restype = C.c_int
        _lib.orc_ndigits.restype = C.c_int
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u64p)


def _arr(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.uint64))


def num_threads() -> int:
    return lib().orc_num_threads()


def set_threads(n: int) -> None:
    lib().orc_set_threads(C.c_int(n))


class Rns:
    """One NegacyclicRing over an RNS basis (crt.jl:282-295 product)."""

    def __init__(self, N: int, qs, psis):
        self.N, self.L = int(N), len(qs)
        self.qs, self.psis = [int(q) for q in qs], [int(p) for p in psis]
        self._q, self._psi = _arr(self.qs), _arr(self.psis)
        self.h = C.c_void_p(lib().orc_rns_create(C.c_uint64(self.N), C.c_int(self.L), _p(self._q), _p(self._psi)))

    def __del__(self):
        try:
            lib().orc_rns_destroy(self.h)
        except Exception:
            pass

    def _rows(self, a):
        assert a.shape[-1] == self.N and a.size % (self.N * self.L) == 0, a.shape
        return a.size // self.N

    def nntt(self, a):
        a = _arr(a); out = np.empty_like(a)
        lib().orc_rns_nntt(self.h, _p(a), _p(out), C.c_long(self._rows(a)))
        return out

    def inntt(self, a):
        a = _arr(a); out = np.empty_like(a)
        lib().orc_rns_inntt(self.h, _p(a), _p(out), C.c_long(self._rows(a)))
        return out

    def _binop(self, op, a, b):
        a, b = _arr(a), _arr(b); out = np.empty_like(a)
        assert a.shape == b.shape
        lib().orc_rns_binop(self.h, C.c_int(op), _p(a), _p(b), _p(out), C.c_long(self._rows(a)))
        return out

    def add(self, a, b): return self._binop(0, a, b)
    def sub(self, a, b): return self._binop(1, a, b)
    def mul(self, a, b): return self._binop(2, a, b)

    def neg(self, a):
        a = _arr(a); out = np.empty_like(a)
        lib().orc_rns_neg(self.h, _p(a), _p(out), C.c_long(self._rows(a)))
        return out

    def scalar_mul(self, a, s: int):
        a = _arr(a); out = np.empty_like(a)
        sr = _arr([int(s) % q for q in self.qs])
        lib().orc_rns_scalar_mul(self.h, _p(a), _p(sr), _p(out), C.c_long(self._rows(a)))
        return out

    def ring_mul(self, a, b):
        a, b = _arr(a), _arr(b); out = np.empty_like(a)
        lib().orc_rns_ring_mul(self.h, _p(a), _p(b), _p(out), C.c_long(self._rows(a)))
        return out

    def galois(self, a, g: int):
        a = _arr(a); out = np.empty_like(a)
        lib().orc_rns_galois(self.h, C.c_uint64(g), _p(a), _p(out), C.c_long(self._rows(a)))
        return out

    def ct_tensor(self, c1, c2):
        c1, c2 = _arr(c1), _arr(c2)
        B = c1.size // (2 * self.L * self.N)
        out = np.empty(c1.shape[:-3] + (3, self.L, self.N), dtype=np.uint64)
        lib().orc_ct_tensor(self.h, _p(c1), _p(c2), _p(out), C.c_long(B))
        return out

    def modswitch(self, a):
        a = _arr(a)
        polys = a.size // (self.L * self.N)
        out = np.empty(a.shape[:-2] + (self.L - 1, self.N), dtype=np.uint64)
        lib().orc_modswitch(self.h, _p(a), _p(out), C.c_long(polys))
        return out

    def crt_expand(self, a, P: int):
        a = _arr(a)
        polys = a.size // (self.L * self.N)
        out = np.empty(a.shape[:-2] + (self.L + 1, self.N), dtype=np.uint64)
        lib().orc_crt_expand(self.h, C.c_uint64(P), _p(a), _p(out), C.c_long(polys))
        return out

    def keyswitch_digits(self, cend, w: int, target_qs=None):
        cend = _arr(cend)
        tq = _arr(self.qs if target_qs is None else target_qs)
        D = self.L if w == 0 else ndigits(self.qs, w)
        out = np.empty((D, len(tq), self.N), dtype=np.uint64)
        lib().orc_keyswitch_digits(C.c_uint64(self.N), C.c_int(self.L), _p(self._q), C.c_int(len(tq)), _p(tq),
                                   C.c_int(w), _p(cend), _p(out))
        return out

    def keyswitch_accum(self, digits, key, c1, c2, which=None):
        """self = (possibly expanded) ring the accumulation happens in."""
        digits, key = _arr(digits), _arr(key)
        c1, c2 = _arr(c1).copy(), _arr(c2).copy()
        D, two, Lk, N = key.shape[-4:]
        assert two == 2 and digits.shape == (digits.shape[0], self.L, self.N)
        which = list(range(self.L)) if which is None else list(which)
        w = (C.c_int * self.L)(*which)
        lib().orc_keyswitch_accum(self.h, C.c_int(digits.shape[0]), C.c_int(Lk), w, _p(digits), _p(key), _p(c1), _p(c2))
        return c1, c2


def ndigits(qs, w: int) -> int:
    q = _arr(qs)
    return lib().orc_ndigits(C.c_int(len(qs)), _p(q), C.c_int(w))


def bfv_switch(N, q_from, q_to, a):
    a = _arr(a); qf, qt = _arr(q_from), _arr(q_to)
    polys = a.size // (len(qf) * N)
    out = np.empty(a.shape[:-2] + (len(qt), N), dtype=np.uint64)
    lib().orc_bfv_switch(C.c_uint64(N), C.c_int(len(qf)), _p(qf), C.c_int(len(qt)), _p(qt), _p(a), _p(out), C.c_long(polys))
    return out


def bfv_contract(N, qs, qs_big, t, a):
    a = _arr(a); q, qb = _arr(qs), _arr(qs_big)
    polys = a.size // (len(qb) * N)
    out = np.empty(a.shape[:-2] + (len(q), N), dtype=np.uint64)
    lib().orc_bfv_contract(C.c_uint64(N), C.c_int(len(q)), _p(q), C.c_int(len(qb)), _p(qb), C.c_uint64(t), _p(a), _p(out), C.c_long(polys))
    return out


def bfv_mul(rq: Rns, rb: Rns, t: int, c1, c2):
    c1, c2 = _arr(c1), _arr(c2)
    B = c1.size // (2 * rq.L * rq.N)
    out = np.empty(c1.shape[:-3] + (3, rq.L, rq.N), dtype=np.uint64)
    lib().orc_bfv_mul(rq.h, rb.h, C.c_uint64(t), _p(c1), _p(c2), _p(out), C.c_long(B))
    return out


def rns_to_ints(N, qs, a):
    """exact CRT reconstruction -> list of Python ints (tests only)."""
    a = _arr(a); q = _arr(qs)
    nl = len(qs) + 1
    out = np.zeros((N, nl), dtype=np.uint64)
    lib().orc_rns_to_limbs(C.c_uint64(N), C.c_int(len(qs)), _p(q), _p(a), _p(out), C.c_int(nl))
    return [sum(int(out[k, j]) << (64 * j) for j in range(nl)) for k in range(N)]
