"""ctypes binding of libtoyfhe_b200.so -- the Python stand-in for the Julia
``ccall`` shim (julia/ToyFHEB200.jl).  Torch is used only for
device memory and streams; every operation is one call through the C-ABI.

There is no CPU fallback: if the library is missing or a CUDA call fails the
binding raises (``EngineError``)."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libtoyfhe_b200.so")

_u64p = C.POINTER(C.c_uint64)
_lib = None

# every symbol include/toyfhe_b200.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "tfb_last_error", "tfb_version", "tfb_kernel_launches",
    "tfb_profile_enable", "tfb_profile_classes", "tfb_profile_class_name", "tfb_profile_read",
    "tfb_debug_force_generic", "tfb_debug_ntt_version", "tfb_debug_ntt_force_harvey", "tfb_debug_ntt_max_mode", "tfb_debug_ntt_cross",
    "tfb_prime_chain", "tfb_minimal_primitive_root", "tfb_ndigits",
    "tfb_ctx_create", "tfb_ctx_destroy", "tfb_ctx_info",
    "tfb_malloc", "tfb_free", "tfb_memcpy_h2d", "tfb_memcpy_d2h", "tfb_sync",
    "tfb_ntt_fwd", "tfb_ntt_inv", "tfb_add", "tfb_sub", "tfb_mul", "tfb_neg", "tfb_scalar_mul", "tfb_mul_plain", "tfb_add_plain", "tfb_lincomb",
    "tfb_ring_mul", "tfb_galois", "tfb_rescale", "tfb_crt_expand",
    "tfb_ct_tensor", "tfb_bfv_switch", "tfb_bfv_contract", "tfb_bfv_mul",
    "tfb_keyswitch_digits", "tfb_keyswitch", "tfb_keyswitch_shard", "tfb_keyswitch_shard_push",
    "tfb_xchg_create", "tfb_xchg_export", "tfb_xchg_attach_ipc", "tfb_xchg_attach_ptr", "tfb_xchg_local", "tfb_xchg_check", "tfb_xchg_destroy",
    "tfb_centered_mod", "tfb_ckks_encode", "tfb_ckks_decode", "tfb_sample_uniform", "tfb_sample_gaussian", "tfb_bfv_encode", "tfb_bfv_decode", "tfb_bfv_encode_host", "tfb_bfv_decode_host",
    "tfb_ntt_fwd_host", "tfb_ntt_inv_host", "tfb_ring_mul_host", "tfb_ct_tensor_host",
    "tfb_bfv_mul_host", "tfb_rescale_host",
]


class EngineError(RuntimeError):
    """Raised for any non-zero return code of the C-ABI (the Julia shim throws
    the same way where the reference would raise / @assert)."""


def load_library():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EngineError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        lib.tfb_last_error.restype = C.c_char_p
        lib.tfb_kernel_launches.restype = C.c_ulonglong
        lib.tfb_ctx_create.argtypes = [C.c_int, C.c_uint32, C.c_uint32, _u64p, _u64p, C.POINTER(C.c_void_p)]
        lib.tfb_ctx_destroy.argtypes = [C.c_void_p]
        _lib = lib
    return _lib


def _check(rc: int):
    if rc != 0:
        raise EngineError(f"toyfhe_b200 error {rc}: {load_library().tfb_last_error().decode()}")


def kernel_launches() -> int:
    return int(load_library().tfb_kernel_launches())


def force_generic(on) -> None:
    """testing hook: 1/True = generic base-conversion kernels instead of the specialised ones; 2 = specialised kernels
    with the generic 128-bit reduction instead of the Solinas folds for 2^60 + e primes; 0/False = default"""
    _check(load_library().tfb_debug_force_generic(C.c_int(int(on))))


def ntt_version(v: int) -> None:
    """testing hook: select the row-kernel generation (1 = one CTA per row, 3 = persistent third generation, the default)"""
    _check(load_library().tfb_debug_ntt_version(C.c_int(int(v))))


def ntt_force_harvey(on: bool) -> None:
    """testing hook: disable the lazy forward ladder"""
    _check(load_library().tfb_debug_ntt_force_harvey(C.c_int(1 if on else 0)))


def ntt_cross(on: bool) -> None:
    """testing hook: N > 2^14 forward out of place with (default) or without the last global level applied on load"""
    _check(load_library().tfb_debug_ntt_cross(C.c_int(1 if on else 0)))


def ntt_max_mode(m: int) -> None:
    """testing hook: cap the forward ladder's range policy (0 Harvey, 1 lazy, 2 lazy + approximate quotient)"""
    _check(load_library().tfb_debug_ntt_max_mode(C.c_int(int(m))))


def profile_enable(on: bool) -> None:
    _check(load_library().tfb_profile_enable(C.c_int(1 if on else 0)))


def profile_read(reset: bool = True) -> dict:
    """per-kernel-class {name: (launches, total_ms)} measured with CUDA events"""
    lib = load_library()
    n = lib.tfb_profile_classes()
    counts = (C.c_ulonglong * n)()
    ms = (C.c_double * n)()
    _check(lib.tfb_profile_read(counts, ms, C.c_int(1 if reset else 0)))
    lib.tfb_profile_class_name.restype = C.c_char_p
    return {lib.tfb_profile_class_name(i).decode(): (int(counts[i]), float(ms[i])) for i in range(n)}


# ---- host-only helpers (no GPU needed) -------------------------------------
def prime_chain(N: int, logqs: Sequence[int]):
    """NegacyclicRing(N, logqs) prime chain + minimal roots (crt.jl:282-295)."""
    lib = load_library()
    n = len(logqs)
    lq = (C.c_int32 * n)(*[int(x) for x in logqs])
    q = (C.c_uint64 * n)()
    psi = (C.c_uint64 * n)()
    _check(lib.tfb_prime_chain(C.c_uint32(N), lq, C.c_uint32(n), q, psi))
    return [int(x) for x in q], [int(x) for x in psi]


def minimal_primitive_root(q: int, n: int) -> int:
    out = C.c_uint64()
    _check(load_library().tfb_minimal_primitive_root(C.c_uint64(q), C.c_uint64(n), C.byref(out)))
    return int(out.value)


def ndigits(qs: Sequence[int], w: int) -> int:
    arr = (C.c_uint64 * len(qs))(*[int(x) for x in qs])
    out = C.c_uint32()
    _check(load_library().tfb_ndigits(arr, C.c_uint32(len(qs)), C.c_uint32(w), C.byref(out)))
    return int(out.value)


# ---- device context ----------------------------------------------------------
def _ptr(t) -> C.c_void_p:
    """device (torch tensor) or host (numpy / pinned torch tensor) buffer -> pointer"""
    if isinstance(t, np.ndarray):
        assert t.dtype == np.uint64 and t.flags["C_CONTIGUOUS"]
        return C.c_void_p(t.ctypes.data)
    assert t.is_contiguous() and t.element_size() == 8, "need a contiguous 64-bit tensor"
    return C.c_void_p(t.data_ptr())


def _stream_ptr(stream, device=None) -> C.c_void_p:
    """``stream`` or torch's current stream OF THE CONTEXT'S DEVICE (not of whatever device happens to be current)"""
    if stream is None:
        import torch
        stream = torch.cuda.current_stream(device)
    return C.c_void_p(stream.cuda_stream)


class Context:
    """One NegacyclicRing{CRTEncoded{L}, N}(psi) resident on a GPU
    (pow2_cyc_rings.jl:27-47; crt.jl:282-295)."""

    def __init__(self, N: int, qs: Sequence[int], psis: Sequence[int], device: Optional[int] = None):
        import torch
        if not torch.cuda.is_available():
            raise EngineError("no CUDA device: the engine has no CPU fallback")
        lib = load_library()
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.N, self.L = int(N), len(qs)
        self.qs, self.psis = [int(q) for q in qs], [int(p) for p in psis]
        q = (C.c_uint64 * self.L)(*self.qs)
        psi = (C.c_uint64 * self.L)(*self.psis)
        h = C.c_void_p()
        _check(lib.tfb_ctx_create(C.c_int(self.device), C.c_uint32(self.N), C.c_uint32(self.L), q, psi, C.byref(h)))
        self.h = h
        self._lib = lib

    def close(self):
        if getattr(self, "h", None):
            self._lib.tfb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- allocation helpers (torch owns the memory)
    def empty(self, *shape):
        import torch
        return torch.empty(*shape, dtype=torch.int64, device=f"cuda:{self.device}")

    def to_device(self, a: np.ndarray):
        import torch
        a = np.ascontiguousarray(a, dtype=np.uint64)
        return torch.from_numpy(a.view(np.int64)).to(f"cuda:{self.device}")

    @staticmethod
    def to_host(t) -> np.ndarray:
        return t.detach().cpu().numpy().view(np.uint64)

    def _rows(self, t) -> int:
        n = t.numel() if hasattr(t, "numel") else t.size
        assert n % (self.N * self.L) == 0, "buffer must hold whole RNS polynomials [..][L][N]"
        return n // self.N

    def _polys(self, t) -> int:
        return self._rows(t) // self.L

    # -- transforms
    def ntt_fwd(self, a, out=None, stream=None):
        out = self.empty(a.shape) if out is None else out
        _check(self._lib.tfb_ntt_fwd(self.h, _ptr(a), _ptr(out), C.c_uint64(self._rows(a)), _stream_ptr(stream, self.device)))
        return out

    def ntt_inv(self, a, out=None, stream=None):
        out = self.empty(a.shape) if out is None else out
        _check(self._lib.tfb_ntt_inv(self.h, _ptr(a), _ptr(out), C.c_uint64(self._rows(a)), _stream_ptr(stream, self.device)))
        return out

    def _bin(self, fn, a, b, out, stream):
        assert a.shape == b.shape
        out = self.empty(a.shape) if out is None else out
        _check(fn(self.h, _ptr(a), _ptr(b), _ptr(out), C.c_uint64(self._rows(a)), _stream_ptr(stream, self.device)))
        return out

    def add(self, a, b, out=None, stream=None): return self._bin(self._lib.tfb_add, a, b, out, stream)
    def sub(self, a, b, out=None, stream=None): return self._bin(self._lib.tfb_sub, a, b, out, stream)
    def mul(self, a, b, out=None, stream=None): return self._bin(self._lib.tfb_mul, a, b, out, stream)
    def ring_mul(self, a, b, out=None, stream=None): return self._bin(self._lib.tfb_ring_mul, a, b, out, stream)

    def neg(self, a, out=None, stream=None):
        out = self.empty(a.shape) if out is None else out
        _check(self._lib.tfb_neg(self.h, _ptr(a), _ptr(out), C.c_uint64(self._rows(a)), _stream_ptr(stream, self.device)))
        return out

    def scalar_mul(self, a, s: int, out=None, stream=None):
        out = self.empty(a.shape) if out is None else out
        sr = (C.c_uint64 * self.L)(*[int(s) % q for q in self.qs])
        _check(self._lib.tfb_scalar_mul(self.h, _ptr(a), sr, _ptr(out), C.c_uint64(self._rows(a)), _stream_ptr(stream, self.device)))
        return out

    def mul_plain(self, a, plain, out=None, accumulate=False, stream=None):
        """out[p] (+)= a[p] (.) plain over every polynomial p of ``a`` (plain: one [L][N] element, same domain)"""
        assert plain.numel() == self.N * self.L
        out = self.empty(a.shape) if out is None else out
        _check(self._lib.tfb_mul_plain(self.h, _ptr(a), _ptr(plain), _ptr(out), C.c_uint64(self._polys(a)), C.c_int(1 if accumulate else 0),
                                       _stream_ptr(stream, self.device)))
        return out

    def lincomb(self, stacked, weights, out=None, stream=None):
        """out[c] = sum_j weights[c][j] * stacked[j]: stacked [J][..][L][N] (one contiguous tensor), weights a device int64 tensor
        [C][J][L] of residues; out [C][..][L][N]"""
        J, Cn = stacked.shape[0], weights.shape[0]
        assert stacked.is_contiguous() and weights.is_contiguous() and tuple(weights.shape) == (Cn, J, self.L)
        polys = self._polys(stacked[0])
        out = self.empty((Cn,) + tuple(stacked.shape[1:])) if out is None else out
        _check(self._lib.tfb_lincomb(self.h, _ptr(stacked), C.c_uint64(stacked[0].numel()), C.c_uint32(J), _ptr(weights), C.c_uint32(Cn),
                                     _ptr(out), C.c_uint64(polys), _stream_ptr(stream, self.device)))
        return out

    def add_plain_first(self, ct, plain, stream=None):
        """ct[b][0] += plain in place for every ciphertext b of ct [B][comps][L][N] (ckksencoding.jl:113-125)"""
        assert plain.numel() == self.N * self.L and ct.is_contiguous()
        B, comps = ct.shape[0], ct.shape[1]
        _check(self._lib.tfb_add_plain(self.h, _ptr(ct), _ptr(plain), _ptr(ct), C.c_uint64(B), C.c_uint64(comps * self.L * self.N),
                                       _stream_ptr(stream, self.device)))
        return ct

    def galois(self, a, g: int, out=None, stream=None):
        out = self.empty(a.shape) if out is None else out
        _check(self._lib.tfb_galois(self.h, C.c_uint64(int(g)), _ptr(a), _ptr(out), C.c_uint64(self._rows(a)), _stream_ptr(stream, self.device)))
        return out

    # -- level changes
    def rescale(self, a, out=None, stream=None):
        out = self.empty(tuple(a.shape[:-2]) + (self.L - 1, self.N)) if out is None else out
        _check(self._lib.tfb_rescale(self.h, _ptr(a), _ptr(out), C.c_uint64(self._polys(a)), _stream_ptr(stream, self.device)))
        return out

    def crt_expand(self, a, P: int, out=None, stream=None):
        out = self.empty(tuple(a.shape[:-2]) + (self.L + 1, self.N)) if out is None else out
        _check(self._lib.tfb_crt_expand(self.h, C.c_uint64(int(P)), _ptr(a), _ptr(out), C.c_uint64(self._polys(a)), _stream_ptr(stream, self.device)))
        return out

    # -- ciphertext multiply
    def _batch(self, c, comps):
        n = c.numel() if hasattr(c, "numel") else c.size
        assert n % (comps * self.L * self.N) == 0
        return n // (comps * self.L * self.N)

    def ct_tensor(self, c1, c2, out=None, stream=None):
        B = self._batch(c1, 2)
        out = self.empty(tuple(c1.shape[:-3]) + (3, self.L, self.N)) if out is None else out
        _check(self._lib.tfb_ct_tensor(self.h, _ptr(c1), _ptr(c2), _ptr(out), C.c_uint64(B), _stream_ptr(stream, self.device)))
        return out

    def bfv_switch(self, to: "Context", a, out=None, stream=None):
        out = to.empty(tuple(a.shape[:-2]) + (to.L, to.N)) if out is None else out
        _check(self._lib.tfb_bfv_switch(self.h, to.h, _ptr(a), _ptr(out), C.c_uint64(self._polys(a)), _stream_ptr(stream, self.device)))
        return out

    def bfv_contract(self, big: "Context", t: int, a, out=None, stream=None):
        out = self.empty(tuple(a.shape[:-2]) + (self.L, self.N)) if out is None else out
        _check(self._lib.tfb_bfv_contract(self.h, big.h, C.c_uint64(int(t)), _ptr(a), _ptr(out), C.c_uint64(big._polys(a)), _stream_ptr(stream, self.device)))
        return out

    def bfv_mul(self, big: "Context", t: int, c1, c2, out=None, stream=None):
        B = self._batch(c1, 2)
        out = self.empty(tuple(c1.shape[:-3]) + (3, self.L, self.N)) if out is None else out
        _check(self._lib.tfb_bfv_mul(self.h, big.h, C.c_uint64(int(t)), _ptr(c1), _ptr(c2), _ptr(out), C.c_uint64(B), _stream_ptr(stream, self.device)))
        return out

    def centered_mod(self, t: int, b, out=None, stream=None):
        """mod(SignedMod(b), t) per coefficient (BGV pi, bgv.jl:22-25): b [polys][L][N] -> [polys][N]"""
        out = self.empty(tuple(b.shape[:-2]) + (self.N,)) if out is None else out
        _check(self._lib.tfb_centered_mod(self.h, C.c_uint64(int(t)), _ptr(b), _ptr(out), C.c_uint64(self._polys(b)), _stream_ptr(stream, self.device)))
        return out

    # -- CKKS encoding (ckksencoding.jl:60-101): slots are complex128 device tensors [polys][N/2]
    def ckks_encode(self, scale: float, slots, out=None, stream=None):
        polys = int(slots.numel() // (self.N // 2))
        out = self.empty(tuple(slots.shape[:-1]) + (self.L, self.N)) if out is None else out
        _check(self._lib.tfb_ckks_encode(self.h, C.c_double(float(scale)), C.c_void_p(slots.data_ptr()), _ptr(out), C.c_uint64(polys), _stream_ptr(stream, self.device)))
        return out

    def ckks_decode(self, scale: float, a, out=None, stream=None):
        import torch
        if out is None:
            out = torch.empty(tuple(a.shape[:-2]) + (self.N // 2,), dtype=torch.complex128, device=a.device)
        _check(self._lib.tfb_ckks_decode(self.h, C.c_double(float(scale)), _ptr(a), C.c_void_p(out.data_ptr()), C.c_uint64(self._polys(a)), _stream_ptr(stream, self.device)))
        return out

    # -- sampling on the device (poly.jl:7-23)
    def sample_uniform(self, seed: int, stream_id: int, polys_shape=(1,), out=None, stream=None):
        out = self.empty(tuple(polys_shape) + (self.L, self.N)) if out is None else out
        _check(self._lib.tfb_sample_uniform(self.h, C.c_uint64(int(seed)), C.c_uint32(int(stream_id)), _ptr(out),
                                            C.c_uint64(self._polys(out)), _stream_ptr(stream, self.device)))
        return out

    def sample_gaussian(self, sigma: float, seed: int, stream_id: int, polys_shape=(1,), out=None, stream=None):
        out = self.empty(tuple(polys_shape) + (self.L, self.N)) if out is None else out
        _check(self._lib.tfb_sample_gaussian(self.h, C.c_double(float(sigma)), C.c_uint64(int(seed)), C.c_uint32(int(stream_id)),
                                             _ptr(out), C.c_uint64(self._polys(out)), _stream_ptr(stream, self.device)))
        return out

    # -- BFV plaintext maps (bfv.jl:21-29)
    @staticmethod
    def _limbs(x: int):
        x = int(x)
        if x <= 0:
            raise EngineError("Delta must be a positive integer")
        n = (x.bit_length() + 63) // 64
        arr = (C.c_uint64 * n)(*[(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)])
        return arr, n

    def bfv_encode(self, t: int, delta: int, m, out=None, stream=None):
        """m [polys][N] (words, reduced mod t) -> Delta*m over this ring's primes [polys][L][N]"""
        arr, n = self._limbs(delta)
        polys = int(m.numel() // self.N)
        out = self.empty(tuple(m.shape[:-1]) + (self.L, self.N)) if out is None else out
        _check(self._lib.tfb_bfv_encode(self.h, C.c_uint64(int(t)), arr, C.c_uint32(n), _ptr(m), _ptr(out), C.c_uint64(polys), _stream_ptr(stream, self.device)))
        return out

    def bfv_decode(self, t: int, delta: int, b, out=None, stream=None):
        """b [polys][L][N] primal -> mod(divround(SignedMod(b), Delta), t) [polys][N]"""
        arr, n = self._limbs(delta)
        out = self.empty(tuple(b.shape[:-2]) + (self.N,)) if out is None else out
        _check(self._lib.tfb_bfv_decode(self.h, C.c_uint64(int(t)), arr, C.c_uint32(n), _ptr(b), _ptr(out), C.c_uint64(self._polys(b)), _stream_ptr(stream, self.device)))
        return out

    # -- key switching
    def keyswitch_digits(self, cend, w: int, target: Optional["Context"] = None, out=None, stream=None):
        target = self if target is None else target
        D = self.L if w == 0 else ndigits(self.qs, w)
        B = self._polys(cend)
        out = self.empty(tuple(cend.shape[:-2]) + (D, target.L, self.N)) if out is None else out
        _check(self._lib.tfb_keyswitch_digits(self.h, target.h, C.c_uint32(w), _ptr(cend), _ptr(out), C.c_uint64(B), _stream_ptr(stream, self.device)))
        return out

    def keyswitch(self, key_dual, ct, w: int, ext: Optional["Context"] = None, out=None, stream=None):
        """ct [B][comps][L][N] primal, key_dual [D][2][L'][N] (NTT domain)."""
        comps = ct.shape[-3]
        B = self._batch(ct, comps)
        D = key_dual.shape[0]
        out = self.empty(tuple(ct.shape[:-3]) + (2, self.L, self.N)) if out is None else out
        _check(self._lib.tfb_keyswitch(self.h, ext.h if ext is not None else None, C.c_uint32(w), _ptr(key_dual),
                                       C.c_uint32(D), _ptr(ct), C.c_uint32(comps), _ptr(out), C.c_uint64(B), _stream_ptr(stream, self.device)))
        return out

    def keyswitch_shard(self, shard: "Context", first: int, key_dual_shard, ct, w: int, out=None, stream=None):
        """Rows first..first+shard.L-1 of keyswitch(key, ct): ct [B][comps][L][N] whole, key_dual_shard [D][2][Ls][N]."""
        comps = ct.shape[-3]
        B = self._batch(ct, comps)
        out = shard.empty(tuple(ct.shape[:-3]) + (2, shard.L, self.N)) if out is None else out
        _check(self._lib.tfb_keyswitch_shard(self.h, shard.h, C.c_uint32(first), C.c_uint32(w), _ptr(key_dual_shard),
                                             C.c_uint32(key_dual_shard.shape[0]), _ptr(ct), C.c_uint32(comps), _ptr(out),
                                             C.c_uint64(B), _stream_ptr(stream, self.device)))
        return out

    def keyswitch_shard_push(self, shard: "Context", first: int, key_dual_shard, ct, w: int, xchg: "PeerExchange", stream=None):
        """keyswitch_shard with the exchange fused into its last kernel: returns the WHOLE result [B][2][L][N] as a view of
        this rank's exchange slot (valid until the call after next); every rank of the exchange makes the same call."""
        comps = ct.shape[-3]
        B = self._batch(ct, comps)
        res = C.c_void_p()
        _check(self._lib.tfb_keyswitch_shard_push(self.h, shard.h, C.c_uint32(first), C.c_uint32(w), _ptr(key_dual_shard),
                                                  C.c_uint32(key_dual_shard.shape[0]), _ptr(ct), C.c_uint32(comps), xchg.h,
                                                  C.byref(res), C.c_uint64(B), _stream_ptr(stream, self.device)))
        return xchg.view(res.value, tuple(ct.shape[:-3]) + (2, self.L, self.N))

    # -- host-buffer entry points (numpy uint64 or pinned torch tensors)
    def ntt_fwd_host(self, a, out, stream=None):
        _check(self._lib.tfb_ntt_fwd_host(self.h, _ptr(a), _ptr(out), C.c_uint64(self._rows(a)), _stream_ptr(stream, self.device)))
        return out

    def ntt_inv_host(self, a, out, stream=None):
        _check(self._lib.tfb_ntt_inv_host(self.h, _ptr(a), _ptr(out), C.c_uint64(self._rows(a)), _stream_ptr(stream, self.device)))
        return out

    def ring_mul_host(self, a, b, out, stream=None):
        _check(self._lib.tfb_ring_mul_host(self.h, _ptr(a), _ptr(b), _ptr(out), C.c_uint64(self._rows(a)), _stream_ptr(stream, self.device)))
        return out

    def ct_tensor_host(self, c1, c2, out, stream=None):
        _check(self._lib.tfb_ct_tensor_host(self.h, _ptr(c1), _ptr(c2), _ptr(out), C.c_uint64(self._batch(c1, 2)), _stream_ptr(stream, self.device)))
        return out

    def bfv_mul_host(self, big: "Context", t: int, c1, c2, out, stream=None):
        _check(self._lib.tfb_bfv_mul_host(self.h, big.h, C.c_uint64(int(t)), _ptr(c1), _ptr(c2), _ptr(out),
                                          C.c_uint64(self._batch(c1, 2)), _stream_ptr(stream, self.device)))
        return out

    def bfv_encode_host(self, t: int, delta: int, m, out, stream=None):
        arr, n = self._limbs(delta)
        polys = int((m.numel() if hasattr(m, "numel") else m.size) // self.N)
        _check(self._lib.tfb_bfv_encode_host(self.h, C.c_uint64(int(t)), arr, C.c_uint32(n), _ptr(m), _ptr(out), C.c_uint64(polys), _stream_ptr(stream, self.device)))
        return out

    def bfv_decode_host(self, t: int, delta: int, b, out, stream=None):
        arr, n = self._limbs(delta)
        _check(self._lib.tfb_bfv_decode_host(self.h, C.c_uint64(int(t)), arr, C.c_uint32(n), _ptr(b), _ptr(out), C.c_uint64(self._polys(b)), _stream_ptr(stream, self.device)))
        return out

    def rescale_host(self, a, out, stream=None):
        _check(self._lib.tfb_rescale_host(self.h, _ptr(a), _ptr(out), C.c_uint64(self._polys(a)), _stream_ptr(stream, self.device)))
        return out


class _RawDevice:
    """a device allocation of the engine exposed to torch (torch.as_tensor reads __cuda_array_interface__)"""

    def __init__(self, ptr: int, words: int):
        self.__cuda_array_interface__ = {"shape": (words,), "typestr": "<i8", "data": (ptr, False), "version": 2}


class PeerExchange:
    """tfb_xchg: this rank's result slots + flag words for the sharded keyswitch whose epilogue pushes its rows into every
    rank's slot over NVLink peer memory (include/toyfhe_b200.h).  ``attach`` takes the handles of ALL ranks in rank order
    (as torch.distributed.all_gather_object returns them); ``attach_local`` wires exchanges that live in one process."""

    def __init__(self, ctx: Context, rank: int, world: int, slot_words: int):
        self._lib = load_library()
        self.ctx, self.rank, self.world, self.slot_words = ctx, rank, world, int(slot_words)
        h = C.c_void_p()
        _check(self._lib.tfb_xchg_create(ctx.h, C.c_uint32(rank), C.c_uint32(world), C.c_uint64(8 * self.slot_words), C.byref(h)))
        self.h = h
        base = C.c_void_p()
        _check(self._lib.tfb_xchg_local(self.h, C.byref(base)))
        self.base = base.value
        import torch
        with torch.cuda.device(ctx.device):
            self._slots = torch.as_tensor(_RawDevice(self.base, 2 * self.slot_words), device=f"cuda:{ctx.device}")

    def handle(self) -> bytes:
        buf = (C.c_uint8 * 64)()
        _check(self._lib.tfb_xchg_export(self.h, buf))
        return bytes(buf)

    def attach(self, handles):
        for p, hb in enumerate(handles):
            if p != self.rank:
                _check(self._lib.tfb_xchg_attach_ipc(self.h, C.c_uint32(p), (C.c_uint8 * 64).from_buffer_copy(hb)))

    def attach_local(self, peers):
        for p, other in enumerate(peers):
            if p != self.rank:
                _check(self._lib.tfb_xchg_attach_ptr(self.h, C.c_uint32(p), C.c_void_p(other.base)))

    def view(self, ptr: int, shape):
        off = (ptr - self.base) // 8
        n = int(np.prod(shape))
        return self._slots[off:off + n].view(*shape)

    def timed_out(self) -> bool:
        t = C.c_int()
        _check(self._lib.tfb_xchg_check(self.h, C.byref(t)))
        return bool(t.value)

    def close(self):
        if getattr(self, "h", None):
            self._slots = None
            _check(self._lib.tfb_xchg_destroy(self.h))
            self.h = None
