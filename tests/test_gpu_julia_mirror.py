"""Replays, through RAW ctypes (no engine.py wrappers), the exact call sequences of the overrides in julia/ToyFHEB200.jl:
same symbols, same argument order, same buffer shapes (a Julia Array{UInt64}(N, L, k) is column-major, i.e. the C-order
numpy array [k][L][N]).  The Julia file cannot run here; this is the executable statement of what it does."""
import ctypes as C

import numpy as np
import pytest

import toyfhe_b200 as T
from oracle import c_oracle as CO
from oracle import toyfhe_oracle as O

pytestmark = pytest.mark.gpu

u64p = C.POINTER(C.c_uint64)
NULL = C.c_void_p(None)


def P(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)


@pytest.fixture(scope="module")
def lib():
    lib = T.load_library()
    lib.tfb_last_error.restype = C.c_char_p
    return lib


def check(lib, rc):
    assert rc == 0, lib.tfb_last_error().decode()


def context(lib, N, qs, psis, device=0):
    """ToyFHEB200.context: tfb_ctx_create(DEVICE, degree, length(q), q, psi, out)"""
    out = C.c_void_p()
    q = (C.c_uint64 * len(qs))(*qs)
    psi = (C.c_uint64 * len(qs))(*psis)
    check(lib, lib.tfb_ctx_create(C.c_int(device), C.c_uint32(N), C.c_uint32(len(qs)), q, psi, C.byref(out)))
    return out


def rnd(rng, qs, N, k):
    a = np.empty((k, len(qs), N), dtype=np.uint64)
    for i, q in enumerate(qs):
        a[:, i] = rng.integers(0, q, size=(k, N), dtype=np.uint64)
    return a


def test_nntt_inntt_modswitch(lib):
    N = 64
    qs, psis = T.prime_chain(N, [60, 40, 40])
    ctx = context(lib, N, qs, psis)
    orc = CO.Rns(N, qs, psis)
    rng = np.random.default_rng(1)
    a = rnd(rng, qs, N, 1)
    buf = a.copy()                                              # staging(N, L, 1, :ntt)
    check(lib, lib.tfb_ntt_fwd_host(ctx, P(buf), P(buf), C.c_uint64(len(qs)), NULL))
    assert np.array_equal(buf, orc.nntt(a))
    check(lib, lib.tfb_ntt_inv_host(ctx, P(buf), P(buf), C.c_uint64(len(qs)), NULL))
    assert np.array_equal(buf, a)
    out = np.empty((1, len(qs) - 1, N), dtype=np.uint64)         # staging(N, L - 1, 1, :rs_out)
    check(lib, lib.tfb_rescale_host(ctx, P(a), P(out), C.c_uint64(1), NULL))
    want = O.modswitch([[int(v) for v in r] for r in a[0]], qs)
    assert [[int(v) for v in r] for r in out[0]] == want
    lib.tfb_ctx_destroy(ctx)


@pytest.mark.parametrize("bfv", [True, False])
def test_enc_mul_override(lib, bfv):
    """ToyFHE.enc_mul(c1, c2): BFVParams -> tfb_bfv_mul_host(ctx, ctx_big, t, a, b, out, 1, NULL); otherwise
    tfb_ct_tensor_host(ctx, a, b, out, 1, NULL); a, b = Array(N, L, 2), out = Array(N, L, 3)"""
    N, L, Lb, t = 2048, 2, 4, 53                                 # test/bfv_crt.jl's ring
    allq, allpsi = T.prime_chain(N, [50] * (L + Lb))
    qs, psis, qb, psib = allq[:L], allpsi[:L], allq[L:], allpsi[L:]
    ctx, ctxb = context(lib, N, qs, psis), context(lib, N, qb, psib)
    oq, ob = CO.Rns(N, qs, psis), CO.Rns(N, qb, psib)
    rng = np.random.default_rng(2)
    a, b = rnd(rng, qs, N, 2), rnd(rng, qs, N, 2)
    out = np.empty((3, L, N), dtype=np.uint64)
    if bfv:
        check(lib, lib.tfb_bfv_mul_host(ctx, ctxb, C.c_uint64(t), P(a), P(b), P(out), C.c_uint64(1), NULL))
        assert np.array_equal(out[None], CO.bfv_mul(oq, ob, t, a[None], b[None]))
    else:
        check(lib, lib.tfb_ct_tensor_host(ctx, P(a), P(b), P(out), C.c_uint64(1), NULL))
        assert np.array_equal(out[None], oq.ct_tensor(a[None], b[None]))
    lib.tfb_ctx_destroy(ctx); lib.tfb_ctx_destroy(ctxb)


@pytest.mark.parametrize("raised,w,comps", [(True, 0, 2), (True, 0, 3), (False, 2, 3), (False, 1, 2)])
def test_keyswitch_override(lib, raised, w, comps):
    """ToyFHE.keyswitch(ek, c): DeviceKey uploaded once as Array(N, L', 2, D) over the residues downswitch_keyelement selects,
    then h2d(ct) -> tfb_keyswitch(ctx, ext, w, key, D, din, NC, dout, 1, NULL) -> d2h -> tfb_sync"""
    N = 256
    qs_key, psis_key = T.prime_chain(N, [60, 40, 40, 60] if raised else [60, 60, 40])
    l = 2 if raised else 3                                        # ciphertext level: primes 1..l of the key ring
    which = list(range(l)) + [len(qs_key) - 1] if raised else list(range(l))
    qs, psis = qs_key[:l], psis_key[:l]
    qsel, psel = [qs_key[i] for i in which], [psis_key[i] for i in which]
    ctx, kctx = context(lib, N, qs, psis), context(lib, N, qsel, psel)
    rng = np.random.default_rng(3 + w)
    D = len(qs_key) if w == 0 else T.ndigits(qs, w)              # length(ek.key)
    key_primal = np.stack([rnd(rng, qsel, N, 2) for _ in range(D)])            # [D][2][L'][N]
    osel = CO.Rns(N, qsel, psel)
    host_key = np.ascontiguousarray(osel.nntt(key_primal))        # coeffs_dual of the crtselect'ed components
    dev = C.c_void_p()
    check(lib, lib.tfb_malloc(kctx, C.c_size_t(host_key.nbytes), C.byref(dev)))
    check(lib, lib.tfb_memcpy_h2d(kctx, dev, P(host_key), C.c_size_t(host_key.nbytes), NULL))
    check(lib, lib.tfb_sync(kctx, NULL))
    ct = rnd(rng, qs, N, comps)                                   # staging(N, L, NC, :ks_in)
    out = np.empty((2, l, N), dtype=np.uint64)
    din, dout = C.c_void_p(), C.c_void_p()
    check(lib, lib.tfb_malloc(ctx, C.c_size_t(ct.nbytes), C.byref(din)))
    check(lib, lib.tfb_malloc(ctx, C.c_size_t(out.nbytes), C.byref(dout)))
    check(lib, lib.tfb_memcpy_h2d(ctx, din, P(ct), C.c_size_t(ct.nbytes), NULL))
    check(lib, lib.tfb_keyswitch(ctx, kctx if raised else NULL, C.c_uint32(w), dev, C.c_uint32(D), din, C.c_uint32(comps), dout,
                                 C.c_uint64(1), NULL))
    check(lib, lib.tfb_memcpy_d2h(ctx, P(out), dout, C.c_size_t(out.nbytes), NULL))
    check(lib, lib.tfb_sync(ctx, NULL))
    # the same operation through the tested Python binding (itself bit-exact against the oracle in test_gpu_parity.py)
    pc, pk = T.Context(N, qs, psis), T.Context(N, qsel, psel)
    want = pc.to_host(pc.keyswitch(pk.to_device(host_key) if raised else pc.to_device(host_key), pc.to_device(ct[None]), w,
                                   ext=pk if raised else None))
    assert np.array_equal(out[None], want)
    for p in (din, dout):
        check(lib, lib.tfb_free(ctx, p))
    check(lib, lib.tfb_free(kctx, dev))
    lib.tfb_ctx_destroy(ctx); lib.tfb_ctx_destroy(kctx)


def test_apply_galois_element_override(lib):
    """NTT.apply_galois_element(re, g): h2d -> tfb_galois(ctx, g, din, dout, L, NULL) -> d2h"""
    N = 128
    qs, psis = T.prime_chain(N, [60, 40])
    ctx = context(lib, N, qs, psis)
    rng = np.random.default_rng(5)
    a = rnd(rng, qs, N, 1)
    din, dout = C.c_void_p(), C.c_void_p()
    check(lib, lib.tfb_malloc(ctx, C.c_size_t(a.nbytes), C.byref(din)))
    check(lib, lib.tfb_malloc(ctx, C.c_size_t(a.nbytes), C.byref(dout)))
    for g in (3, 2 * N - 1, pow(3, 2 * N - 8, 2 * N)):
        check(lib, lib.tfb_memcpy_h2d(ctx, din, P(a), C.c_size_t(a.nbytes), NULL))
        check(lib, lib.tfb_galois(ctx, C.c_uint64(g), din, dout, C.c_uint64(len(qs)), NULL))
        got = np.empty_like(a)
        check(lib, lib.tfb_memcpy_d2h(ctx, P(got), dout, C.c_size_t(a.nbytes), NULL))
        check(lib, lib.tfb_sync(ctx, NULL))
        for i, q in enumerate(qs):
            assert [int(v) for v in got[0, i]] == O.apply_galois_element([int(v) for v in a[0, i]], g, q)
    lib.tfb_ctx_destroy(ctx)
