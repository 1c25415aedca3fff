"""Validate the C restatement (oracle/oracle.c) against the Python big-int
oracle (itself pinned to the reference KATs).  CPU only."""
import math

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import toyfhe_oracle as O


def _rand_poly(rng, N, qs, shape=()):
    out = np.empty(shape + (len(qs), N), dtype=np.uint64)
    for i, q in enumerate(qs):
        out[..., i, :] = rng.integers(0, q, size=shape + (N,), dtype=np.uint64)
    return out


def _tolist(a):
    return [[int(x) for x in row] for row in a]


@pytest.fixture(scope="module")
def small():
    N = 64
    qs, psis = O.prime_chain(N, (60, 60, 40))
    return N, qs, psis, CO.Rns(N, qs, psis)


def test_doc_kat_through_c():
    r = CO.Rns(4, [97], [33])
    a = np.array([[1, 1, 0, 0]], dtype=np.uint64)
    b = np.array([[0, 0, 0, 1]], dtype=np.uint64)
    assert r.nntt(a).tolist() == [[34, 48, 65, 51]]
    assert r.ring_mul(a, b).tolist() == [[96, 0, 0, 1]]
    assert r.ring_mul(a, a).tolist() == [[1, 2, 1, 0]]
    assert r.galois(np.array([[1, 2, 3, 4]], dtype=np.uint64), 3).tolist() == [[1, 4, 94, 2]]


def test_ntt_matches_python(small):
    N, qs, psis, r = small
    rng = np.random.default_rng(0)
    a = _rand_poly(rng, N, qs, (2,))
    f = r.nntt(a)
    for b in range(2):
        assert _tolist(f[b]) == O.rns_nntt(_tolist(a[b]), qs, psis)
    assert np.array_equal(r.inntt(f), a)


def test_ntt_full_size_first_prime(q8, psi8):
    N = 2 ** 14
    r = CO.Rns(N, q8[:1], psi8[:1])
    rng = np.random.default_rng(1)
    a = _rand_poly(rng, N, q8[:1])
    f = r.nntt(a)
    assert [int(x) for x in f[0]] == O.nntt([int(x) for x in a[0]], q8[0], psi8[0])
    assert np.array_equal(r.inntt(f), a)


def test_elementwise_and_ring_mul(small):
    N, qs, psis, r = small
    rng = np.random.default_rng(2)
    a, b = _rand_poly(rng, N, qs), _rand_poly(rng, N, qs)
    A, B = _tolist(a), _tolist(b)
    assert _tolist(r.add(a, b)) == O.rns_add(A, B, qs)
    assert _tolist(r.sub(a, b)) == O.rns_sub(A, B, qs)
    assert _tolist(r.neg(a)) == O.rns_neg(A, qs)
    assert _tolist(r.scalar_mul(a, 12345678901234567890123)) == O.rns_scalar_mul(A, 12345678901234567890123, qs)
    assert _tolist(r.ring_mul(a, b)) == O.rns_ring_multiply(A, B, qs, psis)
    assert _tolist(r.galois(a, 5)) == O.rns_galois(A, 5, qs)
    assert _tolist(r.modswitch(a)) == O.modswitch(A, qs)


def test_crt_expand(small):
    N, qs, psis, _ = small
    r2 = CO.Rns(N, qs[:2], psis[:2])
    rng = np.random.default_rng(3)
    a = _rand_poly(rng, N, qs[:2])
    assert _tolist(r2.crt_expand(a, qs[2])) == O.rns_crt_expand(_tolist(a), qs[:2], qs[2])


def test_ct_tensor(small):
    N, qs, psis, r = small
    rng = np.random.default_rng(4)
    c1, c2 = _rand_poly(rng, N, qs, (2,)), _rand_poly(rng, N, qs, (2,))
    want = O.ct_tensor([_tolist(c1[0]), _tolist(c1[1])], [_tolist(c2[0]), _tolist(c2[1])], qs, psis)
    got = r.ct_tensor(c1, c2)
    for k in range(3):
        assert _tolist(got[k]) == want[k]


def test_bigint_reconstruct(small):
    N, qs, psis, r = small
    rng = np.random.default_rng(5)
    a = _rand_poly(rng, N, qs)
    assert CO.rns_to_ints(N, qs, a) == O.rns_to_ints(_tolist(a), qs)


@pytest.mark.parametrize("L,Lb,t", [(2, 4, 53), (3, 7, 65537)])
def test_bfv_switch_contract_mul(L, Lb, t):
    N = 32
    allq, allpsi = O.prime_chain(N, (60,) * (L + Lb))
    qs, psis, qb, psib = allq[:L], allpsi[:L], allq[L:], allpsi[L:]
    rng = np.random.default_rng(6)
    c1, c2 = _rand_poly(rng, N, qs, (2,)), _rand_poly(rng, N, qs, (2,))
    # edge values around Q/2 for the strict '>' centring rule (bfv.jl:202-220)
    Q = math.prod(qs)
    for k, X in enumerate([Q >> 1, (Q >> 1) + 1, 0, Q - 1, 1]):
        for i, q in enumerate(qs):
            c1[0, i, k] = X % q
    e1 = CO.bfv_switch(N, qs, qb, c1)
    for comp in range(2):
        assert _tolist(e1[comp]) == O.bfv_switch(_tolist(c1[comp]), qs, qb)
    big = _rand_poly(rng, N, qb, (3,))
    Qb = math.prod(qb)
    # ties/edges: x with t*x exactly half-way is impossible for odd Q; exercise +-Qb/2 edges
    for k, X in enumerate([Qb >> 1, (Qb >> 1) + 1, 0, 1, Qb - 1]):
        for j, p in enumerate(qb):
            big[0, j, k] = X % p
    got = CO.bfv_contract(N, qs, qb, t, big)
    want = O.bfv_mul_contract([_tolist(big[k]) for k in range(3)], qs, qb, t)
    for k in range(3):
        assert _tolist(got[k]) == want[k]
    rq, rb = CO.Rns(N, qs, psis), CO.Rns(N, qb, psib)
    gm = CO.bfv_mul(rq, rb, t, c1, c2)
    wm = O.bfv_mul([_tolist(c1[0]), _tolist(c1[1])], [_tolist(c2[0]), _tolist(c2[1])], qs, psis, qb, psib, t)
    for k in range(3):
        assert _tolist(gm[k]) == wm[k]


@pytest.mark.parametrize("w", [0, 1, 2, 7])
def test_keyswitch_digits_and_accum(small, w):
    N, qs, psis, r = small
    rng = np.random.default_rng(7)
    cend = _rand_poly(rng, N, qs)
    dg = r.keyswitch_digits(cend, w)
    want = O.keyswitch_digits(_tolist(cend), qs, w)
    assert len(want) == dg.shape[0]
    for d in range(dg.shape[0]):
        assert _tolist(dg[d]) == want[d]
    if w in (0, 7):
        D = dg.shape[0]
        key = _rand_poly(rng, N, qs, (D, 2))
        ct = _rand_poly(rng, N, qs, (3,))
        ct[2] = cend
        c1, c2 = r.keyswitch_accum(dg, key, ct[0], ct[1])
        kw = O.keyswitch([_tolist(ct[k]) for k in range(3)], [(_tolist(key[d, 0]), _tolist(key[d, 1])) for d in range(D)], qs, psis, w)
        assert _tolist(c1) == kw[0] and _tolist(c2) == kw[1]


def test_ndigits_config4(q8):
    # Q is 481 bits -> 241 base-4 digits
    assert CO.ndigits(q8, 2) == 241 == O.ndigits(math.prod(q8), 4)
