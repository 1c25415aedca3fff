"""BFV ``mul_contract`` on the reference's own RNS test ring (test/bfv_crt.jl:8-28), against an
INLINE transcription of the Julia lines -- not against the oracle's own reading of them.

The transcription below follows the reference statement by statement and shares no code with
``oracle/``; each Python line names the Julia expression it stands for.  The point (round-1
review): ``multround(SignedMod(x), t, Q)`` multiplies by t *in the CRT field* (modulo Q_big)
before the centred lift, so on this ring (Q ~ 2^100, Q_big ~ 2^200, uniformly random tensor
values) it differs from ``rha(t * centre(x), Q)`` on almost every coefficient.
"""
import math
from fractions import Fraction

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import toyfhe_oracle as O

N_CRT = 2048
T_CRT = 53


def bfv_crt_ring():
    """test/bfv_crt.jl:8-22: p1 = nextprime(2^50+1; interval=2n), p2..p6 the following primes = 1 mod 2n."""
    ps = O.prime_chain(N_CRT, [50] * 6)[0]
    assert ps[:2] == [1125899906949121, 1125899906977793]      # SURVEY.md section 4, test/bfv_crt.jl row
    return ps[:2], ps[2:]


# ---- Julia, transcribed -------------------------------------------------------------------------
def jl_convert_Integer_CRTEncoded(c, moduli):
    """crt.jl:105-112  convert(Integer, x::CRTEncoded) = AbstractAlgebra.crt(residues, moduli); pairwise rule crt.jl:98-103:
    g,u,v = gcdx(m1,m2); mod(r1*v*m2 + r2*u*m1, m1*m2)."""
    r, m = int(c[0]), int(moduli[0])
    for r2, m2 in zip(c[1:], moduli[1:]):
        r2, m2 = int(r2), int(m2)
        u, v = pow(m, -1, m2), pow(m2, -1, m)       # u*m = 1 (mod m2), v*m2 = 1 (mod m)
        r, m = (r * v * m2 + r2 * u * m) % (m * m2), m * m2
    return r


def jl_convert_Integer_SignedMod(c, moduli):
    """signedmod.jl:12-19  n = convert(Integer, x.x); n > div(modulus(x.x), 2) ? n - modulus : n."""
    n = jl_convert_Integer_CRTEncoded(c, moduli)
    m = math.prod(int(p) for p in moduli)
    return n - m if n > m // 2 else n


def jl_div_RoundNearestTiesAway(a, b):
    """div_hacks.jl:120-135 (Julia >= 1.4: Base.div(x, y, RoundNearestTiesAway)) on integers, b > 0:
    the exact quotient a/b rounded to the nearest integer, halves away from zero."""
    f = Fraction(a, b)
    fl = f.numerator // f.denominator
    frac = f - fl
    if frac > Fraction(1, 2) or (frac == Fraction(1, 2) and a >= 0):
        return fl + 1
    return fl


def jl_CRTEncoded_of_Integer(x, moduli):
    """crt.jl:91-95  CRTEncoded{N,M}(x::Integer) = map(T -> T(x), fieldtypes(M)); PrimeField(x) reduces to [0,p)."""
    return [x % int(p) for p in moduli]


def jl_multround(c, a, b, moduli):
    """bfv.jl:188 multround(SignedMod(x), a, b).x  with  bfv.jl:172-174 div(e * a, b, RoundNearestTiesAway),
    signedmod.jl:28 e * a = e * oftype(e, a), signedmod.jl:21,24-26 SignedMod{T}(e.x * T(a)) and crt.jl:124-126
    (CRTEncoded * is componentwise), signedmod.jl:30-32 div(e, x, r) = oftype(e, div(convert(Integer, e), x, r))."""
    Ta = jl_CRTEncoded_of_Integer(a, moduli)                               # oftype(e.x, a)
    prod = [(int(x) * y) % int(p) for x, y, p in zip(c, Ta, moduli)]       # e.x * T(a)
    n = jl_convert_Integer_SignedMod(prod, moduli)                         # convert(Integer, e * a)
    return jl_CRTEncoded_of_Integer(jl_div_RoundNearestTiesAway(n, b), moduli)   # oftype(e, div(...))


def jl_switchel(c, from_moduli, to_moduli):
    """bfv.jl:202-220 switchel(T, e): q = modulus(e); halfq = q >> 1; diff = |modulus(T) - q|; en = convert(Integer, e);
    q < modulus(T): en > halfq ? T(en + diff) : T(en);  else: en > halfq ? T(en - diff) : T(en)."""
    q, mT = math.prod(int(p) for p in from_moduli), math.prod(int(p) for p in to_moduli)
    halfq = q >> 1
    diff = mT - q if mT > q else q - mT
    en = jl_convert_Integer_CRTEncoded(c, from_moduli)
    if q < mT:
        return jl_CRTEncoded_of_Integer(en + diff if en > halfq else en, to_moduli)
    return jl_CRTEncoded_of_Integer(en - diff if en > halfq else en, to_moduli)


def jl_mul_contract_coeff(c, qs, qb, t):
    """bfv.jl:35-40: switch(R, multround(e, modulus(base_ring(Rplain)), modulus(coefftype(R)))) for one coefficient."""
    return jl_switchel(jl_multround(c, t, math.prod(qs), qb), qb, qs)


# ---- tests -----------------------------------------------------------------------------------------
def _random_tensor(rng, qb, n):
    big = np.empty((1, len(qb), n), dtype=np.uint64)
    for j, p in enumerate(qb):
        big[0, j] = rng.integers(0, p, size=n, dtype=np.uint64)
    return big


def _edges(qs, qb, t):
    Q, Qb = math.prod(qs), math.prod(qb)
    tinv = pow(t, -1, Qb)
    xs = [0, 1, Qb - 1, Qb >> 1, (Qb >> 1) + 1, Q >> 1, (Q >> 1) + 1, Q, Q - 1]
    # values whose PRODUCT with t lands on the centring boundary / on rounding boundaries
    for y in [Qb >> 1, (Qb >> 1) + 1, (Qb >> 1) - 1, Q >> 1, (Q >> 1) + 1, Qb - (Q >> 1), Qb - (Q >> 1) - 1, 3 * Q + (Q >> 1),
              3 * Q + (Q >> 1) + 1]:
        xs.append((y * tinv) % Qb)
    return xs


def _want(big, qs, qb, t):
    n = big.shape[-1]
    out = np.empty((big.shape[0], len(qs), n), dtype=np.uint64)
    for p in range(big.shape[0]):
        for k in range(n):
            out[p, :, k] = jl_mul_contract_coeff([int(big[p, j, k]) for j in range(len(qb))], qs, qb, t)
    return out


def test_transcription_differs_from_integer_product_on_bfv_crt_ring():
    """The round-1 formula rha(t*centre(x), Q) is NOT what the reference computes here (197 of 200 in the review)."""
    qs, qb = bfv_crt_ring()
    Q, Qb = math.prod(qs), math.prod(qb)
    rng = np.random.default_rng(5)
    big = _random_tensor(rng, qb, 200)
    want = _want(big, qs, qb, T_CRT)
    differ = 0
    for k in range(200):
        x = O.centre(O.crt_reconstruct([int(big[0, j, k]) for j in range(4)], qb), Qb)
        old = [O.rha(T_CRT * x, Q) % q for q in qs]
        differ += old != [int(v) for v in want[0, :, k]]
    assert differ > 150


def test_oracles_follow_the_reference_on_bfv_crt_ring():
    qs, qb = bfv_crt_ring()
    rng = np.random.default_rng(6)
    n = 256
    big = _random_tensor(rng, qb, n)
    for k, X in enumerate(_edges(qs, qb, T_CRT)):
        for j, p in enumerate(qb):
            big[0, j, k] = X % p
    want = _want(big, qs, qb, T_CRT)
    # Python big-int oracle
    got_py = O.bfv_mul_contract([[[int(v) for v in big[0, j]] for j in range(len(qb))]], qs, qb, T_CRT)
    assert [[int(v) for v in row] for row in want[0]] == got_py[0]
    # C oracle
    assert np.array_equal(CO.bfv_contract(n, qs, qb, T_CRT, big), want)


@pytest.mark.parametrize("t", [2, 53, 65537, (1 << 40) + 15])
def test_oracles_follow_the_reference_other_t(t):
    qs, qb = bfv_crt_ring()
    rng = np.random.default_rng(t % 1000)
    big = _random_tensor(rng, qb, 64)
    want = _want(big, qs, qb, t)
    assert np.array_equal(CO.bfv_contract(64, qs, qb, t, big), want)


@pytest.mark.gpu
@pytest.mark.parametrize("generic", [0, 1])
def test_gpu_contract_follows_the_reference_on_bfv_crt_ring(generic):
    """GPU = C oracle = Python oracle = inline Julia transcription on test/bfv_crt.jl's ring with uniformly random
    tensor values (fails on the round-1 kernels, which multiplied by t over the integers)."""
    import toyfhe_b200 as T
    qs, qb = bfv_crt_ring()
    psis = [T.minimal_primitive_root(q, 2 * N_CRT) for q in qs]
    psib = [T.minimal_primitive_root(q, 2 * N_CRT) for q in qb]
    cq, cb = T.Context(N_CRT, qs, psis), T.Context(N_CRT, qb, psib)
    rng = np.random.default_rng(7)
    big = np.empty((3, len(qb), N_CRT), dtype=np.uint64)
    for j, p in enumerate(qb):
        big[:, j] = rng.integers(0, p, size=(3, N_CRT), dtype=np.uint64)
    for k, X in enumerate(_edges(qs, qb, T_CRT)):
        for j, p in enumerate(qb):
            big[0, j, k] = X % p
    T.force_generic(generic)
    try:
        got = cq.to_host(cq.bfv_contract(cb, T_CRT, cb.to_device(big)))
    finally:
        T.force_generic(0)
    assert np.array_equal(got, CO.bfv_contract(N_CRT, qs, qb, T_CRT, big))
    want = _want(big[:1, :, :300], qs, qb, T_CRT)
    assert np.array_equal(got[:1, :, :300], want)


@pytest.mark.gpu
def test_gpu_bfv_mul_on_bfv_crt_ring_wraps_like_the_reference():
    """Whole multiply on the 2+4-prime ring: Q_big (~2^200) < t N Q^2 (~2^217), so the engine must use the caller's
    basis (no joint-basis shortcut) and wrap modulo Q_big exactly as the reference does."""
    import toyfhe_b200 as T
    qs, qb = bfv_crt_ring()
    psis = [T.minimal_primitive_root(q, 2 * N_CRT) for q in qs]
    psib = [T.minimal_primitive_root(q, 2 * N_CRT) for q in qb]
    cq, cb = T.Context(N_CRT, qs, psis), T.Context(N_CRT, qb, psib)
    oq, ob = CO.Rns(N_CRT, qs, psis), CO.Rns(N_CRT, qb, psib)
    rng = np.random.default_rng(8)
    c1 = np.empty((1, 2, 2, N_CRT), dtype=np.uint64)
    c2 = np.empty((1, 2, 2, N_CRT), dtype=np.uint64)
    for i, q in enumerate(qs):
        c1[:, :, i] = rng.integers(0, q, size=(1, 2, N_CRT), dtype=np.uint64)
        c2[:, :, i] = rng.integers(0, q, size=(1, 2, N_CRT), dtype=np.uint64)
    got = cq.to_host(cq.bfv_mul(cb, T_CRT, cq.to_device(c1), cq.to_device(c2)))
    want = CO.bfv_mul(oq, ob, T_CRT, c1, c2)
    assert np.array_equal(got, want)
    # and the tensor -> contract tail against the transcription on a slice
    e1, e2 = CO.bfv_switch(N_CRT, qs, qb, c1), CO.bfv_switch(N_CRT, qs, qb, c2)
    tz = ob.ct_tensor(e1, e2)
    assert np.array_equal(got[0, :, :, :64], _want(tz[0, :, :, :64], qs, qb, T_CRT))
