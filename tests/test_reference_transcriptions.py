"""Statement-by-statement transcriptions of the reference's Julia for the hot-path operations whose parity the oracle
carries (round-1 review: "two restatements cross-check each other, which catches slips but not shared misreadings").
Every function below names the Julia lines it stands for and shares NO code with ``oracle/``; ring products are taken by
the O(N^2) definition of the quotient ring (the naive loop of pow2_cyc_rings.jl:157-164), never through a transform.
Each test asserts  transcription == Python oracle == C oracle  on the CPU, and ``-m gpu`` adds  == CUDA path.

mul_contract / multround / switchel have their own file (tests/test_contract_semantics.py)."""
import math

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import toyfhe_oracle as O


# ------------------------------------------------------------------ Julia, transcribed
def jl_ring_multiply(a, b, q):
    """pow2_cyc_rings.jl:157-164 (the psi == 0 branch: the definition of * in F_q[x]/(x^N+1)):
    for i, j: c[(i+j) % N] += (i+j >= N ? -1 : 1) * a[i] * b[j]"""
    N = len(a)
    c = [0] * N
    for i in range(N):
        if a[i] == 0:
            continue
        for j in range(N):
            k = i + j
            if k >= N:
                c[k - N] = (c[k - N] - a[i] * b[j]) % q
            else:
                c[k] = (c[k] + a[i] * b[j]) % q
    return c


def jl_crt_integer(res, qs):
    """convert(Integer, ::CRTEncoded) crt.jl:105-112 (unique value in [0, Q))"""
    x, m = int(res[0]), int(qs[0])
    for r, q in zip(res[1:], qs[1:]):
        x, m = (x + m * (((int(r) - x) * pow(m, -1, int(q))) % int(q))), m * int(q)
    return x


def jl_modswitch(poly, qs):
    """modswitch(crt::CRTEncoded) crt.jl:215-220 mapped over the primal coefficients (:226-228):
    ct_qk = crt.c[end];  cc -> inv(typeof(cc)(modulus(ct_qk))) * (cc - typeof(cc)(convert(Integer, ct_qk)))"""
    qk = qs[-1]
    out = []
    for cc_row, q in zip(poly[:-1], qs[:-1]):
        inv = pow(qk % q, -1, q)                       # inv(typeof(cc)(modulus(ct_qk)))
        out.append([(inv * ((int(cc) - int(ck) % q) % q)) % q for cc, ck in zip(cc_row, poly[-1])])
    return out


def jl_crt_expand(poly, qs, P):
    """c .* CRTExpand{P}() crt.jl:35-40: (convert(Integer, b) * a).c..., zero(T)"""
    return [[(P * int(c)) % q for c in row] for row, q in zip(poly, qs)] + [[0] * len(poly[0])]


def jl_apply_galois_element(a, g, q):
    """pow2_cyc_rings.jl:321-329: q_, r = divrem(g*i, N); output[r] = (q_ % 2 == 1) ? -val : val"""
    N = len(a)
    out = [0] * N
    for i in range(N):
        d, r = divmod(g * i, N)
        out[r] = (-int(a[i])) % q if d % 2 == 1 else int(a[i])
    return out


def jl_enc_mul(c1, c2, qs):
    """enc_mul without hooks, rlwe_she.jl:255-258: c = [zero ...]; for i, j: c[i+j-1] += c1[i] * c2[j]"""
    N = len(c1[0][0])
    c = [[[0] * N for _ in qs] for _ in range(len(c1) + len(c2) - 1)]
    for i in range(len(c1)):
        for j in range(len(c2)):
            for p, q in enumerate(qs):
                prod = jl_ring_multiply([int(v) for v in c1[i][p]], [int(v) for v in c2[j][p]], q)
                c[i + j][p] = [(x + y) % q for x, y in zip(c[i + j][p], prod)]
    return c


def jl_keyswitch(ct, key, qs, relin_window, raised_qs=None):
    """keyswitch(ek, c) rlwe_she.jl:315-347.  key[i] = (mask, masked) primal polynomials over the key ring.
    raised_qs = (qs..., special): ModulusRaised -- keyswitch_expand = c .* CRTExpand (modulusraising.jl:35-41),
    downswitch_keyelement selects rows [1:l; special] (:43-49, done by the caller), keyswitch_contract = modswitch (:42)."""
    ring_qs = list(raised_qs) if raised_qs is not None else list(qs)
    N = len(ct[0][0])
    expand = (lambda c: jl_crt_expand(c, qs, ring_qs[-1])) if raised_qs is not None else (lambda c: [list(map(int, r)) for r in c])
    c1 = expand(ct[0])                                                        # :323
    c2 = [[0] * N for _ in ring_qs] if len(ct) == 2 else expand(ct[1])        # :324
    cend = ct[-1]
    if relin_window == 0:
        # :326-329  per prime: convert.(Integer, SignedMod.(residue)) re-embedded in every prime of typeof(c1)
        ps = []
        for i, q in enumerate(qs):
            lifted = [int(x) - q if int(x) > q // 2 else int(x) for x in cend[i]]     # signedmod.jl:12-19
            ps.append([[v % p for v in lifted] for p in ring_qs])
    else:
        # :331-338  digits(convert(Integer, x), base = 2^w, pad = nwindows)
        Q = math.prod(qs)
        base = 2 ** relin_window
        nwindows, t = 0, Q
        while t > 0:                                                           # ndigits(Q, base = 2^w)
            t //= base
            nwindows += 1
        ints = [jl_crt_integer([cend[i][n] for i in range(len(qs))], qs) for n in range(N)]
        ps = []
        for k in range(nwindows):
            dig = [(x // base ** k) % base for x in ints]
            ps.append([[d % p for d in dig] for p in ring_qs])
    for i, p in enumerate(ps):                                                 # :340-344
        mask, masked = key[i]
        for r, q in enumerate(ring_qs):
            m2 = jl_ring_multiply([int(v) for v in mask[r]], p[r], q)
            m1 = jl_ring_multiply([int(v) for v in masked[r]], p[r], q)
            c2[r] = [(x + y) % q for x, y in zip(c2[r], m2)]
            c1[r] = [(x + y) % q for x, y in zip(c1[r], m1)]
    if raised_qs is not None:
        return [jl_modswitch(c1, ring_qs), jl_modswitch(c2, ring_qs)]
    return [c1, c2]


# ------------------------------------------------------------------ helpers
def _rnd(rng, qs, N, shape=()):
    a = np.empty(shape + (len(qs), N), dtype=np.uint64)
    for i, q in enumerate(qs):
        a[..., i, :] = rng.integers(0, q, size=shape + (N,), dtype=np.uint64)
    return a


def _ints(a):
    return [[int(v) for v in row] for row in a]


N = 16


def _chain(logqs):
    return O.prime_chain(N, logqs)


# ------------------------------------------------------------------ CPU: transcription == oracles
def test_modswitch_crtexpand_galois():
    qs, psis = _chain([60, 40, 40, 60])
    rng = np.random.default_rng(1)
    a = _rnd(rng, qs, N)
    a[:, 0] = [q - 1 for q in qs]                         # c_L = q_L - 1: the un-centred last residue matters
    assert O.modswitch(_ints(a), qs) == jl_modswitch(_ints(a), qs)
    P = O.nextprime(qs[-1] + 2 * N, 2 * N)
    assert O.rns_crt_expand(_ints(a), qs, P) == jl_crt_expand(_ints(a), qs, P)
    for g in (3, 2 * N - 1, 5, pow(3, 2 * N - 4, 2 * N)):
        for i, q in enumerate(qs):
            assert O.apply_galois_element([int(v) for v in a[i]], g, q) == jl_apply_galois_element(a[i], g, q)


def test_ring_product_and_tensor():
    qs, psis = _chain([60, 40])
    rng = np.random.default_rng(2)
    c1, c2 = _rnd(rng, qs, N, (2,)), _rnd(rng, qs, N, (2,))
    want = jl_enc_mul([_ints(x) for x in c1], [_ints(x) for x in c2], qs)
    got_py = O.ct_tensor([_ints(x) for x in c1], [_ints(x) for x in c2], qs, psis)
    assert got_py == want
    got_c = CO.Rns(N, qs, psis).ct_tensor(c1[None], c2[None])[0]
    assert [_ints(x) for x in got_c] == want


@pytest.mark.parametrize("w,comps", [(1, 3), (7, 2), (2, 3)])
def test_keyswitch_base_2w(w, comps):
    qs, psis = _chain([50, 50, 40])
    rng = np.random.default_rng(10 + w)
    D = CO.ndigits(qs, w)
    key = _rnd(rng, qs, N, (D, 2))
    ct = _rnd(rng, qs, N, (comps,))
    want = jl_keyswitch([_ints(x) for x in ct], [(_ints(k[0]), _ints(k[1])) for k in key], qs, w)
    orc = CO.Rns(N, qs, psis)
    c1 = ct[0]
    c2 = ct[1] if comps == 3 else np.zeros_like(ct[0])
    w1, w2 = orc.keyswitch_accum(orc.keyswitch_digits(ct[-1], w), key, c1, c2)
    assert [_ints(w1), _ints(w2)] == want
    got_py = O.keyswitch([_ints(x) for x in ct], [(_ints(k[0]), _ints(k[1])) for k in key], qs, psis, w)
    assert [list(map(list, got_py[0])), list(map(list, got_py[1]))] == want


# ------------------------------------------------------------------ GPU: CUDA path == transcription
@pytest.mark.gpu
def test_gpu_matches_transcriptions():
    import toyfhe_b200 as T
    H = T.Context.to_host
    # rescale, expand, Galois, tensor
    qs, psis = T.prime_chain(N, [60, 40, 40, 60])
    ctx = T.Context(N, qs, psis)
    rng = np.random.default_rng(3)
    a = _rnd(rng, qs, N, (2,))
    a[0, :, 0] = [q - 1 for q in qs]
    got = H(ctx.rescale(ctx.to_device(a)))
    for p in range(2):
        assert _ints(got[p]) == jl_modswitch(_ints(a[p]), qs)
    P = O.nextprime(qs[-1] + 2 * N, 2 * N)
    got = H(ctx.crt_expand(ctx.to_device(a), P))
    assert _ints(got[0]) == jl_crt_expand(_ints(a[0]), qs, P)
    for g in (3, 2 * N - 1, pow(3, 2 * N - 4, 2 * N)):
        got = H(ctx.galois(ctx.to_device(a), g))
        for i, q in enumerate(qs):
            assert [int(v) for v in got[1, i]] == jl_apply_galois_element(a[1, i], g, q)
    c1, c2 = _rnd(rng, qs, N, (1, 2)), _rnd(rng, qs, N, (1, 2))
    got = H(ctx.ct_tensor(ctx.to_device(c1), ctx.to_device(c2)))[0]
    assert [_ints(x) for x in got] == jl_enc_mul([_ints(x) for x in c1[0]], [_ints(x) for x in c2[0]], qs)


@pytest.mark.gpu
@pytest.mark.parametrize("w,comps,raised", [(1, 3, False), (2, 2, False), (0, 2, True), (0, 3, True), (0, 3, False)])
def test_gpu_keyswitch_matches_transcription(w, comps, raised):
    import toyfhe_b200 as T
    H = T.Context.to_host
    key_qs, key_psis = T.prime_chain(N, [50, 50, 40] + ([60] if raised else []))
    qs, psis = (key_qs[:-1], key_psis[:-1]) if raised else (key_qs, key_psis)
    ctx, kctx = T.Context(N, qs, psis), T.Context(N, key_qs, key_psis)
    rng = np.random.default_rng(20 + w + comps)
    D = len(key_qs) if w == 0 else T.ndigits(qs, w)
    key = _rnd(rng, key_qs, N, (D, 2))                                   # primal key components over the key ring
    ct = _rnd(rng, qs, N, (1, comps))
    ct[0, -1, :, 0] = [q // 2 + 1 for q in qs]                           # centred-lift boundary of the CRT digits
    ct[0, -1, :, 1] = [q // 2 for q in qs]
    want = jl_keyswitch([_ints(x) for x in ct[0]], [(_ints(k[0]), _ints(k[1])) for k in key], qs, w,
                        raised_qs=key_qs if raised else None)
    got = H(ctx.keyswitch(kctx.ntt_fwd(kctx.to_device(key)), ctx.to_device(ct), w, ext=kctx if raised else None))[0]
    assert [_ints(got[0]), _ints(got[1])] == want
