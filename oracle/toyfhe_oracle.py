"""CPU oracle (TEST INFRASTRUCTURE ONLY) -- Python big-int restatement of the
ToyFHE.jl power-of-two-cyclotomic / RNS hot path.

This file is the *checker*, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  Nothing under ``toyfhe.jl_b200/`` imports it.

Every function cites the reference file:line it restates (paths relative to
/root/reference).  The reference's inner arithmetic lives in un-vendored Julia
packages (FourierTransforms.jl @ bed6810d, GaloisFields.jl 0.4.0 @ ec946bde,
Manifest.toml:219-241) and Julia is not installed here, so the reference itself
cannot be run.  Pinning status:

  * primal-domain results (ring products, tensor, rescale, keyswitch, BFV mul)
    are exact field/integer arithmetic, unique by construction, and are pinned
    against every literal KAT the reference holds (tests/test_oracle_kats.py:
    docs/src/man/background/rlwe.md:183-212, docs/src/man/encoding.md:14-24 and
    :69-92, src/crt.jl:23-33 and :50-58, src/cryptparams.jl:22-25).
  * NTT-domain ("dual") *ordering*: parity unpinned -- no reference test or doc
    asserts a forward-NTT output vector.  We follow the docstring definition
    pow2_cyc_rings.jl:279-303 (nntt = NTT(PowMul_psi(a)), natural order,
    slot k <-> evaluation at psi^(2k+1)).

All values are Python ints; polynomials are lists; an RNS polynomial is a list
of L rows (residue-major, the StructArray layout of crt.jl:150-156).
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

# --------------------------------------------------------------------------
# number theory helpers (Primes.jl / GaloisFields.jl call sites)
# --------------------------------------------------------------------------

_MR_BASES = (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37)


def is_prime(n: int) -> bool:
    """Deterministic Miller-Rabin for n < 3.3e24 (Primes.isprime stand-in)."""
    if n < 2:
        return False
    for p in _MR_BASES:
        if n % p == 0:
            return n == p
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in _MR_BASES:
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def nextprime(n: int, interval: int = 1) -> int:
    """Primes.nextprime(n; interval): first prime in n, n+interval, ...
    (call sites crt.jl:287, test/bfv_crt.jl:9-10, infer.jl:97-105)."""
    while not is_prime(n):
        n += interval
    return n


def minimal_primitive_root(q: int, n: int) -> int:
    """GaloisFields.minimal_primitive_root(F_q, n) for n a power of two: the
    smallest integer in [1,q) of multiplicative order exactly n
    (pow2_cyc_rings.jl:40, crt.jl:142-144; doc KAT rlwe.md:183-187 -> 33)."""
    assert n >= 2 and n & (n - 1) == 0 and (q - 1) % n == 0
    e = (q - 1) // n
    a = 2
    while True:
        r = pow(a, e, q)
        if pow(r, n // 2, q) == q - 1:
            break
        a += 1
    # all primitive n-th roots are r^odd
    best = r
    r2 = r * r % q
    x = r
    for _ in range(n // 2 - 1):
        x = x * r2 % q
        if x < best:
            best = x
    return best


def prime_chain(N: int, logqs: Sequence[int]) -> Tuple[List[int], List[int]]:
    """NegacyclicRing(N, logqs) (crt.jl:282-295): returns (primes, psis) in the
    order of ``logqs``; primes generated in ascending-logq order, each
    ``nextprime(max(2^logq+1, last+2N); interval=2N)``; psi_i = minimal
    primitive 2N-th root mod p_i."""
    perm = sorted(range(len(logqs)), key=lambda i: logqs[i])  # sortperm is stable
    primes = [0] * len(logqs)
    last = 0
    for i in perm:
        p = nextprime(max(2 ** logqs[i] + 1, last + 2 * N), 2 * N)
        last = p
        primes[i] = p
    psis = [minimal_primitive_root(p, 2 * N) for p in primes]
    return primes, psis


def centre(x: int, m: int) -> int:
    """SignedMod lift (signedmod.jl:12-19): x > m div 2 ? x-m : x."""
    return x - m if x > m // 2 else x


def rha(a: int, b: int) -> int:
    """div(a, b, RoundNearestTiesAway) (div_hacks.jl:120-135 via bfv.jl:172-174,196)."""
    assert b > 0
    q, r = divmod(abs(a), b)
    if 2 * r >= b:
        q += 1
    return q if a >= 0 else -q


# --------------------------------------------------------------------------
# single-prime negacyclic NTT (pow2_cyc_rings.jl:279-318)
# --------------------------------------------------------------------------

def nntt_def(c: Sequence[int], q: int, psi: int) -> List[int]:
    """O(N^2) *definition*: c^[k] = sum_j c[j] psi^(j(2k+1)) (pow2_cyc_rings.jl:295-303)."""
    N = len(c)
    return [sum(c[j] * pow(psi, j * (2 * k + 1), q) for j in range(N)) % q for k in range(N)]


def _bitrev_permute(a: List[int]) -> List[int]:
    n = len(a)
    lg = n.bit_length() - 1
    out = [0] * n
    for i in range(n):
        out[int(format(i, "0%db" % lg)[::-1], 2) if lg else 0] = a[i]
    return out


def _cyclic_ntt(a: List[int], q: int, w: int) -> List[int]:
    """Natural-order cyclic DFT over F_q with root w (what FourierTransforms'
    CTPlan computes, pow2_cyc_rings.jl:301,315): iterative radix-2 DIT."""
    n = len(a)
    a = _bitrev_permute(list(a))
    m = 2
    while m <= n:
        wm = pow(w, n // m, q)
        half = m // 2
        for s in range(0, n, m):
            t = 1
            for j in range(half):
                u = a[s + j]
                v = a[s + j + half] * t % q
                a[s + j] = (u + v) % q
                a[s + j + half] = (u - v) % q
                t = t * wm % q
        m *= 2
    return a


def nntt(c: Sequence[int], q: int, psi: int) -> List[int]:
    """nntt(p) = NTT(PowMul_psi(p)) (pow2_cyc_rings.jl:295-303)."""
    N = len(c)
    pm = [0] * N
    t = 1
    for i in range(N):
        pm[i] = c[i] * t % q
        t = t * psi % q
    return _cyclic_ntt(pm, q, psi * psi % q)


def inntt(ch: Sequence[int], q: int, psi: int) -> List[int]:
    """inverse of nntt: c[i] = N^-1 psi^-i sum_k c^[k] w^-ik (pow2_cyc_rings.jl:308-318)."""
    N = len(ch)
    ipsi = pow(psi, q - 2, q)
    a = _cyclic_ntt(list(ch), q, ipsi * ipsi % q)
    ninv = pow(N, q - 2, q)
    out = [0] * N
    t = ninv
    for i in range(N):
        out[i] = a[i] * t % q
        t = t * ipsi % q
    return out


def ring_multiply_naive(a: Sequence[int], b: Sequence[int], q: int) -> List[int]:
    """psi==0 branch of ring_multiply (pow2_cyc_rings.jl:157-164)."""
    N = len(a)
    res = [0] * N
    for i in range(N):
        if a[i] == 0:
            continue
        for j in range(N):
            k = i + j
            if k < N:
                res[k] = (res[k] + a[i] * b[j]) % q
            else:
                res[k - N] = (res[k - N] - a[i] * b[j]) % q
    return res


def ring_multiply(a: Sequence[int], b: Sequence[int], q: int, psi: int) -> List[int]:
    """dual(a) .* dual(b), read back in primal (pow2_cyc_rings.jl:147-169,124-130)."""
    A, B = nntt(a, q, psi), nntt(b, q, psi)
    return inntt([x * y % q for x, y in zip(A, B)], q, psi)


def apply_galois_element(a: Sequence[int], g: int, q: int) -> List[int]:
    """out[(g i) mod N] = floor(g i / N) odd ? -a[i] : a[i] (pow2_cyc_rings.jl:321-329)."""
    N = len(a)
    out = [0] * N
    for i in range(N):
        qq, r = divmod(g * i, N)
        out[r] = (-a[i]) % q if qq % 2 == 1 else a[i]
    return out


def galois_element_from_steps(steps: int, N: int) -> int:
    """rlwe_she.jl:304: steps>0 ? 3^(2N-steps) : 3^(-steps), mod 2N."""
    return pow(3, 2 * N - steps, 2 * N) if steps > 0 else pow(3, -steps, 2 * N)


# --------------------------------------------------------------------------
# RNS layer (crt.jl)
# --------------------------------------------------------------------------

def crt_encode(x: int, qs: Sequence[int]) -> List[int]:
    """CRTEncoded{N,M}(x::Integer) (crt.jl:91-95)."""
    return [x % q for q in qs]


def crt_reconstruct(res: Sequence[int], qs: Sequence[int]) -> int:
    """convert(Integer, ::CRTEncoded): unique X in [0,Q) (crt.jl:98-112)."""
    X, M = 0, 1
    for r, q in zip(res, qs):
        # combine X (mod M) with r (mod q)
        t = ((r - X) * pow(M, -1, q)) % q
        X += M * t
        M *= q
    return X


def crt_expand(res: Sequence[int], qs: Sequence[int], p: int) -> List[int]:
    """a * CRTExpand{p}: multiply by p, append residue 0 (crt.jl:35-40)."""
    return [(p * r) % q for r, q in zip(res, qs)] + [0]


def crt_residual(c: int, i: int, qs: Sequence[int]) -> List[int]:
    """CRTEncoded(CRTResidual(c)): residue c at prime i, 0 elsewhere (crt.jl:60-77)."""
    return [c if j == i else 0 for j in range(len(qs))]


def rns_nntt(poly: Sequence[Sequence[int]], qs, psis) -> List[List[int]]:
    """per-prime nntt over StructArray fields (crt.jl:247-256)."""
    return [nntt(r, q, s) for r, q, s in zip(poly, qs, psis)]


def rns_inntt(poly, qs, psis) -> List[List[int]]:
    """crt.jl:258-267."""
    return [inntt(r, q, s) for r, q, s in zip(poly, qs, psis)]


def rns_ring_multiply(a, b, qs, psis) -> List[List[int]]:
    return [ring_multiply(x, y, q, s) for x, y, q, s in zip(a, b, qs, psis)]


def rns_add(a, b, qs):
    return [[(x + y) % q for x, y in zip(r1, r2)] for r1, r2, q in zip(a, b, qs)]


def rns_sub(a, b, qs):
    return [[(x - y) % q for x, y in zip(r1, r2)] for r1, r2, q in zip(a, b, qs)]


def rns_neg(a, qs):
    return [[(-x) % q for x in r] for r, q in zip(a, qs)]


def rns_scalar_mul(a, s: int, qs):
    """scalar_mul (pow2_cyc_rings.jl:177-185) with integer scalar."""
    return [[(x * s) % q for x in r] for r, q in zip(a, qs)]


def rns_from_ints(coeffs: Sequence[int], qs) -> List[List[int]]:
    return [[c % q for c in coeffs] for q in qs]


def rns_to_ints(poly, qs) -> List[int]:
    N = len(poly[0])
    return [crt_reconstruct([poly[i][n] for i in range(len(qs))], qs) for n in range(N)]


def rns_galois(a, g, qs):
    return [apply_galois_element(r, g, q) for r, q in zip(a, qs)]


def modswitch(poly, qs) -> List[List[int]]:
    """CKKS rescale / divide by last prime (crt.jl:215-220, 226-228):
    c'_i = (q_L mod q_i)^-1 (c_i - (c_L mod q_i)), c_L un-centred."""
    qL = qs[-1]
    last = poly[-1]
    out = []
    for r, q in zip(poly[:-1], qs[:-1]):
        inv = pow(qL % q, -1, q)
        out.append([(inv * (c - (cl % q))) % q for c, cl in zip(r, last)])
    return out


def modswitch_drop(poly):
    """crt.jl:222-224, 230-232."""
    return [list(r) for r in poly[:-1]]


def rns_crt_expand(poly, qs, p):
    """keyswitch_expand for ModulusRaised: c .* CRTExpand{p} (modulusraising.jl:35-41)."""
    N = len(poly[0])
    return [[(p * c) % q for c in r] for r, q in zip(poly, qs)] + [[0] * N]


# --------------------------------------------------------------------------
# scheme-layer bodies that the fused kernels compute (rlwe_she.jl, bfv.jl)
# --------------------------------------------------------------------------

def ct_tensor(c1, c2, qs, psis):
    """enc_mul without expand/contract (rlwe_she.jl:255-258): c[i+j-1] += c1[i]*c2[j]."""
    n1, n2 = len(c1), len(c2)
    N = len(c1[0][0])
    out = [[[0] * N for _ in qs] for _ in range(n1 + n2 - 1)]
    for i in range(n1):
        for j in range(n2):
            out[i + j] = rns_add(out[i + j], rns_ring_multiply(c1[i], c2[j], qs, psis), qs)
    return out


def bfv_switch(poly, qs_from, qs_to):
    """switch/switchel (bfv.jl:202-226): centred lift from Q, reduce into target basis."""
    Q = math.prod(qs_from)
    half = Q >> 1
    N = len(poly[0])
    out = [[0] * N for _ in qs_to]
    for n in range(N):
        en = crt_reconstruct([poly[i][n] for i in range(len(qs_from))], qs_from)
        if en > half:
            en -= Q
        for j, p in enumerate(qs_to):
            out[j][n] = en % p
    return out


def bfv_mul_expand(ct, qs, qs_big):
    """mul_expand (bfv.jl:34)."""
    return [bfv_switch(c, qs, qs_big) for c in ct]


def bfv_multround(poly, qs_big, t: int, Q: int):
    """multround(e, t, q) on an R_big element (bfv.jl:182-190 -> :172-174), per coefficient:

        multround(SignedMod(x), t, Q) = div(SignedMod(x) * t, Q, RoundNearestTiesAway).x

    ``SignedMod(x) * t`` is ``SignedMod{T}(x * T(t))`` (signedmod.jl:24-28): the product is taken in the
    CRT field, i.e. modulo Q_big, BEFORE the centred lift of ``div`` (signedmod.jl:12-19, 30-32);
    ``oftype(e, y)`` then re-encodes y modulo Q_big.  Equal to rha(t*centre(x), Q) only while
    t*|x| < Q_big/2 (false on test/bfv_crt.jl's 2+4-prime ring, where the reference wraps)."""
    Qb = math.prod(qs_big)
    N = len(poly[0])
    out = [[0] * N for _ in qs_big]
    for n in range(N):
        tx = [(poly[j][n] * (t % p)) % p for j, p in enumerate(qs_big)]       # e.x * T(t), residue-wise
        y = rha(centre(crt_reconstruct(tx, qs_big), Qb), Q)                     # div(convert(Integer, e), Q, RNTA)
        for j, p in enumerate(qs_big):
            out[j][n] = y % p                                                   # oftype(e, y)
    return out


def bfv_mul_contract(cs, qs, qs_big, t: int):
    """mul_contract (bfv.jl:35-40): switch(R, multround(e, t, Q))."""
    Q = math.prod(qs)
    return [bfv_switch(bfv_multround(c, qs_big, t, Q), qs_big, qs) for c in cs]


def bfv_mul(c1, c2, qs, psis, qs_big, psis_big, t: int):
    """BFV enc_mul (rlwe_she.jl:247-262 with bfv.jl:34-40)."""
    e1 = bfv_mul_expand(c1, qs, qs_big)
    e2 = bfv_mul_expand(c2, qs, qs_big)
    return bfv_mul_contract(ct_tensor(e1, e2, qs_big, psis_big), qs, qs_big, t)


def ndigits(x: int, base: int) -> int:
    n = 0
    while x > 0:
        x //= base
        n += 1
    return max(n, 1)


def keyswitch_digits(cend, qs, relin_window: int, target_qs=None):
    """Digit polys p_i of keyswitch (rlwe_she.jl:326-338), embedded in
    ``target_qs`` (default: qs; ModulusRaised passes the expanded basis)."""
    target_qs = list(qs) if target_qs is None else list(target_qs)
    N = len(cend[0])
    if relin_window == 0:
        # CRT digits: centred residue i re-embedded in every prime (:329)
        return [[[centre(c, qs[i]) % p for c in cend[i]] for p in target_qs] for i in range(len(qs))]
    Q = math.prod(qs)
    base = 2 ** relin_window
    nw = ndigits(Q, base)
    ints = rns_to_ints(cend, qs)
    ps = []
    for k in range(nw):
        dig = [(x >> (relin_window * k)) & (base - 1) for x in ints]
        ps.append([[d % p for d in dig] for p in target_qs])
    return ps


def keyswitch(ct, key, qs, psis, relin_window: int):
    """keyswitch (rlwe_she.jl:315-347), plain params (no modulus raising).
    ``key`` = list of (mask, masked) RNS polys in primal form."""
    assert len(ct) in (2, 3)
    N = len(ct[0][0])
    c1 = [list(r) for r in ct[0]]
    c2 = [[0] * N for _ in qs] if len(ct) == 2 else [list(r) for r in ct[1]]
    ps = keyswitch_digits(ct[-1], qs, relin_window)
    assert len(ps) <= len(key)
    l = len(qs)
    for p, (mask, masked) in zip(ps, key):
        # downswitch_keyelement (crt.jl:238-244): keep the first l residues
        mask, masked = mask[:l], masked[:l]
        c2 = rns_add(c2, rns_ring_multiply(mask, p, qs, psis), qs)
        c1 = rns_add(c1, rns_ring_multiply(masked, p, qs, psis), qs)
    return [c1, c2]


def keyswitch_modraised(ct, key, qs_key, psis_key, relin_window: int = 0):
    """keyswitch with ModulusRaised params (modulusraising.jl:35-49 +
    rlwe_she.jl:315-347).  ``qs_key`` = full key basis (special prime last);
    the ciphertext lives on the first l = len(ct[0]) primes.  Key components
    are over the full key basis; downswitch selects residues [1..l, special]."""
    l = len(ct[0])
    N = len(ct[0][0])
    P = qs_key[-1]
    which = list(range(l)) + [len(qs_key) - 1]
    qs_c = [qs_key[i] for i in range(l)]
    qs_e = [qs_key[i] for i in which]
    psis_e = [psis_key[i] for i in which]
    c1 = rns_crt_expand(ct[0], qs_c, P)
    c2 = [[0] * N for _ in qs_e] if len(ct) == 2 else rns_crt_expand(ct[1], qs_c, P)
    ps = keyswitch_digits(ct[-1], qs_c, relin_window, qs_e)
    for i, p in enumerate(ps):
        mask = [key[i][0][w] for w in which]
        masked = [key[i][1][w] for w in which]
        c2 = rns_add(c2, rns_ring_multiply(mask, p, qs_e, psis_e), qs_e)
        c1 = rns_add(c1, rns_ring_multiply(masked, p, qs_e, psis_e), qs_e)
    return [modswitch(c1, qs_e), modswitch(c2, qs_e)]


# --------------------------------------------------------------------------
# keygen / encrypt / decrypt (rlwe_she.jl:155-216) with an explicit sampler
# --------------------------------------------------------------------------

class Sampler:
    """Seeded sampler standing in for the reference's unseeded global RNG
    (rlwe_she.jl:169-170; Appendix B of SURVEY.md: parity is statistical)."""

    def __init__(self, seed: int):
        import numpy as np
        self.rng = np.random.Generator(np.random.PCG64(seed))

    def uniform(self, N, qs):
        # rand(::CRTEncoded) draws every residue independently (crt.jl:146-148)
        return [[int(x) for x in self.rng.integers(0, q, size=N, dtype="uint64")] for q in qs]

    def gaussian_ints(self, N, sigma):
        import numpy as np
        return [int(x) for x in np.rint(self.rng.normal(0.0, sigma, size=N))]

    def gaussian(self, N, qs, sigma):
        return rns_from_ints(self.gaussian_ints(N, sigma), qs)


def keygen(s: Sampler, N, qs, psis, sigma):
    """rlwe_she.jl:155-167: masked = -(mask*secret + error)."""
    mask = s.uniform(N, qs)
    secret = s.gaussian(N, qs, sigma)
    err = s.gaussian(N, qs, sigma)
    masked = rns_neg(rns_add(rns_ring_multiply(mask, secret, qs, psis), err, qs), qs)
    return secret, (mask, masked)


def encrypt_zero(s: Sampler, pub, N, qs, psis, sigma):
    """rlwe_she.jl:176-186."""
    mask, masked = pub
    u = s.gaussian(N, qs, sigma)
    e1 = s.gaussian(N, qs, sigma)
    e2 = s.gaussian(N, qs, sigma)
    c1 = rns_add(rns_ring_multiply(masked, u, qs, psis), e1, qs)
    c2 = rns_add(rns_ring_multiply(mask, u, qs, psis), e2, qs)
    return [c1, c2]


def encrypt(s: Sampler, pub, plain_rns, N, qs, psis, sigma):
    """rlwe_she.jl:188-195 (plain_rns = pi^-1(plaintext) already in R_cipher)."""
    c = encrypt_zero(s, pub, N, qs, psis, sigma)
    c[0] = rns_add(c[0], plain_rns, qs)
    return c


def decrypt_raw(secret, ct, qs, psis):
    """b = c[1] + sum s^i c[i+1] (rlwe_she.jl:199-213), before pi."""
    b = [list(r) for r in ct[0]]
    spow = secret
    for i in range(1, len(ct)):
        b = rns_add(b, rns_ring_multiply(spow, ct[i], qs, psis), qs)
        spow = rns_ring_multiply(spow, secret, qs, psis)
    return b


def bfv_pi_inv(plain: Sequence[int], t: int, qs):
    """pi^-1 = Delta * plaintext, Delta = Q div t (bfv.jl:21-24, test/bfv_crt.jl:35)."""
    Q = math.prod(qs)
    delta = Q // t
    return rns_from_ints([delta * (m % t) for m in plain], qs)


def bfv_pi(b, t: int, qs) -> List[int]:
    """pi (bfv.jl:26-29): mod(divround(centre(x), Delta), t)."""
    Q = math.prod(qs)
    delta = Q // t
    return [rha(centre(x, Q), delta) % t for x in rns_to_ints(b, qs)]


def make_eval_key(s: Sampler, old, secret, N, qs, psis, sigma, relin_window: int, raise_by: int = 1):
    """make_eval_key (rlwe_she.jl:273-298); ``raise_by`` = P for ModulusRaised
    (modulusraising.jl:28-32: old <- P*old)."""
    if raise_by != 1:
        old = rns_scalar_mul(old, raise_by, qs)
    if relin_window != 0:
        Q = math.prod(qs)
        nw = ndigits(Q, 2 ** relin_window)
        evala = [rns_scalar_mul(old, pow(2, i * relin_window), qs) for i in range(nw)]
    else:
        old_p = old
        evala = [[(old_p[j] if j == i else [0] * N) for j in range(len(qs))] for i in range(len(qs))]
    key = []
    for a in evala:
        mask = s.uniform(N, qs)
        e = s.gaussian(N, qs, sigma)
        masked = rns_sub(a, rns_add(rns_ring_multiply(mask, secret, qs, psis), e, qs), qs)
        key.append((mask, masked))
    return key
