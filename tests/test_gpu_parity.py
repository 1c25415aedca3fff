"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle on
the same seeded inputs -- bit-exact (all arithmetic is integer).  Run with
``pytest -m gpu`` on a B200."""
import math

import numpy as np
import pytest

import toyfhe_b200 as T
from oracle import c_oracle as CO
from oracle import toyfhe_oracle as O

pytestmark = pytest.mark.gpu


def _rand(rng, N, qs, shape=()):
    out = np.empty(shape + (len(qs), N), dtype=np.uint64)
    for i, q in enumerate(qs):
        out[..., i, :] = rng.integers(0, q, size=shape + (N,), dtype=np.uint64)
    return out


def _ring(N, logqs):
    qs, psis = T.prime_chain(N, logqs)
    return qs, psis, T.Context(N, qs, psis), CO.Rns(N, qs, psis)


H = T.Context.to_host


# ---------------------------------------------------------------- doc KATs on the GPU
def test_doc_kats_on_gpu():
    # docs/src/man/background/rlwe.md:183-212 : Z_97[x]/(x^4+1), psi = 33
    ctx = T.Context(4, [97], [33])
    d = lambda v: ctx.to_device(np.array([v], dtype=np.uint64))
    p1, p2, p3, p4 = d([1, 1, 0, 0]), d([0, 0, 0, 1]), d([4, 0, 0, 0]), d([5, 0, 0, 0])
    assert H(ctx.ring_mul(p3, p4)).tolist() == [[20, 0, 0, 0]]
    assert H(ctx.ring_mul(p1, p1)).tolist() == [[1, 2, 1, 0]]
    assert H(ctx.ring_mul(p1, p2)).tolist() == [[96, 0, 0, 1]]
    assert H(ctx.ntt_fwd(p1)).tolist() == [[34, 48, 65, 51]]
    assert H(ctx.galois(d([1, 2, 3, 4]), 3)).tolist() == [[1, 4, 94, 2]]
    # docs/src/man/encoding.md:69-92 : slot product at p = 65537, N = 2048
    N, q = 2048, 65537
    psi = T.minimal_primitive_root(q, 2 * N)
    c = T.Context(N, [q], [psi])
    a_slots = np.zeros((1, N), dtype=np.uint64); a_slots[0, :10] = np.arange(1, 11)
    b_slots = np.full((1, N), 10, dtype=np.uint64)
    a = c.ntt_inv(c.to_device(a_slots)); b = c.ntt_inv(c.to_device(b_slots))
    prod = c.ntt_fwd(c.ring_mul(a, b))
    assert H(prod)[0, :11].tolist() == [10, 20, 30, 40, 50, 60, 70, 80, 90, 100, 0]
    # src/cryptparams.jl:22-25 PALISADE rings: product against the naive O(N^2) definition
    for (q, N, psi) in [(1099511627873, 8, 108163207722), (525313, 512, 513496)]:
        c = T.Context(N, [q], [psi])
        rng = np.random.default_rng(N)
        x = rng.integers(0, q, size=(1, N), dtype=np.uint64); y = rng.integers(0, q, size=(1, N), dtype=np.uint64)
        want = O.ring_multiply_naive([int(v) for v in x[0]], [int(v) for v in y[0]], q)
        assert [int(v) for v in H(c.ring_mul(c.to_device(x), c.to_device(y)))[0]] == want


# ---------------------------------------------------------------------- transforms
@pytest.mark.parametrize("logN", [1, 2, 4, 5, 9, 10, 11, 12, 13, 14, 15, 16])
def test_ntt_parity_all_sizes(logN):
    N = 1 << logN
    qs, psis, ctx, orc = _ring(N, [60, 40, 60] if logN >= 4 else [40, 50, 60])
    rng = np.random.default_rng(logN)
    B = 3 if logN >= 15 else 5
    a = _rand(rng, N, qs, (B,))
    # edge rows: zeros, all q-1, single one
    a[0, 0, :] = 0
    a[0, 1, :] = qs[1] - 1
    a[0, 2, :] = 0; a[0, 2, N - 1] = 1
    d = ctx.to_device(a)
    f = ctx.ntt_fwd(d)
    want = orc.nntt(a)
    assert np.array_equal(H(f), want)
    assert np.array_equal(H(ctx.ntt_inv(f)), a)
    assert np.array_equal(H(ctx.ntt_inv(ctx.to_device(a))), orc.inntt(a))
    # in place
    d2 = d.clone()
    ctx.ntt_fwd(d2, out=d2)
    assert np.array_equal(H(d2), want)
    ctx.ntt_inv(d2, out=d2)
    assert np.array_equal(H(d2), a)


def test_ntt_headline_config_full_batch(q8, psi8):
    """N = 2^14, L = 8 (BASELINE configs[1]): parity on two polys against the oracle,
    then size-independent properties on a batch larger than L2."""
    N = 2 ** 14
    ctx, orc = T.Context(N, q8, psi8), CO.Rns(N, q8, psi8)
    rng = np.random.default_rng(14)
    a = _rand(rng, N, q8, (2,))
    assert np.array_equal(H(ctx.ntt_fwd(ctx.to_device(a))), orc.nntt(a))
    B = 160  # 160 MiB of residues
    big = ctx.to_device(_rand(rng, N, q8, (B,)))
    f = ctx.ntt_fwd(big)
    assert bool((ctx.ntt_inv(f) == big).all())
    # linearity: NTT(x + y) = NTT(x) + NTT(y)
    y = ctx.to_device(_rand(rng, N, q8, (B,)))
    lhs = ctx.ntt_fwd(ctx.add(big, y))
    rhs = ctx.add(f, ctx.ntt_fwd(y))
    assert bool((lhs == rhs).all())
    # every output canonical
    qcol = ctx.to_device(np.array(q8, dtype=np.uint64)).view(1, 8, 1)
    assert bool(((f >= 0) & (f < qcol)).all())  # int64 view is fine: q < 2^62


def test_elementwise_and_ring_ops():
    N = 1024
    qs, psis, ctx, orc = _ring(N, [60, 45, 40, 60])
    rng = np.random.default_rng(2)
    a, b = _rand(rng, N, qs, (3,)), _rand(rng, N, qs, (3,))
    a[0, :, :4] = 0
    b[0, :, :4] = 0
    da, db = ctx.to_device(a), ctx.to_device(b)
    assert np.array_equal(H(ctx.add(da, db)), orc.add(a, b))
    assert np.array_equal(H(ctx.sub(da, db)), orc.sub(a, b))
    assert np.array_equal(H(ctx.mul(da, db)), orc.mul(a, b))
    assert np.array_equal(H(ctx.neg(da)), orc.neg(a))
    s = 123456789012345678901234567890
    assert np.array_equal(H(ctx.scalar_mul(da, s)), orc.scalar_mul(a, s))
    assert np.array_equal(H(ctx.ring_mul(da, db)), orc.ring_mul(a, b))
    for g in (3, 5, 2 * N - 1, pow(3, 2 * N - 1, 2 * N)):
        assert np.array_equal(H(ctx.galois(da, g)), orc.galois(a, g))
    assert np.array_equal(H(ctx.rescale(da)), orc.modswitch(a))
    qs3, psis3 = qs[:3], psis[:3]
    c3, o3 = T.Context(N, qs3, psis3), CO.Rns(N, qs3, psis3)
    a3 = np.ascontiguousarray(a[:, :3, :])
    assert np.array_equal(H(c3.crt_expand(c3.to_device(a3), qs[3])), o3.crt_expand(a3, qs[3]))


def test_small_ring_ops_like_reference_ckks_tests():
    # test/ckks_modswitch.jl: N = 32, three 40-bit primes
    N = 32
    qs, psis, ctx, orc = _ring(N, [40, 40, 40])
    rng = np.random.default_rng(3)
    a, b = _rand(rng, N, qs, (2,)), _rand(rng, N, qs, (2,))
    da, db = ctx.to_device(a), ctx.to_device(b)
    assert np.array_equal(H(ctx.ring_mul(da, db)), orc.ring_mul(a, b))
    assert np.array_equal(H(ctx.rescale(da)), orc.modswitch(a))
    assert np.array_equal(H(ctx.galois(da, 3)), orc.galois(a, 3))


# ------------------------------------------------------------- ciphertext multiply
@pytest.mark.parametrize("N,logqs,B", [(32, [40, 40, 40], 2), (2048, [50, 50], 2), (2 ** 14, [60] * 8, 2)])
def test_ct_tensor(N, logqs, B):
    qs, psis, ctx, orc = _ring(N, logqs)
    rng = np.random.default_rng(N)
    c1, c2 = _rand(rng, N, qs, (B, 2)), _rand(rng, N, qs, (B, 2))
    got = H(ctx.ct_tensor(ctx.to_device(c1), ctx.to_device(c2)))
    assert np.array_equal(got, orc.ct_tensor(c1, c2))


@pytest.mark.parametrize("N,L,Lb,t,B", [(32, 2, 4, 53, 2), (2048, 2, 4, 53, 1), (1024, 8, 17, 65537, 1)])
def test_bfv_switch_contract_mul(N, L, Lb, t, B):
    allq, allpsi = T.prime_chain(N, [60 if L == 8 else 50] * (L + Lb))
    qs, psis, qb, psib = allq[:L], allpsi[:L], allq[L:], allpsi[L:]
    cq, cb = T.Context(N, qs, psis), T.Context(N, qb, psib)
    oq, ob = CO.Rns(N, qs, psis), CO.Rns(N, qb, psib)
    rng = np.random.default_rng(N + L)
    c1, c2 = _rand(rng, N, qs, (B, 2)), _rand(rng, N, qs, (B, 2))
    Q, Qb = math.prod(qs), math.prod(qb)
    # values at and around the centring boundary, including both sides of the window (|x/Q - 1/2| ~ 2L 2^-60) inside which
    # the fast expansion hands over to the exact Garner route
    edge = [Q >> 1, (Q >> 1) + 1, 0, Q - 1, 1, (Q >> 1) - 1]
    for sh in (54, 55, 56, 57, 58, 59, 61):
        edge += [(Q >> 1) + (Q >> sh), (Q >> 1) - (Q >> sh), (Q >> 1) + 1 + (Q >> sh)]
    for k, X in enumerate(edge):   # strict '>' rule, bfv.jl:202-220
        for i, q in enumerate(qs):
            c1[0, 0, i, k] = X % q
    e1 = H(cq.bfv_switch(cb, cq.to_device(c1)))
    assert np.array_equal(e1, CO.bfv_switch(N, qs, qb, c1))
    big = _rand(rng, N, qb, (3,))
    for k, X in enumerate([Qb >> 1, (Qb >> 1) + 1, 0, 1, Qb - 1, Q >> 1, (Q >> 1) + 1, Q, Q - 1]):
        for j, p in enumerate(qb):
            big[0, j, k] = X % p
    got = H(cq.bfv_contract(cb, t, cb.to_device(big)))
    assert np.array_equal(got, CO.bfv_contract(N, qs, qb, t, big))
    gm = H(cq.bfv_mul(cb, t, cq.to_device(c1), cq.to_device(c2)))
    assert np.array_equal(gm, CO.bfv_mul(oq, ob, t, c1, c2))


def test_bfv_mul_headline_config(q8, psi8):
    """BASELINE configs[1]: N = 2^14, L = 8, t = 65537, R_big = 17 further primes."""
    N = 2 ** 14
    allq, allpsi = T.prime_chain(N, [60] * 25)
    assert allq[:8] == q8
    qb, psib = allq[8:], allpsi[8:]
    cq, cb = T.Context(N, q8, psi8), T.Context(N, qb, psib)
    oq, ob = CO.Rns(N, q8, psi8), CO.Rns(N, qb, psib)
    rng = np.random.default_rng(99)
    c1, c2 = _rand(rng, N, q8, (1, 2)), _rand(rng, N, q8, (1, 2))
    gm = H(cq.bfv_mul(cb, 65537, cq.to_device(c1), cq.to_device(c2)))
    assert np.array_equal(gm, CO.bfv_mul(oq, ob, 65537, c1, c2))


@pytest.mark.parametrize("N,L,Lb,t", [(1024, 8, 17, 65537), (4096, 3, 7, 65537), (2048, 1, 3, 257), (64, 4, 9, 7), (256, 2, 6, 53)])
def test_bfv_mul_joint_basis_equals_callers_basis(N, L, Lb, t):
    """tfb_bfv_mul over its own extension basis Q u P' (first primes of R_big) vs the step-by-step
    expand -> tensor -> contract over the caller's R_big (generic kernels) vs the oracle: identical,
    including operands at the centring boundary (bfv.jl:202-220)."""
    allq, allpsi = T.prime_chain(N, [60] * (L + Lb))
    qs, psis, qb, psib = allq[:L], allpsi[:L], allq[L:], allpsi[L:]
    cq, cb = T.Context(N, qs, psis), T.Context(N, qb, psib)
    oq, ob = CO.Rns(N, qs, psis), CO.Rns(N, qb, psib)
    rng = np.random.default_rng(7 * N + L)
    c1, c2 = _rand(rng, N, qs, (3, 2)), _rand(rng, N, qs, (3, 2))
    Q = math.prod(qs)
    # values at and around the centring boundary, including both sides of the window (|x/Q - 1/2| ~ 2L 2^-60) inside which
    # the fast expansion hands over to the exact Garner route
    edge = [Q >> 1, (Q >> 1) + 1, 0, Q - 1, 1, (Q >> 1) - 1]
    for sh in (54, 55, 56, 57, 58, 59, 61):
        edge += [(Q >> 1) + (Q >> sh), (Q >> 1) - (Q >> sh), (Q >> 1) + 1 + (Q >> sh)]
    for k, X in enumerate(edge):
        for i, q in enumerate(qs):
            c1[0, 0, i, k] = X % q
            c2[1, 1, i, (k * 5) % N] = X % q
    c1[2], c2[2] = c1[0], c1[0]                       # extreme magnitudes multiplied together
    want = CO.bfv_mul(oq, ob, t, c1, c2)
    d1, d2 = cq.to_device(c1), cq.to_device(c2)
    fast = H(cq.bfv_mul(cb, t, d1, d2))
    T.force_generic(1)
    try:
        slow = H(cq.bfv_mul(cb, t, d1, d2))
        T.force_generic(2)                            # joint-basis kernels without the Solinas folds
        mid = H(cq.bfv_mul(cb, t, d1, d2))
    finally:
        T.force_generic(0)
    assert np.array_equal(fast, want)
    assert np.array_equal(slow, want)
    assert np.array_equal(mid, want)


def test_squaring_shortcut_equals_the_general_product():
    """c*c with one operand buffer (what `c*c` in the reference's tests and the x^2 activation of
    examples/encrypted_mnist pass): the operand is expanded / transformed once; same result as two distinct buffers"""
    N, L, Lb, t = 4096, 3, 7, 65537
    allq, allpsi = T.prime_chain(N, [60] * (L + Lb))
    qs, psis, qb, psib = allq[:L], allpsi[:L], allq[L:], allpsi[L:]
    cq, cb = T.Context(N, qs, psis), T.Context(N, qb, psib)
    rng = np.random.default_rng(8)
    c = _rand(rng, N, qs, (3, 2))
    d, d2 = cq.to_device(c), cq.to_device(c)
    sq = H(cq.bfv_mul(cb, t, d, d))
    assert np.array_equal(sq, H(cq.bfv_mul(cb, t, d, d2)))
    assert np.array_equal(sq, CO.bfv_mul(CO.Rns(N, qs, psis), CO.Rns(N, qb, psib), t, c, c))
    ten = H(cq.ct_tensor(d, d))
    assert np.array_equal(ten, H(cq.ct_tensor(d, d2)))
    assert np.array_equal(ten, CO.Rns(N, qs, psis).ct_tensor(c, c))


def test_bfv_mul_joint_basis_extreme_residues():
    """the 128-bit sums of the joint-basis kernels are reduced by Solinas folds at bit 60: all-(q-1) operands and
    operands whose tensor values sit at +-Q/2 maximise every accumulation; headline shape (L = 8, K = 9)"""
    N, L, Lb, t = 256, 8, 17, 65537
    allq, allpsi = T.prime_chain(N, [60] * (L + Lb))
    qs, psis, qb, psib = allq[:L], allpsi[:L], allq[L:], allpsi[L:]
    cq, cb = T.Context(N, qs, psis), T.Context(N, qb, psib)
    oq, ob = CO.Rns(N, qs, psis), CO.Rns(N, qb, psib)
    c1 = np.empty((4, 2, L, N), dtype=np.uint64)
    c2 = np.empty_like(c1)
    Q = math.prod(qs)
    rng = np.random.default_rng(3)
    for i, q in enumerate(qs):
        c1[0, :, i, :] = q - 1
        c2[0, :, i, :] = q - 1
        c1[1, :, i, :] = (Q >> 1) % q
        c2[1, :, i, :] = ((Q >> 1) + 1) % q
        c1[2, :, i, :] = rng.integers(0, q, size=(2, N), dtype=np.uint64)
        c2[2, :, i, :] = q - 1
        c1[3, :, i, :] = 0
        c2[3, :, i, :] = rng.integers(0, q, size=(2, N), dtype=np.uint64)
    want = CO.bfv_mul(oq, ob, t, c1, c2)
    got = H(cq.bfv_mul(cb, t, cq.to_device(c1), cq.to_device(c2)))
    assert np.array_equal(got, want)


# --------------------------------------------------------------------- CKKS encoding on the device (ckksencoding.jl:60-101)
from oracle import ckks_oracle as CK


@pytest.mark.parametrize("logN,logqs,scale", [(5, [40, 40, 40], 2.0 ** 40), (13, [60, 40, 40], 2.0 ** 40), (15, [60, 40], 2.0 ** 30), (4, [40, 40], 2.0 ** 60 / 3),
                                              (6, [60, 60, 60], 2.0 ** 70)])   # docs/src/man/ckks.md scale: integers beyond 64 bits
def test_ckks_encode_decode(logN, logqs, scale):
    """tfb_ckks_encode / tfb_ckks_decode against the numpy restatement of ckksencoding.jl: encoded integers equal up
    to one unit where float64 rounding decides a half-integer, decode(encode(z)) = z and decode of arbitrary centred
    coefficients equal to the host formula within the reference tests' tolerances; batches of polynomials"""
    import torch
    N = 1 << logN
    qs, psis, ctx, _ = _ring(N, logqs)
    Q = math.prod(qs)
    rng = np.random.default_rng(logN)
    P = 3
    z = rng.uniform(-4, 4, (P, N // 2)) + 1j * rng.uniform(-4, 4, (P, N // 2))
    z[0, :2] = [1.0, -2.5j]
    d = torch.from_numpy(z).cuda()
    enc = H(ctx.ckks_encode(scale, d))
    mag = scale * 8
    for p in range(P):
        want = CK.encode(z[p], scale, N)
        tol = max(1.0, mag * 2e-15 * logN)                                                       # float64 FFT round-off, then one rounding
        d0 = None
        for i, q in enumerate(qs):                                                               # the same integer under every prime
            di = [((int(a) - w) % q + q // 2) % q - q // 2 for a, w in zip(enc[p, i], want)]      # centred difference mod q_i
            assert max(abs(v) for v in di) <= tol
            assert d0 is None or di == d0
            d0 = di
    back = ctx.ckks_decode(scale, ctx.ckks_encode(scale, d)).cpu().numpy()
    assert np.allclose(back, z, atol=max(1e-9, 4 * N / scale))
    # decode of arbitrary coefficients (large centred values): host formula of ckksencoding.jl:60-70
    xs = [int.from_bytes(rng.bytes(32), "little") % Q for _ in range(N)]
    xs[:4] = [0, 1, Q - 1, Q // 2]
    res = np.array([[x % q for x in xs] for q in qs], dtype=np.uint64)[None]
    want_dec = CK.decode([x - Q if x > Q // 2 else x for x in xs], scale, N)
    got = ctx.ckks_decode(scale, ctx.to_device(res)).cpu().numpy()[0]
    assert np.allclose(got, want_dec, rtol=1e-9, atol=1e-9 * np.abs(want_dec).max())
    with pytest.raises(T.EngineError):
        ctx.ckks_encode(2.0 ** 140, d)                                                           # scale * coefficient beyond 2^126


# --------------------------------------------------------------------- sampling on the device (poly.jl:7-23)
def test_device_sampler_matches_its_cpu_restatement_and_the_distributions():
    """tfb_sample_uniform / tfb_sample_gaussian vs oracle/sampler_oracle.py (same Philox counters): uniform residues
    bit-exact (integer path, including the rejection step that removes modulo bias -- 1 draw in 16 is rejected at
    q ~ 2^60), rounded Gaussian equal up to libm ulps; ranges, reproducibility, stream independence, moments"""
    from oracle import sampler_oracle as SO
    N = 1024
    qs, psis, ctx, _ = _ring(N, [60, 60, 40])
    seed = 0x1234567890ABCDEF
    got = H(ctx.sample_uniform(seed, 7, (3,)))
    assert np.array_equal(got, SO.sample_uniform(seed, 7, 3, qs, N))
    for i, q in enumerate(qs):
        assert got[:, i, :].max() < q
    assert np.array_equal(got, H(ctx.sample_uniform(seed, 7, (3,))))                 # pure function of (seed, stream, position)
    assert not np.array_equal(got, H(ctx.sample_uniform(seed, 8, (3,))))
    assert not np.array_equal(got, H(ctx.sample_uniform(seed + 1, 7, (3,))))
    big = H(ctx.sample_uniform(seed, 9, (64,)))[:, 0, :].astype(np.float64) / qs[0]   # 65536 draws: 16 equal buckets
    counts = np.histogram(big, bins=16, range=(0.0, 1.0))[0]
    assert abs(big.mean() - 0.5) < 0.01 and counts.min() > 3700 and counts.max() < 4500
    sigma = 3.2
    g = H(ctx.sample_gaussian(sigma, seed, 11, (16,)))
    want = SO.sample_gaussian(sigma, seed, 11, 16, qs, N)
    assert (g != want).mean() < 1e-5
    x = g[:, 0, :].astype(np.int64)
    x = np.where(x > qs[0] // 2, x - qs[0], x)                                      # centred lift
    for i, q in enumerate(qs):                                                        # the same integer under every prime
        xi = g[:, i, :].astype(np.int64)
        assert np.array_equal(np.where(xi > q // 2, xi - q, xi), x)
    assert abs(x.mean()) < 0.1 and abs(x.std() - np.sqrt(sigma ** 2 + 1 / 12)) < 0.1 and np.abs(x).max() < 8 * sigma
    with pytest.raises(T.EngineError):
        ctx.sample_gaussian(-1.0, seed, 0)


# --------------------------------------------------------------------- BFV plaintext maps (bfv.jl:21-29)
def _to_rns(xs, qs):
    return np.array([[x % q for x in xs] for q in qs], dtype=np.uint64)


@pytest.mark.parametrize("N,logqs,t", [(64, [60, 60, 40], 53), (1024, [60] * 8, 65537), (32, [50, 50], 256), (16, [40], 7),
                                       (64, [60, 60], (1 << 31) - 1)])
def test_bfv_encode_decode(N, logqs, t):
    """pi^-1 = Delta*m and pi = mod(divround(SignedMod(x), Delta), t): bit-exact against big-integer arithmetic on
    random values, exact ties (2r == Delta), neighbours of ties, 0, +-1 and the ends of the centred range"""
    import math
    qs, psis, ctx, orc = _ring(N, logqs)
    Q = math.prod(qs)
    rng = np.random.default_rng(N + t)
    for delta in (Q // t, Q // t + 1, (Q // t) | 1, (Q // t) & ~1):     # floor(Q/t) as in the reference's tests, odd and even
        m = rng.integers(0, 1 << 63, size=(3, N), dtype=np.uint64)
        got = H(ctx.bfv_encode(t, delta, ctx.to_device(m)))
        for p in range(3):
            assert np.array_equal(got[p], _to_rns([delta * (int(v) % t) for v in m[p]], qs))
        xs = [int.from_bytes(rng.bytes(80), "little") % Q for _ in range(N)]
        half = delta // 2
        special = [0, 1, Q - 1, Q // 2, Q // 2 + 1, Q // 2 - 1, delta, delta - 1, delta + 1, half, half + 1, half - 1 if half else 0,
                   Q - half, Q - half - 1, Q - half + 1, 5 * delta + half, 5 * delta + half + 1, 5 * delta + half - 1,
                   Q - (7 * delta + half), Q - (7 * delta + half) - 1, Q - (7 * delta + half) + 1, (t // 2) * delta, (t // 2) * delta + half]
        for i, v in enumerate(special[:N]):
            xs[i] = v % Q
        want = [O.rha(O.centre(x, Q), delta) % t for x in xs]
        got = H(ctx.bfv_decode(t, delta, ctx.to_device(_to_rns(xs, qs)[None])))[0]
        assert [int(v) for v in got] == want
        # round trip through the plaintext maps alone (no noise): decode(encode(m)) = m mod t
        back = H(ctx.bfv_decode(t, delta, ctx.bfv_encode(t, delta, ctx.to_device(m))))
        if delta == Q // t:
            assert np.array_equal(back, m % np.uint64(t))
    with pytest.raises(T.EngineError):
        ctx.bfv_decode(t, 1, ctx.to_device(_to_rns([0] * N, qs)[None]))          # Delta far too small for a word-size quotient
    # host-buffer variants: same results from numpy buffers
    delta = Q // t
    m = rng.integers(0, t, size=(2, N), dtype=np.uint64)
    enc = np.empty((2, len(qs), N), dtype=np.uint64)
    ctx.bfv_encode_host(t, delta, m, enc)
    assert np.array_equal(enc, H(ctx.bfv_encode(t, delta, ctx.to_device(m))))
    dec = np.empty((2, N), dtype=np.uint64)
    ctx.bfv_decode_host(t, delta, enc, dec)
    assert np.array_equal(dec, m)


# --------------------------------------------------------------------- key switching
@pytest.mark.parametrize("w", [0, 1, 2, 7, 31])
def test_keyswitch_digits(w):
    N = 64
    qs, psis, ctx, orc = _ring(N, [60, 60, 40])
    rng = np.random.default_rng(w)
    cend = _rand(rng, N, qs, (2,))
    got = H(ctx.keyswitch_digits(ctx.to_device(cend), w))
    for b in range(2):
        assert np.array_equal(got[b], orc.keyswitch_digits(cend[b], w))


@pytest.mark.parametrize("N,logqs,w,comps", [(64, [60, 60, 40], 0, 3), (64, [60, 60, 40], 7, 3), (64, [50, 50], 1, 2),
                                             (2048, [50, 50], 1, 3), (1024, [60] * 4, 2, 3),
                                             # N = 2^12, 2^13: base-2^w digits written once and transformed under every prime
                                             (4096, [60, 60, 40], 2, 3), (4096, [50, 50], 7, 2), (8192, [60, 40, 40], 3, 3),
                                             # N = 2^14; w = 5, 7 do not divide 64: digits that straddle two limbs of the integer
                                             (16384, [60, 40], 5, 2), (16384, [60, 60, 60], 7, 3),
                                             # N = 2^15: CRT digits formed inside the first global level of the transform
                                             (32768, [60, 40, 40], 0, 2), (32768, [60, 40], 0, 3)])
def test_keyswitch_plain(N, logqs, w, comps):
    qs, psis, ctx, orc = _ring(N, logqs)
    rng = np.random.default_rng(N + w)
    D = ctx.L if w == 0 else T.ndigits(qs, w)
    key = _rand(rng, N, qs, (D, 2))
    B = 2
    ct = _rand(rng, N, qs, (B, comps))
    key_dual = ctx.ntt_fwd(ctx.to_device(key))
    got = H(ctx.keyswitch(key_dual, ctx.to_device(ct), w))
    for b in range(B):
        dg = orc.keyswitch_digits(ct[b, comps - 1], w)
        c1 = ct[b, 0]
        c2 = ct[b, 1] if comps == 3 else np.zeros_like(ct[b, 0])
        w1, w2 = orc.keyswitch_accum(dg, key, c1, c2)
        assert np.array_equal(got[b, 0], w1) and np.array_equal(got[b, 1], w2)


@pytest.mark.parametrize("N,logqs,w,B", [(2048, [60, 60], 1, 2),      # 8 Ki coefficients: 16 digit lanes per coefficient
                                         (4096, [60] * 4, 2, 4),      # 64 Ki: 8 lanes
                                         (4096, [60] * 4, 2, 12),     # 192 Ki: 4 lanes
                                         (4096, [60] * 4, 2, 24)])    # above the threshold: one thread per coefficient
def test_keyswitch_small_batch_split_accumulate(N, logqs, w, B):
    """One ciphertext (or a residue shard of it) has too few coefficients to fill the GPU with one thread each: the key
    accumulation then splits the digit range over the warps of a CTA (ks_accum_split_kernel).  Same bits as the oracle and
    as the single-thread-per-coefficient kernel."""
    qs, psis, ctx, orc = _ring(N, logqs)
    rng = np.random.default_rng(N + w + B)
    D = T.ndigits(qs, w)
    key = _rand(rng, N, qs, (D, 2))
    ct = _rand(rng, N, qs, (B, 3))
    key_dual = ctx.ntt_fwd(ctx.to_device(key))
    got = H(ctx.keyswitch(key_dual, ctx.to_device(ct), w))
    T.force_generic(1)
    try:
        plain = H(ctx.keyswitch(key_dual, ctx.to_device(ct), w))
    finally:
        T.force_generic(0)
    assert np.array_equal(got, plain)
    for b in range(min(B, 2)):
        dg = orc.keyswitch_digits(ct[b, 2], w)
        w1, w2 = orc.keyswitch_accum(dg, key, ct[b, 0], ct[b, 1])
        assert np.array_equal(got[b, 0], w1) and np.array_equal(got[b, 1], w2)


@pytest.mark.parametrize("N,logqs,w,comps,world", [(64, [60, 60, 40], 7, 3, 2), (1024, [60] * 4, 2, 3, 4), (64, [50, 50, 50], 0, 2, 3),
                                                   (2 ** 14, [60] * 8, 2, 3, 8)])
def test_keyswitch_residue_shards_equal_whole(N, logqs, w, comps, world):
    """BASELINE config 4 (residues sharded over GPUs): the rows each rank computes with tfb_keyswitch_shard, put
    together, are exactly tfb_keyswitch's result (and the oracle's at the sizes it finishes quickly)"""
    from toyfhe_b200 import sharding as S
    qs, psis, ctx, orc = _ring(N, logqs)
    rng = np.random.default_rng(N + w + world)
    D = ctx.L if w == 0 else T.ndigits(qs, w)
    key = _rand(rng, N, qs, (D, 2))
    B = 2 if N < 2 ** 14 else 1
    ct = _rand(rng, N, qs, (B, comps))
    key_dual = ctx.ntt_fwd(ctx.to_device(key))
    d_ct = ctx.to_device(ct)
    whole = H(ctx.keyswitch(key_dual, d_ct, w))
    rows = []
    for rank in range(world):
        lo, hi = S.shard_range(ctx.L, rank, world)
        shard = T.Context(N, qs[lo:hi], psis[lo:hi])
        rows.append(H(ctx.keyswitch_shard(shard, lo, S.key_rows_for_shard(key_dual, lo, hi), d_ct, w)))
    assert np.array_equal(np.concatenate(rows, axis=-2), whole)
    if N <= 1024:
        for b in range(B):
            dg = orc.keyswitch_digits(ct[b, comps - 1], w)
            c2 = ct[b, 1] if comps == 3 else np.zeros_like(ct[b, 0])
            w1, w2 = orc.keyswitch_accum(dg, key, ct[b, 0], c2)
            assert np.array_equal(whole[b, 0], w1) and np.array_equal(whole[b, 1], w2)
    with pytest.raises(T.EngineError):
        ctx.keyswitch_shard(T.Context(N, qs[:1], psis[:1]), 1, S.key_rows_for_shard(key_dual, 0, 1), d_ct, w)   # wrong offset


def test_keyswitch_modulus_raised_full_size_routes_agree():
    """BASELINE config 3 shape (N = 2^15, ModulusRaised, CRT digits): the route that forms the digits inside the first
    global level of the transform equals the route that writes the digit rows and transforms them (kernel generation 1)"""
    N = 1 << 15
    qs, psis = T.prime_chain(N, [40, 40, 40, 60, 60])
    qs, psis = [qs[3]] + qs[:3] + [qs[4]], [psis[3]] + psis[:3] + [psis[4]]            # q0 (60), 3 x 40, special (60)
    ctx_ct, ctx_ext = T.Context(N, qs[:4], psis[:4]), T.Context(N, qs, psis)
    rng = np.random.default_rng(15)
    key_dual = ctx_ext.to_device(_rand(rng, N, qs, (4, 2)))
    ct = ctx_ct.to_device(_rand(rng, N, qs[:4], (5, 2)))
    fused = H(ctx_ct.keyswitch(key_dual, ct, 0, ext=ctx_ext))
    T.ntt_version(1)
    try:
        plain = H(ctx_ct.keyswitch(key_dual, ct, 0, ext=ctx_ext))
    finally:
        T.ntt_version(3)
    assert np.array_equal(fused, plain)


def test_keyswitch_modulus_raised_matches_python_oracle():
    # test/ckks_modraise.jl shape: N = 32, (q0, q1, special), CRT digits
    N = 32
    qs, psis = T.prime_chain(N, [40, 40, 40])
    ctx_ct = T.Context(N, qs[:2], psis[:2])
    ctx_ext = T.Context(N, qs, psis)
    rng = np.random.default_rng(5)
    key = _rand(rng, N, qs, (3, 2))           # one component per prime of the key ring (incl. special)
    ct = _rand(rng, N, qs[:2], (2, 2))
    key_dual = ctx_ext.ntt_fwd(ctx_ext.to_device(key[:2]))   # digits 0..1, rows [q0, q1, special]
    got = H(ctx_ct.keyswitch(key_dual, ctx_ct.to_device(ct), 0, ext=ctx_ext))
    tl = lambda a: [[int(x) for x in r] for r in a]
    for b in range(2):
        want = O.keyswitch_modraised([tl(ct[b, 0]), tl(ct[b, 1])], [(tl(key[d, 0]), tl(key[d, 1])) for d in range(3)], qs, psis, 0)
        assert tl(got[b, 0]) == want[0] and tl(got[b, 1]) == want[1]


# ----------------------------------------------------------------- host entry points
def test_host_entry_points(q8, psi8):
    import torch
    N = 4096
    qs, psis, ctx, orc = _ring(N, [60, 60, 60])
    rng = np.random.default_rng(8)
    a, b = _rand(rng, N, qs, (2,)), _rand(rng, N, qs, (2,))
    out = np.empty_like(a)
    ctx.ntt_fwd_host(a, out)
    assert np.array_equal(out, orc.nntt(a))
    ctx.ntt_inv_host(out, out)
    assert np.array_equal(out, a)
    ctx.ring_mul_host(a, b, out)
    assert np.array_equal(out, orc.ring_mul(a, b))
    c1, c2 = _rand(rng, N, qs, (2, 2)), _rand(rng, N, qs, (2, 2))
    pin = lambda x: torch.from_numpy(x.view(np.int64)).pin_memory()
    o3 = torch.empty((2, 3, 3, N), dtype=torch.int64).pin_memory()
    ctx.ct_tensor_host(pin(c1), pin(c2), o3)
    assert np.array_equal(o3.numpy().view(np.uint64), orc.ct_tensor(c1, c2))
    r = np.empty((2, 2, N), dtype=np.uint64)
    ctx.rescale_host(a, r)
    assert np.array_equal(r, orc.modswitch(a))


def test_error_paths():
    N = 1024
    qs, psis, ctx, _ = _ring(N, [60, 60])
    x = ctx.empty((3, N))  # not a multiple of L rows
    with pytest.raises(AssertionError):
        ctx.ntt_fwd(x)
    import ctypes as C
    lib = T.load_library()
    rc = lib.tfb_ntt_fwd(ctx.h, C.c_void_p(x.data_ptr()), C.c_void_p(x.data_ptr()), C.c_uint64(3), None)
    assert rc == 1 and b"multiple" in lib.tfb_last_error()
    rc = lib.tfb_ntt_fwd(ctx.h, None, None, C.c_uint64(0), None)   # empty input is fine
    assert rc == 0
    with pytest.raises(T.EngineError):
        ctx.galois(ctx.empty((2, N)), 4)                           # even Galois element
    with pytest.raises(T.EngineError):
        T.Context(1 << 17, [97], [33])


# ------------------------------------------------ kernel variants must all agree
@pytest.mark.parametrize("version,mode", [(1, 0), (1, 1), (3, 0), (3, 1), (3, 2)])
@pytest.mark.parametrize("logN,logqs", [(14, [60] * 8), (14, [60, 40, 40]), (15, [60, 60]), (16, [60]), (13, [60, 40]),
                                        (13, [60, 40, 40, 40, 40, 40, 60]), (12, [50, 50, 32]), (15, [60, 40, 40])])
def test_ntt_kernel_variants(version, mode, logN, logqs):
    N = 1 << logN
    qs, psis, ctx, orc = _ring(N, logqs)
    rng = np.random.default_rng(logN + version)
    # more rows than resident CTAs: the persistent kernels loop over several rows
    B = {14: 40, 13: 100, 12: 250, 15: 45, 16: 80}.get(logN, 3) if len(logqs) != 7 else 45   # (2^15, 2^16: more pairs than resident clusters)
    a = _rand(rng, N, qs, (B,))
    a[0, 0, :] = qs[0] - 1                # worst case for the lazy ranges
    want = orc.nntt(a)
    T.ntt_version(version)
    T.ntt_max_mode(mode)
    try:
        d = ctx.to_device(a)
        f = ctx.ntt_fwd(d)
        assert np.array_equal(H(f), want)
        assert np.array_equal(H(ctx.ntt_inv(f)), a)
        f2 = d.clone()
        ctx.ntt_fwd(f2, out=f2)           # in place
        assert np.array_equal(H(f2), want)
        ctx.ntt_inv(f2, out=f2)
        assert np.array_equal(H(f2), a)
    finally:
        T.ntt_version(3)
        T.ntt_max_mode(2)


@pytest.mark.parametrize("logN,logqs,B", [(15, [60, 40, 40], 60), (16, [60], 80), (15, [50], 3), (15, [60] + [40] * 9 + [60], 30)])
def test_ntt_long_rows_last_global_level_on_load(logN, logqs, B):
    """rows of 2^15 / 2^16, forward, out of place: the last global level applied while the sub-block kernel loads (default)
    against every level as its own pass and against the oracle; in place falls back to the separate passes"""
    N = 1 << logN
    qs, psis, ctx, orc = _ring(N, logqs)
    rng = np.random.default_rng(logN + len(logqs))
    a = _rand(rng, N, qs, (B,))
    a[0, 0, :] = qs[0] - 1
    want = orc.nntt(a[:4])
    d = ctx.to_device(a)
    f = ctx.ntt_fwd(d)
    assert np.array_equal(H(f[:4]), want)
    T.ntt_cross(False)
    try:
        g = ctx.ntt_fwd(d)
    finally:
        T.ntt_cross(True)
    assert bool((f == g).all())
    f2 = d.clone()
    ctx.ntt_fwd(f2, out=f2)
    assert bool((f2 == f).all())
    assert np.array_equal(H(ctx.ntt_inv(f)), a)


def test_many_primes_and_conversion_limits():
    """contexts hold up to 64 primes (transforms and element-wise work are per prime); the exact base conversions stop at
    32 primes and say so instead of computing something else"""
    N = 4096
    qs, psis = T.prime_chain(N, [60] * 24 + [50] * 16)
    assert len(qs) == 40
    ctx, orc = T.Context(N, qs, psis), CO.Rns(N, qs, psis)
    rng = np.random.default_rng(40)
    a = _rand(rng, N, qs, (3,))
    d = ctx.to_device(a)
    f = ctx.ntt_fwd(d)
    assert np.array_equal(H(f), orc.nntt(a))
    assert np.array_equal(H(ctx.ntt_inv(f)), a)
    assert np.array_equal(H(ctx.ring_mul(d, d)), orc.ring_mul(a, a))
    with pytest.raises(T.EngineError):
        ctx.keyswitch_digits(d, 2)                       # 40-prime Garner: unsupported, reported
    with pytest.raises(T.EngineError):
        T.Context(N, qs + qs[:30], psis + psis[:30])     # repeated moduli / more than 64 primes


def test_non_lazy_primes_take_the_harvey_ladder():
    # PALISADE prime 2^60 - 16383 (src/cryptparams.jl:25) is not of the form 2^b + small: Harvey path
    q, N, psi = 1152921504606830593, 2048, 811032584449645127
    ctx, orc = T.Context(N, [q], [psi]), CO.Rns(N, [q], [psi])
    rng = np.random.default_rng(0)
    a = rng.integers(0, q, size=(4, 1, N), dtype=np.uint64)
    a[0, 0, :] = q - 1
    f = ctx.ntt_fwd(ctx.to_device(a))
    assert np.array_equal(H(f), orc.nntt(a))
    assert np.array_equal(H(ctx.ntt_inv(f)), a)
    # a 61-bit prime close to 2^61 (no headroom for the lazy ladder)
    N = 1 << 14
    q = O.nextprime(2 ** 61 + 2 ** 59 + 1, 2 * N)
    psi = T.minimal_primitive_root(q, 2 * N)
    ctx, orc = T.Context(N, [q], [psi]), CO.Rns(N, [q], [psi])
    a = rng.integers(0, q, size=(2, 1, N), dtype=np.uint64)
    a[0, 0, :] = q - 1
    f = ctx.ntt_fwd(ctx.to_device(a))
    assert np.array_equal(H(f), orc.nntt(a))
    assert np.array_equal(H(ctx.ntt_inv(f)), a)


@pytest.mark.parametrize("generic", [False, True])
def test_base_conversion_specialisations_agree(generic, q8, psi8):
    N = 1024
    allq, allpsi = T.prime_chain(N, [60] * 25)
    qs, psis, qb, psib = allq[:8], allpsi[:8], allq[8:], allpsi[8:]
    cq, cb = T.Context(N, qs, psis), T.Context(N, qb, psib)
    oq, ob = CO.Rns(N, qs, psis), CO.Rns(N, qb, psib)
    rng = np.random.default_rng(17)
    c1, c2 = _rand(rng, N, qs, (2, 2)), _rand(rng, N, qs, (2, 2))
    big = _rand(rng, N, qb, (3,))
    Q, Qb = math.prod(qs), math.prod(qb)
    for k, X in enumerate([Qb >> 1, (Qb >> 1) + 1, 0, 1, Qb - 1, Q >> 1, (Q >> 1) + 1, Q, Q - 1, Qb // 65537, Qb // 65537 + 1]):
        for j, p in enumerate(qb):
            big[0, j, k] = X % p
    for k, X in enumerate([Q >> 1, (Q >> 1) + 1, 0, Q - 1, 1]):
        for i, q in enumerate(qs):
            c1[0, 0, i, k] = X % q
    T.force_generic(generic)
    try:
        assert np.array_equal(H(cq.bfv_switch(cb, cq.to_device(c1))), CO.bfv_switch(N, qs, qb, c1))
        assert np.array_equal(H(cq.bfv_contract(cb, 65537, cb.to_device(big))), CO.bfv_contract(N, qs, qb, 65537, big))
        # a plaintext modulus large enough that the sub-basis trick must fall back to the full basis
        tbig = (1 << 61) - 1
        assert np.array_equal(H(cq.bfv_contract(cb, tbig, cb.to_device(big))), CO.bfv_contract(N, qs, qb, tbig, big))
        assert np.array_equal(H(cq.bfv_mul(cb, 65537, cq.to_device(c1), cq.to_device(c2))), CO.bfv_mul(oq, ob, 65537, c1, c2))
    finally:
        T.force_generic(False)


def test_one_context_from_two_streams_and_two_devices_guard():
    """the scratch of a context is ordered across streams by the library (include/toyfhe_b200.h "Threading"): composite
    operations issued alternately on two streams, without any synchronisation in between, give the single-stream results;
    and every call leaves the caller's current device as it found it"""
    import torch
    N = 4096
    qs, psis, ctx, orc = _ring(N, [60, 60, 40])
    rng = np.random.default_rng(11)
    a, b = _rand(rng, N, qs, (6, 2)), _rand(rng, N, qs, (6, 2))
    da, db = ctx.to_device(a), ctx.to_device(b)
    want = orc.ct_tensor(a, b)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    outs = []
    for i in range(6):                                   # ct_tensor stages its transforms in the context's scratch
        st = s1 if i % 2 == 0 else s2
        outs.append(ctx.ct_tensor(da[i:i + 1], db[i:i + 1], stream=st))
    torch.cuda.synchronize()
    for i in range(6):
        assert np.array_equal(H(outs[i]), want[i:i + 1]), i
    assert torch.cuda.current_device() == ctx.device


@pytest.mark.parametrize("logN,logqs,B,comps", [(12, [60, 40, 40, 60], 5, 2), (13, [60, 40, 40, 40, 40, 40, 60], 40, 2), (14, [60, 60, 40, 60], 20, 3),
                                                (13, [40, 40, 60], 3, 3)])
def test_crt_digits_formed_inside_the_transform(logN, logqs, B, comps):
    """CRT keyswitch (relin_window = 0, special prime) at N = 2^12 .. 2^14: the digit polynomials are formed while the forward
    transform loads its rows (60-bit residues re-embedded under 40-bit primes take the Barrett branch) -- same result as
    the separate digit kernel + transform (force_generic), which test_keyswitch_* and tests/test_reference_transcriptions.py
    pin to the oracle and to the Julia transcription"""
    N = 1 << logN
    key_qs, key_psis = T.prime_chain(N, logqs)
    qs, psis = key_qs[:-1], key_psis[:-1]
    ctx, ext = T.Context(N, qs, psis), T.Context(N, key_qs, key_psis)
    rng = np.random.default_rng(logN + B)
    key = ext.ntt_fwd(ext.to_device(_rand(rng, N, key_qs, (len(key_qs), 2))))
    ct = _rand(rng, N, qs, (B, comps))
    ct[0, -1, :, 0] = [q // 2 for q in qs]
    ct[0, -1, :, 1] = [q // 2 + 1 for q in qs]
    ct[0, -1, :, 2] = [q - 1 for q in qs]
    d = ctx.to_device(ct)
    fused = H(ctx.keyswitch(key, d, 0, ext=ext))
    T.force_generic(1)
    try:
        plain = H(ctx.keyswitch(key, d, 0, ext=ext))
    finally:
        T.force_generic(0)
    assert np.array_equal(fused, plain)


def test_empty_batches_every_device_entry_point():
    """batch = 0 is a valid call everywhere (the reference maps over possibly empty arrays): OK, no launch, no pointer touched"""
    import ctypes as C
    N = 1024
    qs, psis = T.prime_chain(N, [60, 60, 60, 60, 60])
    cq, cb = T.Context(N, qs[:2], psis[:2]), T.Context(N, qs[2:], psis[2:])
    lib = T.load_library()
    z64, z32, nul = C.c_uint64(0), C.c_uint32(0), None
    before = T.kernel_launches()
    calls = [
        lib.tfb_ntt_fwd(cq.h, nul, nul, z64, nul), lib.tfb_ntt_inv(cq.h, nul, nul, z64, nul),
        lib.tfb_add(cq.h, nul, nul, nul, z64, nul), lib.tfb_sub(cq.h, nul, nul, nul, z64, nul), lib.tfb_mul(cq.h, nul, nul, nul, z64, nul),
        lib.tfb_neg(cq.h, nul, nul, z64, nul), lib.tfb_ring_mul(cq.h, nul, nul, nul, z64, nul),
        lib.tfb_galois(cq.h, C.c_uint64(3), nul, nul, z64, nul), lib.tfb_rescale(cq.h, nul, nul, z64, nul),
        lib.tfb_ct_tensor(cq.h, nul, nul, nul, z64, nul),
        lib.tfb_bfv_mul(cq.h, cb.h, C.c_uint64(65537), nul, nul, nul, z64, nul),
        lib.tfb_keyswitch(cq.h, None, C.c_uint32(2), nul, C.c_uint32(61), nul, C.c_uint32(3), nul, z64, nul),
        lib.tfb_keyswitch_shard(cq.h, cq.h, z32, C.c_uint32(2), nul, C.c_uint32(61), nul, C.c_uint32(3), nul, z64, nul),
    ]
    assert calls == [0] * len(calls), (calls, lib.tfb_last_error())
    assert T.kernel_launches() == before
    # the same calls with work to do and a null buffer are rejected, not dereferenced
    one = C.c_uint64(2)
    assert lib.tfb_ntt_fwd(cq.h, nul, nul, one, nul) == 1
    assert lib.tfb_ct_tensor(cq.h, nul, nul, nul, C.c_uint64(1), nul) == 1
    assert lib.tfb_keyswitch(cq.h, None, C.c_uint32(2), nul, C.c_uint32(61), nul, C.c_uint32(3), nul, C.c_uint64(1), nul) == 1
    assert lib.tfb_bfv_mul(cq.h, cb.h, C.c_uint64(65537), nul, nul, nul, C.c_uint64(1), nul) == 1
