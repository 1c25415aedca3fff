"""Replays of the reference's own tests for the RNS / power-of-two path through
the mirrored host interface + CUDA engine (same parameters, same assertions), and
ciphertext-level bit-exact parity against the oracle with a shared seeded sampler
(the reference's tests only pin decrypt-level results)."""
import math

import numpy as np
import pytest

import toyfhe_b200 as T
from oracle import c_oracle as CO
from oracle import toyfhe_oracle as O

pytestmark = pytest.mark.gpu


def tl(a):
    return [[int(x) for x in r] for r in a]


# ---------------------------------------------------------------- test/bfv_crt.jl
def _bfv_crt_params():
    n = 2048
    p1 = O.nextprime(2 ** 50 + 1, 2 * n)
    p2 = O.nextprime(p1 + 2 * n, 2 * n)
    chain = [p1, p2]
    for _ in range(4):
        chain.append(O.nextprime(chain[-1] + 2 * n, 2 * n))
    R = T.NegacyclicRing(n, qs=chain[:2])
    Rbig = T.NegacyclicRing(n, qs=chain[2:])
    return n, R, Rbig, T.BFVParams(R, Rbig, 53, relin_window=1, sigma=3.2)


def test_bfv_crt_replay():
    n, R, Rbig, params = _bfv_crt_params()
    s = T.Sampler(2024)
    kp = T.keygen(s, params)
    plain = [0] * n
    plain[0] = 6
    c = T.encrypt(s, kp, plain)
    assert T.decrypt(kp, c)[0] == 6
    y = c * c
    assert len(y) == 3
    assert T.decrypt(kp, y)[0] == 0x24 % 53 == 36
    # general hook path (expand -> component tensor -> contract) agrees with the fused kernel bit for bit
    e1 = params.mul_expand(c)
    acc = [e1[0] * e1[0], e1[0] * e1[1] + e1[1] * e1[0], e1[1] * e1[1]]
    hooks = params.mul_contract(acc)
    for a, b in zip(hooks, y.cs):
        assert np.array_equal(a.residues(), b.residues())


def test_bfv_simd_replay():
    """test/bfv_simd.jl:10-31 (BASELINE config "BFV SIMD") over the word-size RNS route: SlotEncoding plaintexts
    (slots = values at psi_t^(2k+1), t = 65537), encrypt, ciphertext multiply, decrypt, read the slots back.
    Encoding (NTT over F_t), pi^-1, the multiply and pi all run on the device."""
    n, t = 4096, 65537
    chain = [O.nextprime((1 << 60) + 1, 2 * n)]
    while len(chain) < 7:
        chain.append(O.nextprime(chain[-1] + 2 * n, 2 * n))
    R = T.NegacyclicRing(n, qs=chain[:2])          # Q ~ 2^120
    Rbig = T.NegacyclicRing(n, qs=chain[2:])       # 5 primes: P_big ~ 2^300 > N Q^2
    Rplain = T.NegacyclicRing(n, qs=[t])           # psi = minimal primitive 2N-th root mod t (pow2_cyc_rings.jl:38-41)
    params = T.BFVParams(R, Rbig, t, relin_window=1, sigma=3.2)
    s = T.Sampler(7)
    kp = T.keygen(s, params)
    plain = T.SlotEncoding(Rplain)
    plain[0] = 1
    plain[1] = 1
    plain2 = T.SlotEncoding(Rplain)
    plain2[:] = 10
    plain2[0] = 5
    c1, c2 = T.encrypt(s, kp, plain), T.encrypt(s, kp, plain2)
    y = c1 * c2
    data = T.SlotEncoding.from_coeffs(Rplain, T.decrypt(kp, y))
    assert data[0] == 5 and data[1] == 10
    assert all(v == 0 for v in data.slots[2:])
    # the slot maps invert each other and follow the oracle's NTT definition over F_t
    orc = CO.Rns(n, [t], Rplain.psis)
    rng = np.random.default_rng(1)
    v = rng.integers(0, t, size=n, dtype=np.uint64)
    se = T.SlotEncoding(Rplain, v)
    assert se.coeffs() == [int(x) for x in orc.inntt(v.reshape(1, 1, n)).reshape(-1)]
    assert T.SlotEncoding.from_coeffs(Rplain, se.coeffs()).slots == [int(x) for x in v]
    # full-width SIMD: slot-wise product of two random slot vectors
    a, b = rng.integers(0, t, size=n, dtype=np.uint64), rng.integers(0, t, size=n, dtype=np.uint64)
    prod = T.encrypt(s, kp, T.SlotEncoding(Rplain, a)) * T.encrypt(s, kp, T.SlotEncoding(Rplain, b))
    got = T.SlotEncoding.from_coeffs(Rplain, T.decrypt(kp, prod)).slots
    assert got == [int(x) * int(y) % t for x, y in zip(a, b)]


def test_bfv_triv_replay():
    """test/bfv_triv.jl (BASELINE configs[0]) over its word-size equivalent: N = 2^12, ONE 60-bit prime, t = 53,
    big ring of three further chain primes; decrypt(c)[0] == 6, decrypt(c*c)[0] == 0x24 mod t"""
    n = 1 << 12
    chain = [O.nextprime((1 << 60) + 1, 2 * n)]
    while len(chain) < 4:
        chain.append(O.nextprime(chain[-1] + 2 * n, 2 * n))
    assert chain[0] == 1152921504606904321
    R, Rbig = T.NegacyclicRing(n, qs=chain[:1]), T.NegacyclicRing(n, qs=chain[1:])
    params = T.BFVParams(R, Rbig, 53, relin_window=1, sigma=3.2)
    s = T.Sampler(12)
    kp = T.keygen(s, params)
    plain = [0] * n
    plain[0] = 6
    c = T.encrypt(s, kp, plain)
    assert T.decrypt(kp, c)[0] == 6
    dec = T.decrypt(kp, c * c)
    assert dec[0] == 0x24 % 53 and not any(dec[1:])


def test_bgv_triv_replay():
    """test/bgv_triv.jl: BGV over the PALISADE ring m = 4096 of src/cryptparams.jl:25 (q = 2^60 - 16383..., N = 2048 --
    not a 2^b + small prime, so the Harvey kernels run), t = 256, sigma = 8/sqrt(2 pi)"""
    q, n, psi = 1152921504606830593, 2048, 811032584449645127
    R = T.NegacyclicRing(n, qs=[q], psis=[psi])
    params = T.BGVParams(R, 256, 8 / math.sqrt(2 * math.pi))
    s = T.Sampler(31)
    kp = T.keygen(s, params)
    plain = [0] * n
    plain[0] = 6
    c = T.encrypt(s, kp, plain)
    assert T.decrypt(kp, c)[0] == 6
    dec = T.decrypt(kp, c * c)
    assert dec[0] == 0x24 and not any(dec[1:])
    # the plaintext map against big-integer arithmetic over an RNS ring (centred lift, negative values, 0, Q-1)
    R3 = T.NegacyclicRing(64, logqs=[60, 60, 40])
    Q = R3.modulus()
    rng = np.random.default_rng(3)
    xs = [int.from_bytes(rng.bytes(24), "little") % Q for _ in range(64)]
    xs[:5] = [0, 1, Q - 1, Q // 2, Q // 2 + 1]
    for t in (256, 53, 65537, (1 << 61) - 1):
        got = T.BGVParams(R3, t, 3.2).pi(R3(xs))
        assert got == [(x - Q if x > Q // 2 else x) % t for x in xs]


def test_bfv_crt_replay_with_device_sampler():
    """test/bfv_crt.jl with every random draw of keygen / encrypt made on the device (tfb_sample_*): nothing but the
    6 and the decrypted 36 crosses the host boundary"""
    n, R, Rbig, params = _bfv_crt_params()
    s = T.Sampler(77, device=True)
    kp = T.keygen(s, params)
    plain = [0] * n
    plain[0] = 6
    c = T.encrypt(s, kp, plain)
    assert T.decrypt(kp, c)[0] == 6
    assert T.decrypt(kp, c * c)[0] == 36
    ek = T.keygen_evalmult(s, kp.priv)
    assert T.decrypt(kp, T.keyswitch(ek, c * c))[0] == 36


def test_bfv_keyswitch_replay():
    """test/bfv_keyswitch.jl semantics on the RNS route: relinearise c*c with an
    EvalMultKey (relin_window = 1, base-2 digits), then multiply again"""
    n, R, Rbig, params = _bfv_crt_params()
    s = T.Sampler(7)
    kp = T.keygen(s, params)
    ek = T.keygen_evalmult(s, kp.priv)
    plain = [0] * n
    plain[0] = 2
    c1 = T.encrypt(s, kp, plain)
    c2 = c1 * c1
    cswitch = T.keyswitch(ek, c2)
    assert len(cswitch) == 2
    assert T.decrypt(kp, c2)[0] == 4
    assert T.decrypt(kp, cswitch)[0] == 4


def test_ciphertext_add_sub():
    n, R, Rbig, params = _bfv_crt_params()
    s = T.Sampler(3)
    kp = T.keygen(s, params)
    a = [0] * n; a[0] = 20; a[5] = 7
    b = [0] * n; b[0] = 11; b[5] = 9
    ca, cb = T.encrypt(s, kp, a), T.encrypt(s, kp, b)
    d = T.decrypt(kp, ca + cb)
    assert d[0] == 31 and d[5] == 16
    d = T.decrypt(kp, ca - cb)
    assert d[0] == 9 and d[5] == (7 - 9) % 53
    same = T.BFVParams(R, Rbig, 53)      # egal by value, as the reference's `!==` on immutable structs sees it
    assert T.decrypt(kp, ca + T.CipherText(same, cb.cs))[0] == 31
    other = T.BFVParams(R, Rbig, 59)     # a different plaintext modulus is a different parameter set
    with pytest.raises(T.UsageError):
        ca + T.CipherText(other, cb.cs)
    with pytest.raises(T.UsageError):
        ca * T.CipherText(other, cb.cs)


# ---------------------------------------------------------------- CKKS tests
def _ckks_ring(N, n_primes):
    q = [O.nextprime(2 ** 40 + 1, 2 * N)]
    for _ in range(n_primes - 1):
        q.append(O.nextprime(q[-1] + 2 * N, 2 * N))
    return T.NegacyclicRing(N, qs=q)


def test_ckks_triv_replay():
    """test/ckks_triv.jl: 2048 slots (N = 4096), scale 2^40, the encoder in isolation (re*re decoded at scale^2), then
    encrypt / decrypt and a ciphertext square.  The reference borrows a single ~180-bit prime from the BFV estimator;
    the word-size equivalent is an RNS ring of three 60-bit chain primes (same bit budget)."""
    N = 4096
    R = T.NegacyclicRing(N, logqs=[60, 60, 60])
    scale = 2.0 ** 40
    x = np.linspace(0.0, 1.0, N // 2)
    plain = T.CKKSEncoding(scale, x.astype(np.complex128))
    re = plain.to_ring_element(R)
    sq = T.CKKSEncoding.from_ring_element(re * re, scale * scale)
    assert np.allclose(np.real(sq.data), x ** 2, atol=1e-4)
    params = T.CKKSParams(R, 1, 3.2)
    s = T.Sampler(41)
    kp = T.keygen(s, params)
    c = T.encrypt(s, kp, plain)
    assert np.allclose(np.real(T.decrypt(kp, c).data), x, atol=1e-4)
    assert np.allclose(np.real(T.decrypt(kp, c * c).data), x ** 2, atol=1e-4)


def test_ckks_modswitch_replay():
    """test/ckks_modswitch.jl"""
    N = 32
    R = _ckks_ring(N, 3)
    scale = 2.0 ** 60
    plain = T.CKKSEncoding.zeros(scale, N)
    plain.data[:] = 2
    ps = R.qs[-1]
    switched = plain.to_ring_element(R).modswitch()
    assert abs(T.CKKSEncoding.from_ring_element(switched, scale / ps).data[0] - 2.0) < 1e-5
    params = T.CKKSParams(R, 1, 3.2)
    s = T.Sampler(11)
    kp = T.keygen(s, params)
    c = T.modswitch(T.encrypt(s, kp, plain))
    assert c.ring().L == 2
    dec = T.decrypt(kp, c)
    assert np.allclose(dec.data, plain.data, atol=1e-3)


def test_add_and_multiply_after_separate_modswitches():
    """examples/encrypted_mnist/infer.jl:137-160: ciphertexts rescaled by SEPARATE modswitch calls are added and
    multiplied.  The reference's `!==` on immutable parameter structs is egality by value, so the two DropLastParams
    wrappers are the same parameters (round-1 advisor finding: the mirror compared them by identity)."""
    N = 64
    R = _ckks_ring(N, 3)
    params = T.CKKSParams(R, 1, 3.2)
    s = T.Sampler(12)
    kp = T.keygen(s, params)
    scale = 2.0 ** 60
    x = np.linspace(0.25, 1.0, N // 2)
    y = np.linspace(-1.0, 0.5, N // 2)
    ca = T.modswitch(T.encrypt(s, kp, T.CKKSEncoding(scale, x.astype(np.complex128))))
    cb = T.modswitch(T.encrypt(s, kp, T.CKKSEncoding(scale, y.astype(np.complex128))))
    assert ca.params is not cb.params and ca.params == cb.params and hash(ca.params) == hash(cb.params)
    assert np.allclose(np.real(T.decrypt(kp, ca + cb).data), x + y, atol=1e-3)
    assert np.allclose(np.real(T.decrypt(kp, ca - cb).data), x - y, atol=1e-3)
    prod = ca * cb                                      # scale (2^60 / q_last)^2 ~ 2^40 on the two remaining primes
    assert np.allclose(np.real(T.decrypt(kp, prod).data), x * y, atol=1e-3)
    assert T.modswitch_drop(ca).params == T.modswitch_drop(cb).params
    assert T.modswitch_drop(ca).params != ca.params     # one level further down is a different parameter set
    with pytest.raises(T.UsageError):
        ca + T.modswitch_drop(cb)


def test_ckks_rotate_replay():
    """test/ckks_rotate.jl"""
    N = 16
    R = _ckks_ring(N, 2)
    scale = 2.0 ** 60
    plain = T.CKKSEncoding.zeros(scale, N)
    plain.data[:] = np.arange(1, N // 2 + 1)
    plain.data[0] += 1j
    re = plain.to_ring_element(R)
    rot = T.CKKSEncoding.from_ring_element(re.apply_galois_element(3), scale)
    assert np.allclose(rot.data, np.roll(plain.data, -1), atol=1e-6)
    params = T.CKKSParams(R, 1, 3.2)
    s = T.Sampler(5)
    kp = T.keygen(s, params)
    c = T.encrypt(s, kp, plain)
    cg = T.apply_galois_element(c, 3)
    ek = T.make_eval_key(s, kp.priv.secret.apply_galois_element(3), kp.priv)
    rt = T.decrypt(kp, T.keyswitch(ek, cg))
    assert np.allclose(rt.data, np.roll(plain.data, -1), atol=1e-4)
    gk = T.keygen_galois(s, kp.priv, steps=1)
    rt = T.decrypt(kp, T.rotate(gk, T.encrypt(s, kp, plain)))
    assert np.allclose(rt.data, np.roll(plain.data, 1), atol=1e-4)


def test_ckks_matmul_replay():
    """test/ckks_matmul.jl: 4x4 diagonal-method matmul, 3 rotations + 4 plaintext-vector mults"""
    N = 32
    R = _ckks_ring(N, 3)
    scale = 2.0 ** 40
    plain = T.CKKSEncoding.zeros(scale, N)
    plain.data[:] = np.arange(1, N // 2 + 1)
    W = np.ones((4, 4), dtype=np.float32)
    params = T.CKKSParams(R, 1, 3.2)
    s = T.Sampler(13)
    kp = T.keygen(s, params)
    c = T.encrypt(s, kp, plain)
    gk = T.keygen_galois(s, kp.priv, steps=4)

    def diag_rep(M, k):
        return np.tile(np.diag(np.roll(M, k, axis=1)), M.shape[1]).astype(np.float64)

    result = T.ckks_mul_plain_vector(diag_rep(W, 0), c)
    rotated = c
    for k in range(1, W.shape[1]):
        rotated = T.rotate(gk, rotated)
        result = result + T.ckks_mul_plain_vector(diag_rep(W, k), rotated)
    dec = T.decrypt(kp, result)
    got = np.real(dec.data).reshape(4, 4, order="F")
    want = (W.astype(np.float64) @ np.real(plain.data).reshape(4, 4, order="F").T).T
    assert np.allclose(got, want, atol=1e-5)


def test_ckks_modraise_replay():
    """test/ckks_modraise.jl: special-prime keyswitch (CRT digits) from the secret to itself"""
    N = 32
    R = _ckks_ring(N, 3)
    params = T.ModulusRaised(T.CKKSParams(R, 0, 3.2))
    s = T.Sampler(17)
    kp = T.keygen(s, params)
    scale = 2.0 ** 40
    plain = T.CKKSEncoding.zeros(scale, N)
    plain.data[:] = np.arange(1, N // 2 + 1)
    c = T.encrypt(s, kp, plain)
    assert c.ring().L == 2
    ek = T.make_eval_key(s, kp.priv.secret, kp.priv)
    assert len(ek.key) == 3
    dec = T.decrypt(kp, T.keyswitch(ek, c))
    assert np.allclose(dec.data, plain.data, atol=1e-8)


def test_ckks_ct_mul_and_rescale():
    """ciphertext * ciphertext then relinearise and rescale (docs/src/man/ckks.md flow)"""
    N = 64
    qs, psis = T.prime_chain(N, [60, 40, 40])
    R = T.NegacyclicRing(N, qs=[qs[1], qs[2], qs[0]], psis=[psis[1], psis[2], psis[0]])   # 60-bit prime last
    params = T.CKKSParams(R, 2, 3.2)   # base-4 digits: CRT digits without a special prime are too noisy
    s = T.Sampler(23)
    kp = T.keygen(s, params)
    ek = T.keygen_evalmult(s, kp.priv)
    scale = float(2 ** 30)
    a = T.CKKSEncoding(scale, np.linspace(-1, 1, N // 2))
    b = T.CKKSEncoding(scale, np.linspace(0.5, 2, N // 2))
    prod = T.keyswitch(ek, T.encrypt(s, kp, a) * T.encrypt(s, kp, b))
    dec = T.decrypt(kp, prod)
    assert np.allclose(dec.data, a.data * b.data, atol=1e-4)


def _mnist_style_ring(N, n40):
    """examples/encrypted_mnist/infer.jl:97-110: (q0 60-bit, n40 x 40-bit, special 60-bit), primes found the same way"""
    q0 = O.nextprime(2 ** 60 + 1, 2 * N)
    ps = O.nextprime(q0 + 2 * N, 2 * N)
    qs = [O.nextprime(2 ** 40 + 1, 2 * N)]
    for _ in range(n40 - 1):
        qs.append(O.nextprime(qs[-1] + 2 * N, 2 * N))
    return T.NegacyclicRing(N, qs=[q0] + qs + [ps])


@pytest.mark.parametrize("logN,n40", [(13, 5), (15, 9)])
def test_ckks_full_size_rotate_multiply_rescale(logN, n40):
    """BASELINE configs 5 and 3 at their full sizes (N = 2^13 with 7 primes; N = 2^15 with a 10-level chain + special
    prime, ModulusRaised CRT-digit keyswitch as in examples/encrypted_mnist/infer.jl:110): encrypt, rotate with a
    GaloisKey, multiply by a plaintext vector, rescale, decrypt -- the decrypted slots equal the plaintext computation
    (the size-independent property the reference's own CKKS tests check, test/ckks_rotate.jl, ckks_matmul.jl,
    ckks_modswitch.jl)"""
    N = 1 << logN
    R = _mnist_style_ring(N, n40)
    params = T.ModulusRaised(T.CKKSParams(R, 0, 3.2))
    s = T.Sampler(100 + logN)
    kp = T.keygen(s, params)
    scale = 2.0 ** 40
    rng = np.random.default_rng(logN)
    v = rng.uniform(-1, 1, N // 2)
    w = rng.uniform(-2, 2, N // 2)
    c = T.encrypt(s, kp, T.CKKSEncoding(scale, v.astype(np.complex128)))
    assert c.ring().L == n40 + 1                       # special prime dropped from the ciphertext ring
    gk = T.keygen_galois(s, kp.priv, steps=1)
    rot = T.rotate(gk, c)
    tol = 1e-6 if logN == 13 else 3e-5              # fresh + keyswitch noise grows with N at a fixed scale of 2^40
    assert np.allclose(np.real(T.decrypt(kp, rot).data), np.roll(v, 1), atol=tol)
    prod = T.modswitch(T.ckks_mul_plain_vector(w, rot))  # scale^2 / q_last ~ 2^40
    assert prod.ring().L == n40
    got = np.real(T.decrypt(kp, prod).data)
    assert np.allclose(got, w * np.roll(v, 1), atol=10 * tol)


# ------------------------------------------ ciphertext-level bit-exact parity vs the oracle
def test_scheme_bit_exact_vs_oracle():
    N = 64
    qs, psis = T.prime_chain(N, [60, 60, 40])
    R = T.NegacyclicRing(N, qs=qs, psis=psis)
    sigma = 3.2
    for w in (0, 2):
        params = T.CKKSParams(R, w, sigma)
        s_gpu, s_cpu = T.Sampler(99), O.Sampler(99)
        kp = T.keygen(s_gpu, params)
        sec, (mask, masked) = O.keygen(s_cpu, N, qs, psis, sigma)
        assert tl(kp.priv.secret.residues()) == sec
        assert tl(kp.pub.key.masked.residues()) == masked
        m = list(range(N))
        c = T.encrypt(s_gpu, kp, R(m))
        oc = O.encrypt(s_cpu, (mask, masked), O.rns_from_ints(m, qs), N, qs, psis, sigma)
        assert [tl(x.residues()) for x in c.cs] == oc
        c2 = c * c
        oc2 = O.ct_tensor(oc, oc, qs, psis)
        assert [tl(x.residues()) for x in c2.cs] == oc2
        ek = T.make_eval_key(s_gpu, kp.priv.secret ** 2, kp.priv)
        osec2 = O.rns_ring_multiply(sec, sec, qs, psis)
        okey = O.make_eval_key(s_cpu, osec2, sec, N, qs, psis, sigma, w)
        assert len(okey) == len(ek.key)
        assert tl(ek.key[-1].masked.residues()) == okey[-1][1]
        c3 = T.keyswitch(ek, c2)
        oc3 = O.keyswitch(oc2, okey, qs, psis, w)
        assert [tl(x.residues()) for x in c3.cs] == oc3
        # rotation
        g = T.galois_element_from_steps(1, N)
        assert g == O.galois_element_from_steps(1, N)
        cg = T.apply_galois_element(c, g)
        assert [tl(x.residues()) for x in cg.cs] == [O.rns_galois(x, g, qs) for x in oc]
        # rescale
        cr = [x.modswitch() for x in c.cs]
        assert [tl(x.residues()) for x in cr] == [O.modswitch(x, qs) for x in oc]
        # decrypt inner product
        b = kp.priv.secret * c.cs[1] + c.cs[0]
        assert tl(b.residues()) == O.decrypt_raw(sec, oc, qs, psis)
