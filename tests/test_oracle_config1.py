"""BASELINE configs[0] ("BFV triv", test/bfv_triv.jl:5-22): one encrypt -> mul -> decrypt on the CPU, through the
oracle alone.  The reference test lets the parameter estimator pick a 70-bit Int128 prime; SURVEY.md section 8d fixes
the word-size equivalent -- N = 2^12, the single 60-bit chain prime q = 1152921504606904321, t = 53, a big ring of
further chain primes with P_big > N Q^2 -- and the same assertions: decrypt(c)[0] == 6, decrypt(c*c)[0] == 0x24 (mod t).
CPU only; the GPU replay of the RNS variant is tests/test_gpu_scheme.py::test_bfv_crt_replay."""
from oracle import toyfhe_oracle as O


def test_bfv_triv_on_the_cpu_oracle():
    N, t, sigma = 1 << 12, 53, 3.2
    chain, psis = O.prime_chain(N, (60,) * 4)
    assert chain[0] == 1152921504606904321
    qs, ps = chain[:1], psis[:1]
    qb, pb = chain[1:], psis[1:]                    # 3 x 60 bits > log2(N Q^2) = 132
    s = O.Sampler(2025)
    secret, pub = O.keygen(s, N, qs, ps, sigma)
    plain = [0] * N
    plain[0] = 6
    c = O.encrypt(s, pub, O.bfv_pi_inv(plain, t, qs), N, qs, ps, sigma)
    assert O.bfv_pi(O.decrypt_raw(secret, c, qs, ps), t, qs)[0] == 6
    y = O.bfv_mul(c, c, qs, ps, qb, pb, t)
    assert len(y) == 3
    dec = O.bfv_pi(O.decrypt_raw(secret, y, qs, ps), t, qs)
    assert dec[0] == 0x24 % t == 36
    assert all(v == 0 for v in dec[1:])
