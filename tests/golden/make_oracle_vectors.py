"""Writes tests/golden/oracle_vectors.npz: seeded input/output vectors of the hot-path operations, produced by the CPU
oracle (oracle/c_oracle.py, itself pinned to the reference's literal KATs by tests/test_oracle_kats.py and
tests/golden/reference_kats.json).  The reference is Julia with un-vendored dependencies and cannot be imported or run
here, so these fixtures are ORACLE outputs, not reference outputs: they freeze the oracle's
behaviour (a regression pin for both the oracle and the CUDA path) and let the GPU tests check the engine without
executing anything under oracle/.

    python tests/golden/make_oracle_vectors.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import c_oracle as CO          # noqa: E402
from oracle import toyfhe_oracle as O      # noqa: E402


def rnd(rng, qs, shape, N):
    out = np.empty(shape + (len(qs), N), dtype=np.uint64)
    for i, q in enumerate(qs):
        out[..., i, :] = rng.integers(0, q, size=shape + (N,), dtype=np.uint64)
    return out


def main():
    rng = np.random.default_rng(20240229)
    v = {}
    # transforms: mixed 60/40-bit chain at N = 64 and one 60-bit row at N = 4096 (third-generation kernels)
    for tag, N, logqs in (("ntt64", 64, (60, 40, 40)), ("ntt4096", 4096, (60,))):
        qs, psis = O.prime_chain(N, logqs)
        a = rnd(rng, qs, (2,), N)
        a[0, 0, :3] = [qs[0] - 1, 0, 1]
        orc = CO.Rns(N, qs, psis)
        v[tag + "_q"], v[tag + "_psi"], v[tag + "_in"], v[tag + "_fwd"] = np.array(qs, dtype=np.uint64), np.array(psis, dtype=np.uint64), a, orc.nntt(a)
        v[tag + "_inv"] = orc.inntt(a)
    # ciphertext tensor + BFV multiply (test/bfv_crt.jl shape scaled down: 2 + 4 primes; joint-basis shape 3 + 7 primes)
    for tag, N, L, Lb, t in (("bfv_a", 64, 2, 4, 53), ("bfv_b", 256, 3, 7, 65537)):
        allq, allpsi = O.prime_chain(N, (60,) * (L + Lb))
        qs, psis, qb, psib = allq[:L], allpsi[:L], allq[L:], allpsi[L:]
        c1, c2 = rnd(rng, qs, (2, 2), N), rnd(rng, qs, (2, 2), N)
        oq, ob = CO.Rns(N, qs, psis), CO.Rns(N, qb, psib)
        v[tag + "_q"], v[tag + "_psi"] = np.array(allq, dtype=np.uint64), np.array(allpsi, dtype=np.uint64)
        v[tag + "_meta"] = np.array([N, L, Lb, t], dtype=np.uint64)
        v[tag + "_c1"], v[tag + "_c2"] = c1, c2
        v[tag + "_tensor"] = oq.ct_tensor(c1, c2)
        v[tag + "_mul"] = CO.bfv_mul(oq, ob, t, c1, c2)
    # keyswitch (base-2^7 digits, 3 components), rescale, Galois at N = 64 over (60, 60, 40)
    N, w = 64, 7
    qs, psis = O.prime_chain(N, (60, 60, 40))
    orc = CO.Rns(N, qs, psis)
    D = CO.ndigits(qs, w)
    key, ct = rnd(rng, qs, (D, 2), N), rnd(rng, qs, (1, 3), N)
    w1, w2 = orc.keyswitch_accum(orc.keyswitch_digits(ct[0, 2], w), key, ct[0, 0], ct[0, 1])
    v["ks_q"], v["ks_psi"], v["ks_w"] = np.array(qs, dtype=np.uint64), np.array(psis, dtype=np.uint64), np.array([w], dtype=np.uint64)
    v["ks_key"], v["ks_ct"], v["ks_out"] = key, ct, np.stack([w1, w2])[None]
    a = rnd(rng, qs, (2,), N)
    v["rs_in"] = a
    v["rs_out"] = np.array([[[int(x) for x in row] for row in O.modswitch([[int(x) for x in r] for r in a[p]], qs)] for p in range(2)], dtype=np.uint64)
    v["gal_g"] = np.array([3, 2 * N - 1, 5], dtype=np.uint64)
    v["gal_out"] = np.array([[[O.apply_galois_element([int(x) for x in a[0, i]], int(g), qs[i]) for i in range(len(qs))]] for g in v["gal_g"]], dtype=np.uint64)
    # BFV plaintext maps at t = 65537 over the same ring
    import math
    Q = math.prod(qs)
    t = 65537
    m = rng.integers(0, t, size=(1, N), dtype=np.uint64)
    xs = [int.from_bytes(rng.bytes(24), "little") % Q for _ in range(N)]
    v["pt_t"], v["pt_m"] = np.array([t], dtype=np.uint64), m
    v["pt_enc"] = np.array([[[(Q // t) * int(x) % q for x in m[0]] for q in qs]], dtype=np.uint64)
    v["pt_b"] = np.array([[[x % q for x in xs] for q in qs]], dtype=np.uint64)
    v["pt_dec"] = np.array([[O.rha(O.centre(x, Q), Q // t) % t for x in xs]], dtype=np.uint64)
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **v)
    print("wrote", len(v), "arrays,", os.path.getsize(os.path.join(HERE, "oracle_vectors.npz")), "bytes")


if __name__ == "__main__":
    main()
