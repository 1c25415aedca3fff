"""TEST INFRASTRUCTURE ONLY (imported by tests/): CPU restatement of the device sampler
(toyfhe.jl_b200/csrc/sample_kernels.cu) -- Philox4x32-10 keyed by the seed, counter = (position, stream, attempt).
It stands in for the reference's RingSampler (poly.jl:7-23; uniform residues crt.jl:146-148, rounded Gaussian
bfv.jl:31-32 / ckks.jl:24-25), whose own RNG is unseeded (rlwe_she.jl:169-170): parity with the REFERENCE is
statistical only; parity between the device sampler and this file is bit-exact on the integer path."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """vectorised over uint64 arrays holding 32-bit words"""
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint64) for v in (c0, c1, c2, c3))
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n1 = p1 & MASK
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        n3 = p0 & MASK
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def _draw(seed: int, stream: int, index: np.ndarray, attempt: int):
    idx = np.asarray(index, dtype=np.uint64)
    z = np.zeros_like(idx)
    return philox4x32_10(idx & MASK, idx >> np.uint64(32), z + np.uint64(stream), z + np.uint64(attempt),
                         seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)


def sample_uniform(seed: int, stream: int, polys: int, qs, N: int) -> np.ndarray:
    L = len(qs)
    out = np.empty((polys, L, N), dtype=np.uint64)
    for p in range(polys):
        for i, q in enumerate(qs):
            base = (p * L + i) * N
            idx = np.arange(base, base + N, dtype=np.uint64)
            limit = (1 << 64) - ((1 << 64) % q)          # multiples of q below 2^64
            res = np.zeros(N, dtype=np.uint64)
            todo = np.ones(N, dtype=bool)
            attempt = 0
            while todo.any():
                x, y, z, w = _draw(seed, stream, idx[todo], attempt)
                r_a = (y << np.uint64(32)) | x
                r_b = (w << np.uint64(32)) | z
                ok_a = np.array([int(v) < limit for v in r_a], dtype=bool)
                ok_b = np.array([int(v) < limit for v in r_b], dtype=bool)
                val = np.where(ok_a, r_a, r_b)
                done = ok_a | ok_b
                pos = np.flatnonzero(todo)
                res[pos[done]] = val[done] % np.uint64(q)
                todo[pos[done]] = False
                attempt += 1
            out[p, i] = res
    return out


def gaussian_ints(sigma: float, seed: int, stream: int, polys: int, N: int) -> np.ndarray:
    idx = np.arange(polys * N, dtype=np.uint64)
    x, y, z, w = _draw(seed, stream, idx, 0)
    r0 = (y << np.uint64(32)) | x
    r1 = (w << np.uint64(32)) | z
    u1 = ((r0 >> np.uint64(11)) + np.uint64(1)).astype(np.float64) * 2.0 ** -53
    u2 = (r1 >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
    zz = np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
    return np.rint(sigma * zz).astype(np.int64).reshape(polys, N)


def sample_gaussian(sigma: float, seed: int, stream: int, polys: int, qs, N: int) -> np.ndarray:
    x = gaussian_ints(sigma, seed, stream, polys, N)
    out = np.empty((polys, len(qs), N), dtype=np.uint64)
    for i, q in enumerate(qs):
        out[:, i, :] = np.mod(x, np.int64(q)).astype(np.uint64) if q < (1 << 63) else x.astype(np.uint64)
    return out
