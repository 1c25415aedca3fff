"""BASELINE.json configs[2]: CKKS diagonal-method matrix product (test/ckks_matmul.jl:8-44 scaled up) with rescale.

Shape: N = 2^15, chain q0 (60-bit) + 9 x 40-bit + special (60-bit) built like examples/encrypted_mnist/infer.jl:97-105,
ModulusRaised(CKKSParams(R, 0, 3.2)), scale 2^40; x = a d x d matrix packed in the N/2 = d * (N/2/d) slots
(slot c + cols * r = entry [r][c], the layout of test/ckks_matmul.jl with cols = N/2/d columns), W a d x d plaintext matrix:

    result = repeat(diag(W), cols) .* x;  for k = 2..d: rotated = rotate(gk, rotated);
    result += repeat(diag(circshift(W, (0, k-1))), cols) .* rotated                      test/ckks_matmul.jl:34-42
    rescale

i.e. d - 1 rotations + keyswitches, d plaintext-vector multiplies, 1 rescale per ciphertext (d = 128: 127 / 128 / 1).
Checked against W @ X in float64 (atol 1e-5 as in the reference test).

    python -m workloads.ckks_matmul [--batch B] [--d 128] [--logn 15]      # one JSON line
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

import toyfhe_b200 as T  # noqa: E402
from workloads.ckks_batch import CtBatch, Graphed, MatDiagonals, Pipeline, decrypt_slots, diag_matmul  # noqa: E402

SCALE = float(2 ** 40)


def chain_ring(N: int, n40: int) -> T.NegacyclicRing:
    qs, psis = T.prime_chain(N, [60] + [40] * n40 + [60])
    return T.NegacyclicRing(N, qs=qs, psis=psis)


def diag_vectors(W: np.ndarray, cols: int):
    d = W.shape[0]
    return [np.repeat(np.array([W[l, (l - k) % d] for l in range(d)]), cols) for k in range(d)]


def run(batch: int, d: int, N: int, n40: int = 9, reps: int = 1, seed: int = 0, check: bool = True, graph: bool = True) -> dict:
    cols = (N // 2) // d
    assert cols * d == N // 2
    R = chain_ring(N, n40)
    params = T.ModulusRaised(T.CKKSParams(R, 0, 3.2))
    s = T.Sampler(seed + 1, device=True)
    kp = T.keygen(s, params)
    gk = T.keygen_galois(s, kp.priv, steps=cols)
    pipe = Pipeline(params)
    rng = np.random.default_rng(seed)
    X = rng.uniform(-1, 1, size=(d, cols))
    W = rng.normal(0, 1 / np.sqrt(d), size=(d, d))
    c = T.encrypt(s, kp, T.CKKSEncoding(SCALE, X.reshape(-1).astype(np.complex128)))
    x1 = CtBatch.from_ciphertexts(pipe, [c])
    diags = MatDiagonals(pipe.level(0), x1.scale, diag_vectors(W, cols))
    want = W @ X
    out = {"N": N, "d": d, "primes": R.L, "batch": batch, "rotations": d - 1, "plaintext_multiplies": d, "rescales": 1}
    if check:
        got = np.real(decrypt_slots(kp, diag_matmul(x1, gk, diags).rescale())).reshape(d, cols)
        out["max_abs_err"] = float(np.max(np.abs(got - want)))
    if batch > 0:
        xb = x1.replicate(batch)
        launches0 = T.kernel_launches()
        diag_matmul(xb, gk, diags).rescale()
        torch.cuda.synchronize()
        launches = T.kernel_launches() - launches0
        op = (lambda: diag_matmul(xb, gk, diags).rescale())
        fwd = Graphed(op) if graph else op
        fwd()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            res = fwd()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        last = np.real(decrypt_slots(kp, res, batch - 1)).reshape(d, cols)
        out.update({"ms_per_batch": ms, "matmuls_per_s": batch / (ms * 1e-3), "rotations_per_s": batch * (d - 1) / (ms * 1e-3),
                    "cuda_graph": bool(graph), "kernel_launches_per_batch": launches,
                    "last_of_batch_max_abs_err": float(np.max(np.abs(last - want)))})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--d", type=int, default=128)
    ap.add_argument("--logn", type=int, default=15)
    ap.add_argument("--n40", type=int, default=9)
    ap.add_argument("--reps", type=int, default=1)
    a = ap.parse_args()
    t0 = time.time()
    r = run(a.batch, a.d, 1 << a.logn, a.n40, a.reps)
    r["wall_s"] = time.time() - t0
    print(json.dumps({"workload": "ckks_matmul (test/ckks_matmul.jl:30-44 at d x d)", **r}))


if __name__ == "__main__":
    main()
