"""Workload pipelines (BASELINE.json configs 3 and 5) at reduced and at full size: every operation on the device,
one decrypt at the end, compared with the plaintext computation in float64 (tolerances of the reference's CKKS tests)."""
import numpy as np
import pytest

import toyfhe_b200 as T
from workloads import ckks_matmul, mnist

pytestmark = pytest.mark.gpu


def test_mul_plain_broadcast_and_accumulate():
    N = 256
    qs, psis = T.prime_chain(N, [60, 40, 40])
    ctx = T.Context(N, qs, psis)
    rng = np.random.default_rng(3)
    rnd = lambda shape: np.stack([rng.integers(0, q, size=shape + (N,), dtype=np.uint64) for q in qs], axis=-2)
    a, p, acc = rnd((5, 2)), rnd(()), rnd((5, 2))
    want = np.empty_like(a)
    for i, q in enumerate(qs):
        want[..., i, :] = (a[..., i, :].astype(object) * p[i].astype(object) % q).astype(np.uint64)
    got = ctx.to_host(ctx.mul_plain(ctx.to_device(a), ctx.to_device(p)))
    assert np.array_equal(got, want)
    d = ctx.to_device(acc)
    ctx.mul_plain(ctx.to_device(a), ctx.to_device(p), out=d, accumulate=True)
    for i, q in enumerate(qs):
        want[..., i, :] = ((want[..., i, :].astype(object) + acc[..., i, :].astype(object)) % q).astype(np.uint64)
    assert np.array_equal(ctx.to_host(d), want)


def test_lincomb_matches_scalar_mul_and_add():
    N = 256
    qs, psis = T.prime_chain(N, [60, 60, 40])
    ctx = T.Context(N, qs, psis)
    rng = np.random.default_rng(9)
    J, Cn = 7, 3
    stacked = np.stack([np.stack([rng.integers(0, q, size=(2, 2, N), dtype=np.uint64) for q in qs], axis=-2) for _ in range(J)])   # [J][2][2][L][N]
    wint = rng.integers(-2 ** 45, 2 ** 45, size=(Cn, J))
    w = np.array([[[int(v) % q for q in qs] for v in row] for row in wint], dtype=np.uint64)
    got = ctx.to_host(ctx.lincomb(ctx.to_device(stacked), ctx.to_device(w)))
    want = np.zeros((Cn,) + stacked.shape[1:], dtype=object)
    for c in range(Cn):
        for j in range(J):
            for i, q in enumerate(qs):
                want[c, :, :, i, :] = (want[c, :, :, i, :] + stacked[j, :, :, i, :].astype(object) * int(w[c, j, i])) % q
    assert np.array_equal(got, want.astype(np.uint64))
    with pytest.raises(T.EngineError):
        ctx.lincomb(ctx.to_device(np.zeros((64, 1, 1, 3, N), dtype=np.uint64)), ctx.to_device(np.zeros((1, 64, 3), dtype=np.uint64)))


def test_add_plain_first_component():
    N = 128
    qs, psis = T.prime_chain(N, [60, 40])
    ctx = T.Context(N, qs, psis)
    rng = np.random.default_rng(4)
    rnd = lambda shape: np.stack([rng.integers(0, q, size=shape + (N,), dtype=np.uint64) for q in qs], axis=-2)
    ct, p = rnd((5, 3)), rnd(())
    d = ctx.to_device(ct)
    ctx.add_plain_first(d, ctx.to_device(p))
    want = ct.copy()
    for i, q in enumerate(qs):
        want[:, 0, i, :] = ((ct[:, 0, i, :].astype(object) + p[i].astype(object)) % q).astype(np.uint64)
    assert np.array_equal(ctx.to_host(d), want)


def test_mnist_pipeline_graph_replay_equals_eager():
    """the pipeline captured into a CUDA graph and replayed gives the eager run's ciphertext bit for bit"""
    import torch
    P = mnist.MnistPipeline(N=512, m=16, seed=3)
    C = [c.replicate(2) for c in P.encrypt_inputs(mnist.make_inputs(5, 16, P.n_img))]
    eager = P.forward(C).ct.clone()
    torch.cuda.synchronize()
    for _ in range(2):
        got = P.forward_graphed(C)
    torch.cuda.synchronize()
    assert bool((got.ct == eager).all())


def test_ckks_matmul_reference_shape():
    """test/ckks_matmul.jl: N = 32, 4 x 4, atol 1e-5 (here with the special-prime CRT keyswitch of the workload)"""
    r = ckks_matmul.run(batch=3, d=4, N=32, n40=1)
    assert r["max_abs_err"] < 1e-5 and r["last_of_batch_max_abs_err"] < 1e-5


def test_ckks_matmul_small():
    r = ckks_matmul.run(batch=2, d=16, N=1024, n40=3)
    assert r["max_abs_err"] < 1e-5 and r["last_of_batch_max_abs_err"] < 1e-5


def test_ckks_matmul_full_size():
    """BASELINE configs[2]: N = 2^15, 10-level chain + special prime, 128 x 128: 127 rotations, 128 multiplies, 1 rescale"""
    r = ckks_matmul.run(batch=2, d=128, N=2 ** 15, n40=9)
    assert r["primes"] == 11 and r["rotations"] == 127
    assert r["max_abs_err"] < 1e-5 and r["last_of_batch_max_abs_err"] < 1e-5


def test_mnist_pipeline_small():
    r = mnist.run(batch=2, m=16, N=512)
    assert r["max_abs_err"] < 1e-3 * max(1.0, r["max_abs_value"]) and r["labels_agree"]
    assert r["last_of_batch_max_abs_err"] < 1e-3 * max(1.0, r["max_abs_value"])


def test_mnist_pipeline_full_size():
    """BASELINE configs[4] shape: N = 2^13, (60, 5 x 40, 60) primes, 64 images per pipeline, 315 rotations"""
    r = mnist.run(batch=2, m=64, N=2 ** 13)
    assert r["images_per_pipeline"] == 64
    assert r["max_abs_err"] < 1e-3 * max(1.0, r["max_abs_value"]) and r["labels_agree"]
    assert r["last_of_batch_max_abs_err"] < 1e-3 * max(1.0, r["max_abs_value"])
