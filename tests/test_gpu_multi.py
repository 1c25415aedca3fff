"""Ciphertext-parallel sharding on real GPUs (NCCL): the sharded BFV multiply over 2 ranks equals the
oracle.  Skipped on boxes with fewer than 2 GPUs (the CPU/gloo version is tests/test_sharding_gloo.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, batch, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        import toyfhe_b200 as T
        from toyfhe_b200 import sharding as S
        from oracle import c_oracle as CO
        N, L, Lb, t = 1024, 3, 7, 65537
        allq, allpsi = T.prime_chain(N, [60] * (L + Lb))
        qs, psis, qb, psib = allq[:L], allpsi[:L], allq[L:], allpsi[L:]
        cq, cb = T.Context(N, qs, psis, device=rank), T.Context(N, qb, psib, device=rank)
        full = (None, None)
        if rank == 0:
            rng = np.random.default_rng(11)
            a = np.empty((batch, 2, L, N), dtype=np.uint64)
            b = np.empty_like(a)
            for i, q in enumerate(qs):
                a[:, :, i, :] = rng.integers(0, q, size=(batch, 2, N), dtype=np.uint64)
                b[:, :, i, :] = rng.integers(0, q, size=(batch, 2, N), dtype=np.uint64)
            full = (cq.to_device(a), cq.to_device(b))
        tail = (2, L, N)
        before = T.kernel_launches()
        res = S.sharded_apply(lambda x, y: cq.bfv_mul(cb, t, x, y), batch, full, (tail, tail), device=f"cuda:{rank}")
        torch.cuda.synchronize()
        assert T.kernel_launches() > before        # every rank ran engine kernels on its shard
        if rank == 0:
            want = CO.bfv_mul(CO.Rns(N, qs, psis), CO.Rns(N, qb, psib), t, a, b)
            assert np.array_equal(T.Context.to_host(res), want)
            np.save(os.path.join(outdir, "ok.npy"), np.array([1]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_sharded_bfv_mul_two_gpus(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mp.spawn(_worker, args=(2, _free_port(), 5, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok.npy")


def _ks_worker(rank, world, port, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        import toyfhe_b200 as T
        from toyfhe_b200 import sharding as S
        N, L, w = 1024, 4, 2
        qs, psis = T.prime_chain(N, [60] * L)
        ctx = T.Context(N, qs, psis, device=rank)
        rng = np.random.default_rng(21)            # replicated ciphertext and key

        def rnd(shape):
            out = np.empty(shape + (L, N), dtype=np.uint64)
            for i, q in enumerate(qs):
                out[..., i, :] = rng.integers(0, q, size=shape + (N,), dtype=np.uint64)
            return out

        D = T.ndigits(qs, w)
        key, ct = rnd((D, 2)), rnd((1, 3))
        key_dual = ctx.ntt_fwd(ctx.to_device(key))
        d_ct = ctx.to_device(ct)
        lo, hi = S.shard_range(L, rank, world)
        shard = T.Context(N, qs[lo:hi], psis[lo:hi], device=rank)
        krows = S.key_rows_for_shard(key_dual, lo, hi)
        res = S.keyswitch_residue_sharded(lambda a, b: ctx.keyswitch_shard(shard, a, krows, d_ct, w), L)
        whole = ctx.keyswitch(key_dual, d_ct, w)
        assert torch.equal(res, whole)
        # the same with the exchange fused into the epilogue kernel (peer stores over NVLink, IPC-mapped buffers): three
        # calls in a row exercise both result slots and the epoch flags
        x = S.open_peer_exchange(ctx, L, 2)
        try:
            for it in range(3):
                ct_i = ctx.to_device(rnd((1 + it % 2, 3)))
                got = S.keyswitch_residue_sharded_push(ctx, shard, lo, krows, ct_i, w, x)
                assert torch.equal(got, ctx.keyswitch(key_dual, ct_i, w))
            assert not x.timed_out()
            dist.barrier()
        finally:
            x.close()
        np.save(os.path.join(outdir, f"ks{rank}.npy"), np.array([1]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_residue_sharded_keyswitch_two_gpus(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mp.spawn(_ks_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ks0.npy") and os.path.exists(tmp_path / "ks1.npy")
