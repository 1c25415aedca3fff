"""Multi-process host logic of the ciphertext-parallel sharding (toyfhe.jl_b200/sharding.py) on CPU:
world_size 2 and 3 over gloo.  The per-shard operation here is the ORACLE's ciphertext tensor (tests may
use the oracle as the checker); on a GPU box the same code runs with `ctx.ct_tensor` / `ctx.bfv_mul` on
cuda tensors over NCCL (tests/test_gpu_multi.py, bench.py --gpus N)."""
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


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, batch, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import toyfhe_b200 as T
        from toyfhe_b200 import sharding as S
        from oracle import c_oracle as CO
        N = 64
        qs, psis = T.prime_chain(N, [50, 50, 50])
        orc = CO.Rns(N, qs, psis)
        full = None
        if rank == 0:
            rng = np.random.default_rng(5)
            a = np.empty((batch, 2, len(qs), N), dtype=np.uint64)
            b = np.empty_like(a)
            for i, q in enumerate(qs):
                a[:, :, i, :] = rng.integers(0, q, size=(batch, 2, N), dtype=np.uint64)
                b[:, :, i, :] = rng.integers(0, q, size=(batch, 2, N), dtype=np.uint64)
            full = (torch.from_numpy(a.view(np.int64)), torch.from_numpy(b.view(np.int64)))

        def op(x, y):
            r = orc.ct_tensor(x.numpy().view(np.uint64), y.numpy().view(np.uint64))
            return torch.from_numpy(np.ascontiguousarray(r).view(np.int64))

        tail = (2, len(qs), N)
        res = S.sharded_apply(op, batch, full if rank == 0 else (None, None), (tail, tail))
        lo, hi = S.shard_range(batch, rank, world)
        np.save(os.path.join(outdir, f"range{rank}.npy"), np.array([lo, hi]))
        if rank == 0:
            want = orc.ct_tensor(full[0].numpy().view(np.uint64), full[1].numpy().view(np.uint64))
            assert res is not None and res.shape == (batch, 3, len(qs), N)
            assert np.array_equal(res.numpy().view(np.uint64), want)
            np.save(os.path.join(outdir, "ok.npy"), np.array([1]))
        else:
            assert res is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,batch", [(2, 5), (2, 1), (3, 7)])
def test_sharded_ct_tensor_over_gloo(tmp_path, world, batch):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, batch, str(tmp_path)), nprocs=world, join=True)
    assert os.path.exists(tmp_path / "ok.npy")
    covered = []
    for r in range(world):
        lo, hi = np.load(tmp_path / f"range{r}.npy")
        covered += list(range(lo, hi))
    assert covered == list(range(batch))


def test_shard_range_properties():
    from toyfhe_b200 import sharding as S
    for n in (0, 1, 7, 8, 128, 4096):
        for w in (1, 2, 3, 4, 8):
            sizes = S.shard_sizes(n, w)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
            assert [S.shard_range(n, r, w)[0] for r in range(w)] == [sum(sizes[:r]) for r in range(w)]
    with pytest.raises(ValueError):
        S.shard_range(4, 2, 2)


def _ks_worker(rank, world, port, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import toyfhe_b200 as T
        from toyfhe_b200 import sharding as S
        from oracle import c_oracle as CO
        N, w = 16, 7
        qs, psis = T.prime_chain(N, [50, 50, 50])
        L = len(qs)
        orc = CO.Rns(N, qs, psis)
        rng = np.random.default_rng(3)     # same seed on every rank: the ciphertext and the key are replicated here

        def rnd(shape):
            out = np.empty(shape + (L, N), dtype=np.uint64)
            for i, q in enumerate(qs):
                out[..., i, :] = rng.integers(0, q, size=shape + (N,), dtype=np.uint64)
            return out

        ct = rnd((1, 3))
        D = CO.ndigits(qs, w)
        key = rnd((D, 2))
        # the reference algorithm on the whole ring (rlwe_she.jl:315-347): [1][2][L][N]
        w1, w2 = orc.keyswitch_accum(orc.keyswitch_digits(ct[0, 2], w), key, ct[0, 0], ct[0, 1])
        want = np.stack([w1, w2])[None]

        def shard_op(lo, hi):                                 # rows lo..hi-1 (the oracle stands in for tfb_keyswitch_shard)
            return torch.from_numpy(np.ascontiguousarray(want[:, :, lo:hi, :]).view(np.int64))

        res = S.keyswitch_residue_sharded(shard_op, L)
        assert res.shape == (1, 2, L, N)
        assert np.array_equal(res.numpy().view(np.uint64), want)
        krows = S.key_rows_for_shard(torch.from_numpy(key.view(np.int64)), *S.shard_range(L, rank, world))
        assert krows.shape == (D, 2, S.shard_sizes(L, world)[rank], N) and krows.is_contiguous()
        np.save(os.path.join(outdir, f"ok{rank}.npy"), np.array([1]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_residue_sharded_keyswitch_gather_over_gloo(tmp_path, world):
    """prime rows computed per rank (ragged: 3 primes over 2 ranks) are assembled by one all-gather"""
    mp.spawn(_ks_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}.npy") for r in range(world))
