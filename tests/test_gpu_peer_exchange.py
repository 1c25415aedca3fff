"""The sharded keyswitch whose epilogue kernel pushes its rows into every rank's result buffer (tfb_keyswitch_shard_push,
BASELINE config 4 "residues sharded over GPUs").  Here the ranks are exchanges of ONE process on ONE GPU, each with its own
contexts and its own stream -- the kernels, the flag protocol and the two-slot reuse are the ones the multi-GPU run uses
(tests/test_gpu_multi.py maps the buffers through CUDA IPC on two GPUs instead)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

import toyfhe_b200 as T
from toyfhe_b200 import sharding as S


def _rnd(rng, qs, N, shape):
    out = np.empty(shape + (len(qs), N), dtype=np.uint64)
    for i, q in enumerate(qs):
        out[..., i, :] = rng.integers(0, q, size=shape + (N,), dtype=np.uint64)
    return out


@pytest.mark.parametrize("N,logqs,w,world", [(4096, [60] * 4, 2, 2), (16384, [60] * 8, 2, 8), (1024, [60, 60, 40], 7, 3)])
def test_push_exchange_equals_whole_keyswitch(N, logqs, w, world):
    qs, psis = T.prime_chain(N, logqs)
    L = len(qs)
    ref = T.Context(N, qs, psis)
    rng = np.random.default_rng(N + world)
    D = T.ndigits(qs, w)
    key_dual = ref.ntt_fwd(ref.to_device(_rnd(rng, qs, N, (D, 2))))
    ranks = []
    for r in range(world):
        lo, hi = S.shard_range(L, r, world)
        ctx = T.Context(N, qs, psis)                       # one whole-ring context per rank, as in one process per GPU
        ranks.append(dict(ctx=ctx, lo=lo, shard=T.Context(N, qs[lo:hi], psis[lo:hi]), krows=S.key_rows_for_shard(key_dual, lo, hi),
                          x=T.PeerExchange(ctx, r, world, 2 * 2 * L * N), stream=torch.cuda.Stream()))
    for rk in ranks:
        rk["x"].attach_local([o["x"] for o in ranks])
    warm = ref.to_device(_rnd(rng, qs, N, (2, 3)))
    for rk in ranks:   # sizes every context's scratch now: a first-use cudaMalloc synchronises the DEVICE, which here (all ranks on
        rk["ctx"].keyswitch_shard(rk["shard"], rk["lo"], rk["krows"], warm, w)   # one GPU) would wait on a peer's spinning kernel
    torch.cuda.synchronize()
    try:
        for it in range(4):                                # both slots twice; batch 1 and 2
            B = 1 + it % 2
            ct = ref.to_device(_rnd(rng, qs, N, (B, 3)))
            whole = ref.keyswitch(key_dual, ct, w)
            torch.cuda.synchronize()
            outs = []
            for rk in ranks:                               # asynchronous launches, one stream per rank: the waits overlap
                with torch.cuda.stream(rk["stream"]):
                    outs.append(rk["ctx"].keyswitch_shard_push(rk["shard"], rk["lo"], rk["krows"], ct, w, rk["x"]))
            torch.cuda.synchronize()
            for rk, o in zip(ranks, outs):
                assert not rk["x"].timed_out()
                assert torch.equal(o, whole)
    finally:
        torch.cuda.synchronize()
        for rk in ranks:
            rk["x"].close()


def test_push_exchange_gives_up_when_a_peer_never_arrives():
    N, w = 4096, 2
    qs, psis = T.prime_chain(N, [60, 60])
    ctx = T.Context(N, qs, psis)
    other = T.Context(N, qs, psis)
    x0, x1 = T.PeerExchange(ctx, 0, 2, 2 * 2 * N), T.PeerExchange(other, 1, 2, 2 * 2 * N)
    x0.attach_local([x0, x1]); x1.attach_local([x0, x1])
    try:
        rng = np.random.default_rng(3)
        D = T.ndigits(qs, w)
        key_dual = ctx.ntt_fwd(ctx.to_device(_rnd(rng, qs, N, (D, 2))))
        ct = ctx.to_device(_rnd(rng, qs, N, (1, 3)))
        shard = T.Context(N, qs[:1], psis[:1])
        ctx.keyswitch_shard_push(shard, 0, S.key_rows_for_shard(key_dual, 0, 1), ct, w, x0)   # rank 1 never calls
        assert x0.timed_out()                              # after the 2 s bound, not a hung GPU
        with pytest.raises(T.EngineError):
            ctx.keyswitch_shard_push(shard, 0, S.key_rows_for_shard(key_dual, 0, 1), ctx.to_device(_rnd(rng, qs, N, (3, 3))), w, x0)   # larger than the slot
    finally:
        x0.close(); x1.close()
