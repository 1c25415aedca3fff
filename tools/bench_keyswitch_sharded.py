"""BASELINE config 4: BFV relinearisation keyswitch, N=2^14, 8x60-bit primes, base-4 digits (D=241), ONE ciphertext
key-switched by all ranks together with the RNS residues sharded over the GPUs (toyfhe.jl_b200/sharding.py).
Run:  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P tools/bench_keyswitch_sharded.py [out.json]
(G = 1 works without torchrun).  Time = max over ranks, CUDA events, shard kernels + the all-gather of the result rows."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import toyfhe_b200 as T
from toyfhe_b200 import sharding as S

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
N, L, w = 1 << 14, 8, 2
qs, psis = T.prime_chain(N, [60] * L)
ctx = T.Context(N, qs, psis, device=local)
D = T.ndigits(qs, w)
rng = np.random.default_rng(7)

def rnd(shape):
    out = np.empty(shape + (L, N), dtype=np.uint64)
    for i, q in enumerate(qs):
        out[..., i, :] = rng.integers(0, q, size=shape + (N,), dtype=np.uint64)
    return out

res = {"config": f"N=2^14, L=8x60-bit, relin_window=2 (D={D} digit polynomials), residues sharded over {world} GPU(s)", "n_gpus": world}
lo, hi = S.shard_range(L, rank, world)
shard = T.Context(N, qs[lo:hi], psis[lo:hi], device=local)
key_dual = ctx.ntt_fwd(ctx.to_device(rnd((D, 2))))           # every rank builds the same key, keeps only its rows
krows = S.key_rows_for_shard(key_dual, lo, hi)
del key_dual
xchg = S.open_peer_exchange(ctx, L, 8) if world > 1 else None        # result slots mapped into every rank (CUDA IPC over NVLink)
for B, mode in ((1, "allgather"), (8, "allgather"), (1, "push"), (8, "push")):
    if mode == "push" and xchg is None:
        continue
    ct = ctx.to_device(rnd((B, 3)))
    if mode == "allgather":   # shard kernels, then ONE NCCL all-gather of the result rows
        fn = lambda: S.keyswitch_residue_sharded(lambda a, b: ctx.keyswitch_shard(shard, a, krows, ct, w), L)
    else:                     # the epilogue kernel stores its rows into every rank's slot and waits for the peers' flags: no collective
        fn = lambda: S.keyswitch_residue_sharded_push(ctx, shard, lo, krows, ct, w, xchg)
    for _ in range(3):
        out = fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    it = 10
    e0.record()
    for _ in range(it):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / it], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    if mode == "push":
        same = torch.equal(out, S.keyswitch_residue_sharded(lambda a, b: ctx.keyswitch_shard(shard, a, krows, ct, w), L))
        ok = torch.tensor([int(same and not xchg.timed_out())], device=f"cuda:{local}")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        assert int(ok.item()) == 1, "push exchange result differs from the all-gather path"
    res[f"batch{B}" + ("_push" if mode == "push" else "")] = {"ms_per_call": round(ms, 4), "keyswitches_per_s": round(B / ms * 1e3, 1), "key_bytes_per_rank": int(krows.numel() * 8)}
    if rank == 0:
        print(f"G={world} batch {B} [{mode}]: {ms:.3f} ms per call = {B / ms * 1e3:.0f} keyswitches/s (key rows per rank: {krows.numel() * 8 / 2**20:.0f} MiB)", flush=True)
if rank == 0 and len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
if world > 1:
    dist.barrier()
    xchg.close()
    dist.destroy_process_group()
