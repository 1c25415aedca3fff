"""Ciphertext-parallel sharding across GPUs (one process per GPU, torch.distributed).

The reference is single-process; its "batches" are plain `map`s over arrays of
independent ciphertexts (examples/encrypted_mnist/infer.jl:120-137), and every
ring operation is independent per ciphertext (and per RNS prime, crt.jl:250-254).
So a batch shards over ranks with NO data-path collective: each rank runs the
engine on its contiguous slice, context tables and evaluation keys are replicated.
NCCL (or gloo in the CPU tests) is used only to scatter a batch that lives on one
rank and to gather the results back.

Layout: batches are int64/uint64 tensors `[batch, ...]` (engine layout
`[batch][components][L][N]`); sharding is along dim 0.
"""
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_units: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [start, stop) of `n_units` independent units for `rank`
    (the first n_units % world ranks get one extra unit)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    if n_units < 0:
        raise ValueError("n_units must be >= 0")
    base, extra = divmod(n_units, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_sizes(n_units: int, world: int) -> List[int]:
    return [shard_range(n_units, r, world)[1] - shard_range(n_units, r, world)[0] for r in range(world)]


def _world(group=None) -> Tuple[int, int]:
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def scatter_batch(full: Optional[torch.Tensor], batch: int, tail_shape: Sequence[int], dtype=torch.int64,
                  device=None, src: int = 0, group=None) -> torch.Tensor:
    """Rank `src` holds `full` = [batch, *tail_shape]; every rank gets its shard_range slice.
    Ragged shards are sent point-to-point (scatter needs equal sizes)."""
    rank, world = _world(group)
    lo, hi = shard_range(batch, rank, world)
    if world == 1:
        return full[lo:hi].contiguous()
    mine = torch.empty((hi - lo, *tail_shape), dtype=dtype, device=device)
    if rank == src:
        reqs = []
        for r in range(world):
            a, b = shard_range(batch, r, world)
            if r == src:
                mine.copy_(full[a:b])
            elif b > a:
                reqs.append(dist.isend(full[a:b].contiguous(), dst=r, group=group))
        for q in reqs:
            q.wait()
    elif hi > lo:
        dist.recv(mine, src=src, group=group)
    return mine


def gather_batch(local: torch.Tensor, batch: int, dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """Inverse of scatter_batch: rank `dst` returns [batch, ...], the others None."""
    rank, world = _world(group)
    if world == 1:
        return local
    if rank == dst:
        out = torch.empty((batch, *local.shape[1:]), dtype=local.dtype, device=local.device)
        for r in range(world):
            a, b = shard_range(batch, r, world)
            if r == dst:
                out[a:b].copy_(local)
            elif b > a:
                dist.recv(out[a:b], src=r, group=group)
        return out
    if local.shape[0] > 0:
        dist.send(local.contiguous(), dst=dst, group=group)
    return None


def sharded_apply(op: Callable[..., torch.Tensor], batch: int, inputs: Sequence[Optional[torch.Tensor]],
                  tail_shapes: Sequence[Sequence[int]], device=None, root: int = 0, group=None) -> Optional[torch.Tensor]:
    """scatter -> `op` on the local shard -> gather.  `op` is an engine call such as
    `lambda a, b: ctx.bfv_mul(big, t, a, b)`; ranks with an empty shard skip it."""
    rank, world = _world(group)
    shards = [scatter_batch(x, batch, ts, device=device, src=root, group=group) for x, ts in zip(inputs, tail_shapes)]
    if shards[0].shape[0] > 0:
        res = op(*shards)
        meta = torch.tensor(list(res.shape[1:]), dtype=torch.int64, device=res.device)
    else:
        res, meta = None, None
    if world > 1:
        # every rank needs the result's tail shape to build an empty shard / the gather buffer
        lo0, hi0 = shard_range(batch, 0, world)
        holder = 0 if hi0 > lo0 else None
        if holder is None:
            raise ValueError("empty batch")
        n = torch.zeros(1, dtype=torch.int64, device=shards[0].device)
        if rank == holder:
            n[0] = meta.numel()
        dist.broadcast(n, src=holder, group=group)
        if rank != holder:
            meta = torch.zeros(int(n.item()), dtype=torch.int64, device=shards[0].device)
        dist.broadcast(meta, src=holder, group=group)
        if res is None:
            res = torch.empty((0, *[int(v) for v in meta.tolist()]), dtype=shards[0].dtype, device=shards[0].device)
    return gather_batch(res, batch, dst=root, group=group)


# --------------------------------------------------------------------------- residue-parallel keyswitch
# BASELINE config 4 ("residues sharded 2/4/8 GPU"): ONE ciphertext (or a small batch) is
# key-switched by all ranks together.  Rank r owns the contiguous primes shard_range(L, r, world); the ciphertext is
# replicated (2-3 MiB), the evaluation key -- the large object, D*2*L*N words -- is split by prime rows so each rank
# holds and streams only its part.  Every rank extracts the digits of the whole integer itself (rlwe_she.jl:328-337
# needs all residues of a coefficient; recomputing them is cheaper than exchanging them), transforms and accumulates
# them under its own primes, and the result rows are assembled with ONE all-gather -- the only data-path collective.

def key_rows_for_shard(key_dual: torch.Tensor, lo: int, hi: int) -> torch.Tensor:
    """key_dual [D][2][L][N] -> the contiguous copy [D][2][hi-lo][N] a rank keeps"""
    return key_dual[:, :, lo:hi, :].contiguous()


_GATHER_BUF = {}


def allgather_prime_rows(local: torch.Tensor, L: int, group=None) -> torch.Tensor:
    """local [..., Ls, N] (this rank's prime rows, ragged over ranks) -> [..., L, N] on every rank.
    Equal shard sizes (L divisible by the world size: 8 primes on 2/4/8 GPUs): ONE all_gather_into_tensor into a cached
    [world][...] buffer and one strided copy into the [..., L, N] result -- no padding, no per-rank tensors, no cat."""
    rank, world = _world(group)
    if world == 1:
        return local
    sizes = shard_sizes(L, world)
    lead, N = tuple(local.shape[:-2]), local.shape[-1]
    if len(set(sizes)) == 1 and hasattr(dist, "all_gather_into_tensor"):
        Ls = sizes[0]
        key = (world, lead, Ls, N, local.dtype, str(local.device))
        buf = _GATHER_BUF.get(key)
        if buf is None:
            buf = _GATHER_BUF[key] = torch.empty((world,) + lead + (Ls, N), dtype=local.dtype, device=local.device)
        try:
            dist.all_gather_into_tensor(buf, local.contiguous(), group=group)
            out = torch.empty(lead + (L, N), dtype=local.dtype, device=local.device)
            # out[..., r*Ls + i, :] = buf[r, ..., i, :]
            out.view(lead + (world, Ls, N)).copy_(buf.movedim(0, len(lead)))
            return out
        except (RuntimeError, NotImplementedError):
            pass                                        # backend without the fused collective: the general path below
    mx = max(sizes)
    pad = torch.zeros(lead + (mx, N), dtype=local.dtype, device=local.device)
    pad[..., : local.shape[-2], :] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[..., :s, :] for p, s in zip(parts, sizes) if s > 0], dim=-2).contiguous()


def keyswitch_residue_sharded(shard_op: Callable[[int, int], torch.Tensor], L: int, group=None) -> torch.Tensor:
    """`shard_op(lo, hi)` computes rows lo..hi-1 of the key-switched ciphertext ([B][2][hi-lo][N]); on a GPU it is
    `lambda lo, hi: ctx.keyswitch_shard(shard_ctx, lo, key_rows, ct, w)`.  Returns [B][2][L][N] on every rank."""
    rank, world = _world(group)
    lo, hi = shard_range(L, rank, world)
    if hi <= lo:
        raise ValueError("more ranks than RNS primes: residue sharding needs world <= L")
    return allgather_prime_rows(shard_op(lo, hi), L, group=group)


def open_peer_exchange(ctx, L: int, max_batch: int, group=None):
    """One PeerExchange per rank for results of up to max_batch ciphertexts ([B][2][L][N]), the IPC handles swapped with
    all_gather_object.  Collective: every rank of the group calls it.  Either every rank ends up with a fully attached
    exchange or every rank raises (a rank whose allocation or IPC mapping fails does not leave the others waiting)."""
    from .engine import PeerExchange
    rank, world = _world(group)
    x, err, handle = None, None, None
    try:
        x = PeerExchange(ctx, rank, world, max_batch * 2 * L * ctx.N)
        handle = x.handle()
    except Exception as e:                      # reported to the peers through the missing handle
        err = e
    handles = [None] * world
    dist.all_gather_object(handles, handle, group=group)
    ok = err is None and all(h is not None for h in handles)
    if ok:
        try:
            x.attach(handles)
        except Exception as e:                  # e.g. no peer access between two of the GPUs
            err, ok = e, False
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=f"cuda:{ctx.device}")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if int(flag.item()) == 0:
        dist.barrier(group=group)               # nobody frees a buffer a peer may still be mapping
        if x is not None:
            x.close()
        raise RuntimeError(f"peer exchange unavailable on at least one rank (this rank: {err!r})")
    dist.barrier(group=group)
    return x


def keyswitch_residue_sharded_push(ctx, shard, lo: int, key_rows: torch.Tensor, ct: torch.Tensor, w: int, xchg) -> torch.Tensor:
    """keyswitch_residue_sharded with the gather fused into the epilogue kernel (tfb_keyswitch_shard_push): no collective
    call on the data path -- the rows travel as peer stores of the kernel that computes them.  [B][2][L][N] on every rank, a
    view of the exchange slot (copy it if it has to outlive the next two calls)."""
    return ctx.keyswitch_shard_push(shard, lo, key_rows, ct, w, xchg)
