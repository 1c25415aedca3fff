"""Driver for ncu captures of the keyswitch kernels: BASELINE config 4 (N=2^14, 8x60-bit primes, base-4 digits, D=241) at
batch 1 and 8, and config 3's rotation keyswitch (N=2^15, 60+9x40+special 60, CRT digits, ModulusRaised) at batch 64.
    python tools/prof_keyswitch.py c4 1 | c4 8 | c3 64     (prints per-class CUDA-event times as well)
    python tools/prof_keyswitch.py c4s 1                   one-prime residue shard of config 4 (what each of 8 GPUs runs)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import toyfhe_b200 as T

which, B = sys.argv[1], int(sys.argv[2])
rng = np.random.default_rng(0)

def rnd(qs, N, shape):
    out = np.empty(shape + (len(qs), N), dtype=np.uint64)
    for i, q in enumerate(qs):
        out[..., i, :] = rng.integers(0, q, size=shape + (N,), dtype=np.uint64)
    return out

if which == "c4s":
    N, w = 1 << 14, 2
    qs, psis = T.prime_chain(N, [60] * 8)
    ctx, shard = T.Context(N, qs, psis), T.Context(N, qs[:1], psis[:1])
    D = T.ndigits(qs, w)
    key = shard.sample_uniform(9, 1, (D, 2))
    ct = ctx.to_device(rnd(qs, N, (B, 3)))
    out = shard.empty((B, 2, 1, N))
    run = lambda: ctx.keyswitch_shard(shard, 0, key, ct, w, out=out)
elif which == "c4":
    N, w = 1 << 14, 2
    qs, psis = T.prime_chain(N, [60] * 8)
    ctx = T.Context(N, qs, psis)
    D = T.ndigits(qs, w)
    key = ctx.ntt_fwd(ctx.to_device(rnd(qs, N, (D, 2))))
    ct = ctx.to_device(rnd(qs, N, (B, 3)))
    run = lambda: ctx.keyswitch(key, ct, w)
else:
    N, w = 1 << 15, 0
    qk, pk = T.prime_chain(N, [60] + [40] * 9 + [60])
    qs, psis = qk[:-1], pk[:-1]
    ctx, ext = T.Context(N, qs, psis), T.Context(N, qk, pk)
    key = ext.ntt_fwd(ext.to_device(rnd(qk, N, (len(qk), 2))))
    ct = ctx.to_device(rnd(qs, N, (B, 2)))
    g = pow(3, 2 * N - 128, 2 * N)
    run = lambda: ctx.keyswitch(key, ctx.galois(ct, g), w, ext=ext)
for _ in range(2):
    run()
torch.cuda.synchronize()
T.profile_read(reset=True)
T.profile_enable(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
it = 5
for _ in range(it):
    run()
e1.record()
torch.cuda.synchronize()
T.profile_enable(False)
prof = T.profile_read(reset=True)
tot = e0.elapsed_time(e1) / it
print(f"{which} batch {B}: {tot:.3f} ms per call; per class (ms per call, launches per call): " +
      ", ".join(f"{k} {ms / it:.3f} ({cnt // it})" for k, (cnt, ms) in prof.items() if cnt))
