"""Per-kernel-class device time (tfb_profile_*, CUDA events on the launching stream) of the keyswitch-heavy ops:
C3 rotate+keyswitch (N=2^15, CRT digits, special prime), C5 rotate+keyswitch (N=2^13), C4 relinearise (base 4)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import toyfhe_b200 as T
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench_configs as BC

def profile(name, fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    T.profile_enable(True)
    T.profile_read(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    wall = e0.elapsed_time(e1) / iters
    pr = T.profile_read(True)
    T.profile_enable(False)
    tot = sum(ms for _, ms in pr.values()) / iters
    print(f"{name}: {wall:.3f} ms per call (sum of kernels {tot:.3f} ms)")
    for k, (n, ms) in sorted(pr.items(), key=lambda kv: -kv[1][1]):
        if n:
            print(f"    {k:14s} {n // iters:3d} launches  {ms / iters:.3f} ms  {100 * ms / iters / wall:5.1f}%")

rng = np.random.default_rng(1)
for label, N, logs, B in (("C3 rotate+keyswitch N=2^15 B=8", 2 ** 15, [60] + [40] * 9, 8), ("C3 rotate+keyswitch N=2^15 B=64", 2 ** 15, [60] + [40] * 9, 64),
                          ("C5 rotate+keyswitch N=2^13 B=64", 2 ** 13, [60] + [40] * 5, 64)):
    qs, psis, sp, spsi = BC.ckks_chain(N, logs)
    ctx, ext, key = BC.keyswitch_setup(N, qs, psis, sp, spsi, rng)
    ct = ctx.to_device(BC.rand_res(rng, qs, N, (B, 2)))
    tmp, out = torch.empty_like(ct), torch.empty_like(ct)
    g = T.galois_element_from_steps(1, N)
    profile(label, lambda: (ctx.galois(ct, g, out=tmp), ctx.keyswitch(key, tmp, 0, ext=ext, out=out)))
N, L, w = 2 ** 14, 8, 2
qs, psis = T.prime_chain(N, [60] * L)
ctx = T.Context(N, qs, psis)
D = T.ndigits(qs, w)
key = ctx.to_device(BC.rand_res(rng, qs, N, (8, 2))).repeat((D + 7) // 8, 1, 1, 1)[:D].contiguous()
for B in (1, 4, 8):
    ct = ctx.to_device(BC.rand_res(rng, qs, N, (B, 3)))
    out = ctx.empty((B, 2, L, N))
    profile(f"C4 relinearise base 4 (D={D}) N=2^14 L=8 B={B}", lambda: ctx.keyswitch(key, ct, w, out=out), iters=3)
