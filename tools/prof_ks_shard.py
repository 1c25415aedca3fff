"""Where does the residue-sharded base-4 keyswitch of ONE ciphertext spend its time on a rank that owns Ls of the 8 primes?
One GPU, no collective: per-class kernel times (engine profile), the eager call, and the same call replayed from a CUDA
graph (no launch gaps).  The difference to the G-GPU figure of tools/bench_keyswitch_sharded.py is the exchange.
    python tools/prof_ks_shard.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import toyfhe_b200 as T

N, L, w = 1 << 14, 8, 2
qs, psis = T.prime_chain(N, [60] * L)
ctx = T.Context(N, qs, psis)
D = T.ndigits(qs, w)
for B in (1, 8):
    ct = ctx.sample_uniform(5, 0, (B, 3))
    for Ls in (1, 2, 4, 8):
        shard = T.Context(N, qs[:Ls], psis[:Ls])
        krows = shard.sample_uniform(9, 1, (D, 2))
        out = shard.empty((B, 2, Ls, N))
        run = lambda: ctx.keyswitch_shard(shard, 0, krows, ct, w, out=out)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        T.profile_read(reset=True); T.profile_enable(True)
        for _ in range(5):
            run()
        torch.cuda.synchronize()
        T.profile_enable(False)
        prof = T.profile_read(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        it = 20
        e0.record()
        for _ in range(it):
            run()
        e1.record(); torch.cuda.synchronize()
        eager = e0.elapsed_time(e1) / it
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            run()
        side.synchronize()
        with torch.cuda.graph(g, stream=side):
            run()
        g.replay(); torch.cuda.synchronize()
        e0.record()
        for _ in range(it):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        graph = e0.elapsed_time(e1) / it
        print(f"B={B} Ls={Ls}: eager {eager*1e3:.0f} us, graph {graph*1e3:.0f} us; classes (us): " +
              ", ".join(f"{k} {ms / 5 * 1e3:.0f}" for k, (cnt, ms) in prof.items() if cnt), flush=True)
        del krows, shard
