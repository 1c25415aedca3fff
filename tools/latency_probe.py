"""Latency of one BFV ciphertext multiply and one NTT call at small batch sizes (device-resident, CUDA events)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import toyfhe_b200 as T
import bench

qs, psis, qb, psib = bench.rings()
cq, cb = T.Context(bench.N_RING, qs, psis), T.Context(bench.N_RING, qb, psib)
rng = np.random.default_rng(0)
for B in (1, 2, 4, 8, 16, 32, 64, 128):
    c1 = cq.to_device(bench.rand_ct(rng, qs, (B, 2)))
    c2 = cq.to_device(bench.rand_ct(rng, qs, (B, 2)))
    out = cq.empty((B, 3, bench.L_Q, bench.N_RING))
    tmp = torch.empty_like(c1)
    def timeit(fn, it=50):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(it):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / it
    tm = timeit(lambda: cq.bfv_mul(cb, bench.T_PLAIN, c1, c2, out=out))
    tn = timeit(lambda: cq.ntt_fwd(c1, out=tmp))
    print(f"batch {B:4d}: BFV multiply {tm * 1e3:8.1f} us per call ({B / tm * 1e3:8.0f} /s)   forward NTT of {2 * B} polys {tn * 1e3:7.1f} us", flush=True)
