"""Times the forward / inverse NTT kernel variants at the headline shape (N=2^14, L=8)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import toyfhe_b200 as T
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
qs, psis, qb, psib = bench.rings()
cq = T.Context(bench.N_RING, qs, psis)
rng = np.random.default_rng(0)
a = cq.to_device(bench.rand_ct(rng, qs, (B, 2)))
out = torch.empty_like(a)
rows = 2 * B * len(qs)
nbytes = rows * bench.N_RING * 8 * 2

def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

for version, mode in ((1, 0), (1, 1), (2, 0), (2, 1), (3, 0), (3, 1), (3, 2)):
    if True:
        T.ntt_version(version)
        T.ntt_max_mode(mode)
        f = timeit(lambda: cq.ntt_fwd(a, out=out))
        i = timeit(lambda: cq.ntt_inv(a, out=out))
        print(f"v{version} mode {mode} ({('harvey', 'lazy', 'lazy+approx quotient')[mode]}): fwd {f:.3f} ms = {rows / f / 1e3:.2f} Mrows/s {nbytes / f / 1e6:.0f} GB/s | "
              f"inv {i:.3f} ms = {rows / i / 1e3:.2f} Mrows/s {nbytes / i / 1e6:.0f} GB/s", flush=True)
