"""Short driver for ncu captures: a few launches of each hot kernel at the headline shape."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import toyfhe_b200 as T
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
qs, psis, qb, psib = bench.rings()
cq, cb = T.Context(bench.N_RING, qs, psis), T.Context(bench.N_RING, qb, psib)
rng = np.random.default_rng(0)
c1 = cq.to_device(bench.rand_ct(rng, qs, (B, 2)))
c2 = cq.to_device(bench.rand_ct(rng, qs, (B, 2)))
out = cq.empty((B, 3, bench.L_Q, bench.N_RING))
tmp = torch.empty_like(c1)
for _ in range(3):
    cq.ntt_fwd(c1, out=tmp)
    cq.ntt_inv(tmp, out=tmp)
for _ in range(2):
    cq.bfv_mul(cb, bench.T_PLAIN, c1, c2, out=out)
torch.cuda.synchronize()
print("driver done")
