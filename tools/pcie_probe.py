"""PCIe probe: pinned H2D / D2H alone and concurrently (32 MiB copies), then the host-buffer BFV multiply
(tfb_bfv_mul_host) at several batch sizes.  Chunk size of the pipeline: env TFB_HOST_CHUNK_MIB (read at first call)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import toyfhe_b200 as T
import bench

def ev():
    return torch.cuda.Event(enable_timing=True)

if "--copy" in sys.argv:
    n = 512 << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    ck = 32 << 20
    def h2d():
        with torch.cuda.stream(s1):
            for o in range(0, n, ck):
                d_in[o:o + ck].copy_(h_in[o:o + ck], non_blocking=True)
    def d2h():
        with torch.cuda.stream(s2):
            for o in range(0, n, ck):
                h_out[o:o + ck].copy_(d_out[o:o + ck], non_blocking=True)
    for name, fns in (("H2D alone", (h2d,)), ("D2H alone", (d2h,)), ("H2D + D2H concurrently", (h2d, d2h))):
        for _ in range(2):
            for f in fns:
                f()
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            for f in fns:
                f()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
        print(f"{name:26s} {n / dt / 1e9:6.1f} GB/s per direction", flush=True)

qs, psis, qb, psib = bench.rings()
cq, cb = T.Context(bench.N_RING, qs, psis), T.Context(bench.N_RING, qb, psib)
rng = np.random.default_rng(0)
for B in (64, 128, 256):
    c1 = torch.from_numpy(bench.rand_ct(rng, qs, (B, 2))).pin_memory()
    c2 = torch.from_numpy(bench.rand_ct(rng, qs, (B, 2))).pin_memory()
    out = torch.empty((B, 3, bench.L_Q, bench.N_RING), dtype=torch.int64).pin_memory()
    for _ in range(3):
        cq.bfv_mul_host(cb, bench.T_PLAIN, c1, c2, out)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    it = 8
    for _ in range(it):
        cq.bfv_mul_host(cb, bench.T_PLAIN, c1, c2, out)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / it
    print(f"chunk {os.environ.get('TFB_HOST_CHUNK_MIB', '32')} MiB  batch {B:4d}: {dt * 1e3:7.3f} ms  {B / dt:8.0f} ct-mul/s   H2D {B * 4 * 2**20 / dt / 1e9:5.1f} GB/s  D2H {B * 3 * 2**20 / dt / 1e9:5.1f} GB/s", flush=True)
