"""Does write-combined pinned memory (cudaHostAllocWriteCombined) change the host->device rate of the pipelined
BFV multiply?  Inputs are only ever written by the CPU and read by the GPU, the case WC memory is meant for."""
import ctypes as C, glob, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import toyfhe_b200 as T
import bench

libs = glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*")) + glob.glob("/usr/local/cuda/lib64/libcudart.so*")
rt = C.CDLL(libs[0])
rt.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]

def host_alloc(shape, flags):
    n = int(np.prod(shape)) * 8
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), n, flags) == 0
    return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint64)), shape=(int(np.prod(shape)),)).reshape(shape)

torch.cuda.init(); torch.zeros(1).cuda()
qs, psis, qb, psib = bench.rings()
cq, cb = T.Context(bench.N_RING, qs, psis), T.Context(bench.N_RING, qb, psib)
rng = np.random.default_rng(0)
B = 128
src1, src2 = bench.rand_ct(rng, qs, (B, 2)), bench.rand_ct(rng, qs, (B, 2))
for name, flags in (("default pinned", 0), ("write-combined inputs", 4), ("portable", 1)):
    c1, c2 = host_alloc(src1.shape, flags), host_alloc(src2.shape, flags)
    c1[...] = src1; c2[...] = src2
    out = host_alloc((B, 3, bench.L_Q, bench.N_RING), 0)
    for _ in range(3):
        cq.bfv_mul_host(cb, bench.T_PLAIN, c1, c2, out)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    it = 10
    for _ in range(it):
        cq.bfv_mul_host(cb, bench.T_PLAIN, c1, c2, out)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / it
    print(f"{name:24s} batch {B}: {dt * 1e3:7.3f} ms  {B / dt:8.0f} ct-mul/s   H2D {B * 4 * 2**20 / dt / 1e9:5.1f} GB/s  D2H {B * 3 * 2**20 / dt / 1e9:5.1f} GB/s", flush=True)
