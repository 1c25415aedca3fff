"""Forward / inverse NTT rate at every ring degree and prime-chain shape of the BASELINE configs
(C2: N=2^14 8x60; C3: N=2^15 60+9x40+60; C5: N=2^13 60+5x40+60; bfv_crt-like N=2^12 3x50)."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import toyfhe_b200 as T

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

out = {}
for name, logN, logqs, polys in (("C2 N=2^14 8x60", 14, [60] * 8, 256), ("C3 N=2^15 60+9x40+60", 15, [60] + [40] * 9 + [60], 96),
                                 ("C5 N=2^13 60+5x40+60", 13, [60] + [40] * 5 + [60], 512), ("N=2^12 3x50", 12, [50] * 3, 2048),
                                 ("N=2^16 4x60", 16, [60] * 4, 64)):
    N = 1 << logN
    qs, psis = T.prime_chain(N, sorted(logqs))
    ctx = T.Context(N, qs, psis)
    rng = np.random.default_rng(0)
    a = np.empty((polys, len(qs), N), dtype=np.uint64)
    for i, q in enumerate(qs):
        a[:, i, :] = rng.integers(0, q, size=(polys, N), dtype=np.uint64)
    d = ctx.to_device(a)
    o = torch.empty_like(d)
    rows = polys * len(qs)
    nbytes = rows * N * 16
    res = {}
    for label, ver, mode in (("gen1 lazy", 1, 1), ("gen3", 3, 2)):
        T.ntt_version(ver); T.ntt_max_mode(mode)
        f = timeit(lambda: ctx.ntt_fwd(d, out=o))
        i = timeit(lambda: ctx.ntt_inv(d, out=o))
        res[label] = {"fwd_ms": round(f, 4), "fwd_gbs": round(nbytes / f / 1e6, 1), "inv_ms": round(i, 4), "inv_gbs": round(nbytes / i / 1e6, 1),
                      "fwd_rns_ntt_per_s": round(polys / f * 1e3, 1)}
        print(f"{name:28s} {label:10s} fwd {f:.3f} ms {nbytes / f / 1e6:7.0f} GB/s | inv {i:.3f} ms {nbytes / i / 1e6:7.0f} GB/s  ({rows} rows)", flush=True)
    T.ntt_version(3); T.ntt_max_mode(2)
    out[name] = res
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
