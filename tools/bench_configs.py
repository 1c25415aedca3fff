"""Times the engine on the BASELINE.json configurations other than the headline bench.py line
(synthetic uniform residues, inputs resident in HBM, CUDA events, >= 3 warm-up passes):

  C2  N=2^14, L=8      forward / inverse NTT, ring product, ciphertext tensor (CKKS/BGV form)
  C3  N=2^15, 11 primes (60, 9x40, special 60)  rotate+keyswitch (CRT digits, ModulusRaised), plaintext
      multiply, rescale, and the 128x128 diagonal matmul loop of test/ckks_matmul.jl scaled up
  C4  N=2^14, L=8      keyswitch / relinearise with relin_window w=2 (D = 241 digit polynomials)
  C5  N=2^13, 7 primes (60, 5x40, special 60)   op mix of examples/encrypted_mnist/infer.jl per pipeline

Writes one JSON object to stdout (and to --out).  Each entry: ms per call, units/s, algorithmic bytes
per call and achieved GB/s against the measured HBM peak."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import toyfhe_b200 as T  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def rand_res(rng, qs, N, shape):
    out = np.empty(shape + (len(qs), N), dtype=np.uint64)
    for i, q in enumerate(qs):
        out[..., i, :] = rng.integers(0, q, size=shape + (N,), dtype=np.uint64)
    return out


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def entry(ms, units, unit_name, alg_bytes, pk):
    gbs = alg_bytes / (ms * 1e-3) / 1e9 if alg_bytes else None
    return {"ms": round(ms, 4), "per_s": units / (ms * 1e-3), "unit": unit_name, "algorithmic_bytes": alg_bytes,
            "achieved_gbs": None if gbs is None else round(gbs, 1), "frac_of_measured_hbm": None if gbs is None else round(gbs / pk, 4)}


def c2(res, pk, B):
    N, L = 2 ** 14, 8
    qs, psis = T.prime_chain(N, [60] * L)
    ctx = T.Context(N, qs, psis)
    rng = np.random.default_rng(0)
    blk = min(B, 16)
    h = rand_res(rng, qs, N, (blk, 2))
    a = ctx.to_device(h).repeat((B + blk - 1) // blk, 1, 1, 1)[:B].contiguous()
    b = torch.roll(a, 1, dims=3)
    out = torch.empty_like(a)
    out3 = ctx.empty((B, 3, L, N))
    rows = 2 * B * L
    rb = N * 8
    res["C2 fwd NTT N=2^14 L=8"] = entry(timeit(lambda: ctx.ntt_fwd(a, out=out)), 2 * B, "RNS-NTT/s", 2 * rows * rb, pk)
    res["C2 inv NTT N=2^14 L=8"] = entry(timeit(lambda: ctx.ntt_inv(a, out=out)), 2 * B, "RNS-NTT/s", 2 * rows * rb, pk)
    res["C2 ring product (primal in/out)"] = entry(timeit(lambda: ctx.ring_mul(a, b, out=out)), 2 * B, "ring-products/s", 3 * rows * rb, pk)
    res["C2 ciphertext tensor (CKKS/BGV form)"] = entry(timeit(lambda: ctx.ct_tensor(a, b, out=out3)), B, "ciphertext-muls/s", 7 * B * L * rb, pk)


def ckks_chain(N, logs_main, special=60):
    """q0 + scale primes + special prime, built like examples/encrypted_mnist/infer.jl:97-105 (one ascending
    chain; the special prime is the last 60-bit one)."""
    n60 = sum(1 for x in logs_main if x == 60) + 1
    n40 = sum(1 for x in logs_main if x == 40)
    q40, p40 = T.prime_chain(N, [40] * n40)
    q60, p60 = T.prime_chain(N, [60] * n60)
    qs = [q60[0]] + q40
    ps = [p60[0]] + p40
    return qs, ps, q60[-1], p60[-1]


def keyswitch_setup(N, qs, psis, sp_q, sp_psi, rng):
    """CRT-digit evaluation key over the raised ring (modulusraising.jl): synthetic uniform residues in the
    NTT domain (what a key looks like to the kernels)."""
    ctx = T.Context(N, qs, psis)
    ext = T.Context(N, qs + [sp_q], psis + [sp_psi])
    D = len(qs)
    key = ext.to_device(rand_res(rng, qs + [sp_q], N, (D, 2)))
    return ctx, ext, key


def c3(res, pk):
    N = 2 ** 15
    qs, psis, sp, spsi = ckks_chain(N, [60] + [40] * 9)
    rng = np.random.default_rng(1)
    ctx, ext, key = keyswitch_setup(N, qs, psis, sp, spsi, rng)
    L = len(qs)
    rb = N * 8
    B = 8
    ct = ctx.to_device(rand_res(rng, qs, N, (B, 2)))
    pt = ctx.to_device(rand_res(rng, qs, N, (B, 1)))
    tmp = torch.empty_like(ct)
    out = torch.empty_like(ct)
    g = T.galois_element_from_steps(1, N)

    def rot():
        ctx.galois(ct, g, out=tmp)
        ctx.keyswitch(key, tmp, 0, ext=ext, out=out)

    t_rot = timeit(rot, iters=5)
    ks_bytes = B * (2 * L + 2 * L) * rb + L * 2 * (L + 1) * rb
    res["C3 rotate+keyswitch N=2^15 (CRT digits, special prime)"] = entry(t_rot, B, "rotations/s", ks_bytes, pk)
    dual = ctx.ntt_fwd(ct)
    ptd = ctx.ntt_fwd(pt)
    t_pm = timeit(lambda: ctx.mul(dual, ptd.expand(B, 2, L, N).contiguous(), out=tmp), iters=5)
    res["C3 plaintext multiply (dual domain)"] = entry(t_pm, B, "ct*pt/s", B * 3 * 2 * L * rb, pk)
    outr = ctx.empty((B, 2, L - 1, N))
    t_rs = timeit(lambda: ctx.rescale(ct, out=outr), iters=5)
    res["C3 rescale N=2^15 L=10->9"] = entry(t_rs, B, "rescales/s", B * 2 * (2 * L - 1) * rb, pk)
    t_f = timeit(lambda: ctx.ntt_fwd(ct, out=tmp), iters=5)
    res["C3 fwd NTT N=2^15 L=10"] = entry(t_f, 2 * B, "RNS-NTT/s", 2 * 2 * B * L * rb, pk)
    # 128x128 diagonal matmul (ckks_matmul.jl:34-42 scaled): 127 rotations + 128 plaintext mults + 127 adds + 1 rescale
    diags = 128

    def matmul():
        acc = None
        cur = ct
        for d in range(diags):
            if d:
                ctx.galois(ct, g, out=tmp)
                cur = ctx.keyswitch(key, tmp, 0, ext=ext, out=out)
            prod = ctx.ring_mul(cur, pt.expand(B, 2, L, N).contiguous())
            acc = prod if acc is None else ctx.add(acc, prod, out=acc)
        return ctx.rescale(acc, out=outr)

    t_mm = timeit(matmul, iters=2, warm=1)
    res["C3 CKKS 128x128 diagonal matmul + rescale (batch of 8 ciphertexts)"] = entry(t_mm, B, "matmuls/s", None, pk)


def c4(res, pk):
    N, L, w = 2 ** 14, 8, 2
    qs, psis = T.prime_chain(N, [60] * L)
    ctx = T.Context(N, qs, psis)
    D = T.ndigits(qs, w)
    rng = np.random.default_rng(2)
    rb = N * 8
    blk = rand_res(rng, qs, N, (8, 2))
    key = ctx.to_device(blk).repeat((D + 7) // 8, 1, 1, 1)[:D].contiguous()     # [D][2][L][N] synthetic key
    B = 4
    ct = ctx.to_device(rand_res(rng, qs, N, (B, 3)))
    out = ctx.empty((B, 2, L, N))
    t = timeit(lambda: ctx.keyswitch(key, ct, w, out=out), iters=3, warm=2)
    alg = B * (3 + 2) * L * rb + D * 2 * L * rb
    res[f"C4 relinearise w=2 (D={D} digit polys) N=2^14 L=8"] = entry(t, B, "keyswitches/s", alg, pk)
    res[f"C4 relinearise w=2 (D={D} digit polys) N=2^14 L=8"]["digit_row_ntts_per_call"] = B * D * L


def c5(res, pk):
    N = 2 ** 13
    qs, psis, sp, spsi = ckks_chain(N, [60] + [40] * 5)
    rng = np.random.default_rng(3)
    ctx, ext, key = keyswitch_setup(N, qs, psis, sp, spsi, rng)
    L = len(qs)
    B = 64   # independent pipelines processed together
    ct = ctx.to_device(rand_res(rng, qs, N, (B, 2)))
    ct3 = ctx.to_device(rand_res(rng, qs, N, (B, 3)))
    pt = ctx.to_device(rand_res(rng, qs, N, (B, 1))).expand(B, 2, L, N).contiguous()
    tmp, out = torch.empty_like(ct), torch.empty_like(ct)
    out3 = ctx.empty((B, 3, L, N))
    outr = ctx.empty((B, 2, L - 1, N))
    g = T.galois_element_from_steps(1, N)
    ops = {
        "encrypt (2 ring products + adds)": (49, lambda: (ctx.ring_mul(ct, pt, out=tmp), ctx.add(tmp, ct, out=out))),
        "ct*scalar": (196, lambda: ctx.scalar_mul(ct, 12345, out=out)),
        "ct*ct + relinearise": (5, lambda: (ctx.ct_tensor(ct, ct, out=out3), ctx.keyswitch(key, out3, 0, ext=ext, out=out))),
        "rescale": (10, lambda: ctx.rescale(ct, out=outr)),
        "rotate + keyswitch": (315, lambda: (ctx.galois(ct, g, out=tmp), ctx.keyswitch(key, tmp, 0, ext=ext, out=out))),
        "plaintext-vector multiply": (320, lambda: ctx.ring_mul(ct, pt, out=out)),
    }
    total = 0.0
    detail = {}
    for name, (count, fn) in ops.items():
        ms = timeit(fn, iters=5)
        detail[name] = {"count_per_pipeline": count, "ms_per_batched_call": round(ms, 4)}
        total += count * ms
    res["C5 encrypted-MNIST op mix N=2^13, 7 primes (per GPU)"] = {
        "ms_per_batch_of_pipelines": round(total, 2), "pipelines_per_batch": B, "per_s": B / (total * 1e-3), "unit": "pipelines/s",
        "ops": detail, "note": "all ops at the top level (7 primes); the real pipeline drops primes as it rescales"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--only", default="")
    ap.add_argument("--batch", type=int, default=128)
    args = ap.parse_args()
    pk = peak()
    res = {"hbm_peak_gbs_measured": pk, "gpu": torch.cuda.get_device_name(0)}
    todo = args.only.split(",") if args.only else ["c2", "c3", "c4", "c5"]
    if "c2" in todo:
        c2(res, pk, args.batch)
    if "c3" in todo:
        c3(res, pk)
    if "c4" in todo:
        c4(res, pk)
    if "c5" in todo:
        c5(res, pk)
    txt = json.dumps(res, indent=1)
    print(txt)
    if args.out:
        with open(args.out, "w") as f:
            f.write(txt + "\n")


if __name__ == "__main__":
    main()
