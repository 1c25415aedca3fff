"""Small driver for compute-sanitizer (memcheck / racecheck): a few rows through every NTT kernel family and
one small BFV multiply, keyswitch and rescale; results checked against the oracle."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import toyfhe_b200 as T
from oracle import c_oracle as CO

def rnd(rng, qs, shape, N):
    out = np.empty(shape + (len(qs), N), dtype=np.uint64)
    for i, q in enumerate(qs):
        out[..., i, :] = rng.integers(0, q, size=shape + (N,), dtype=np.uint64)
    return out

rng = np.random.default_rng(0)
H = T.Context.to_host
for logN, logqs, B in ((14, [60, 60], 100), (13, [60, 40], 200), (12, [50], 700), (15, [60, 40], 90), (10, [60], 3), (5, [40], 3)):
    N = 1 << logN
    qs, psis = T.prime_chain(N, sorted(logqs))
    ctx, orc = T.Context(N, qs, psis), CO.Rns(N, qs, psis)
    a = rnd(rng, qs, (B,), N)
    d = ctx.to_device(a)
    f = ctx.ntt_fwd(d)
    back = ctx.ntt_inv(f)
    assert np.array_equal(H(f[:2]), orc.nntt(a[:2])) and np.array_equal(H(back), a), logN
    print("ntt ok", logN, logqs, flush=True)
N, L, Lb, t = 1024, 8, 17, 65537
allq, allpsi = T.prime_chain(N, [60] * (L + Lb))
cq, cb = T.Context(N, allq[:L], allpsi[:L]), T.Context(N, allq[L:], allpsi[L:])
c1, c2 = rnd(rng, allq[:L], (3, 2), N), rnd(rng, allq[:L], (3, 2), N)
got = H(cq.bfv_mul(cb, t, cq.to_device(c1), cq.to_device(c2)))
assert np.array_equal(got, CO.bfv_mul(CO.Rns(N, allq[:L], allpsi[:L]), CO.Rns(N, allq[L:], allpsi[L:]), t, c1, c2))
print("bfv_mul ok", flush=True)
D = T.ndigits(allq[:L], 2)
key = cq.ntt_fwd(cq.to_device(rnd(rng, allq[:L], (D, 2), N)))
ks = cq.keyswitch(key, cq.to_device(rnd(rng, allq[:L], (2, 3), N)), 2)
rs = cq.rescale(cq.to_device(c1))
from toyfhe_b200 import sharding as S
lo, hi = S.shard_range(L, 1, 3)
shard = T.Context(N, allq[lo:hi], allpsi[lo:hi])
ct3 = cq.to_device(rnd(rng, allq[:L], (2, 3), N))
part = cq.keyswitch_shard(shard, lo, S.key_rows_for_shard(key, lo, hi), ct3, 2)
assert torch.equal(part, cq.keyswitch(key, ct3, 2)[:, :, lo:hi, :])
import math
Q = math.prod(allq[:L])
m = rng.integers(0, t, size=(2, N), dtype=np.uint64)
dec = cq.bfv_decode(t, Q // t, cq.bfv_encode(t, Q // t, cq.to_device(m)))
assert np.array_equal(H(dec), m)
u = cq.sample_uniform(1, 2, (2,))
g = cq.sample_gaussian(3.2, 1, 3, (2,))
torch.cuda.synchronize()
print("keyswitch / keyswitch_shard / rescale / encode / decode / samplers ran", flush=True)
# ---- round-2 kernels: gathered forward transform (ct_tensor, BFV joint basis with compact expansions), last global level on
# load (N = 2^15 out of place), CRT digits formed inside the transform, broadcast plaintext multiply / add
N = 4096
allq, allpsi = T.prime_chain(N, [60] * 10)
cq, cb = T.Context(N, allq[:3], allpsi[:3]), T.Context(N, allq[3:], allpsi[3:])
oq, ob = CO.Rns(N, allq[:3], allpsi[:3]), CO.Rns(N, allq[3:], allpsi[3:])
c1, c2 = rnd(rng, allq[:3], (5, 2), N), rnd(rng, allq[:3], (5, 2), N)
assert np.array_equal(H(cq.ct_tensor(cq.to_device(c1), cq.to_device(c2))), oq.ct_tensor(c1, c2))
assert np.array_equal(H(cq.bfv_mul(cb, 65537, cq.to_device(c1), cq.to_device(c2))), CO.bfv_mul(oq, ob, 65537, c1, c2))
print("gathered forward transforms ok", flush=True)
N = 1 << 15
qs, psis = T.prime_chain(N, [40, 60])
ctx, orc = T.Context(N, qs, psis), CO.Rns(N, qs, psis)
a = rnd(rng, qs, (40,), N)
f = ctx.ntt_fwd(ctx.to_device(a))
assert np.array_equal(H(f[:2]), orc.nntt(a[:2])) and np.array_equal(H(ctx.ntt_inv(f)), a)
print("2^15 forward with the last global level on load ok", flush=True)
N = 1 << 13
kq, kp = T.prime_chain(N, [60, 40, 40, 60])
ctx, ext = T.Context(N, kq[:-1], kp[:-1]), T.Context(N, kq, kp)
key = ext.ntt_fwd(ext.to_device(rnd(rng, kq, (4, 2), N)))
ct = ctx.to_device(rnd(rng, kq[:-1], (30, 2), N))
fused = ctx.keyswitch(key, ct, 0, ext=ext)
T.force_generic(1)
plain = ctx.keyswitch(key, ct, 0, ext=ext)
T.force_generic(0)
assert torch.equal(fused, plain)
p = ctx.to_device(rnd(rng, kq[:-1], (), N))
acc = ctx.mul_plain(ct, p)
ctx.mul_plain(ct, p, out=acc, accumulate=True)
ctx.add_plain_first(acc, p)
torch.cuda.synchronize()
print("CRT digits inside the transform / mul_plain / add_plain ok", flush=True)
# ---- base-2^w digits cut out of the binary limbs inside the transform (incl. digits that straddle two limbs), the split
# accumulate for few coefficients, lincomb, and the epilogue that pushes result rows into every rank's buffer (two ranks of
# this process on this GPU, one stream each)
N = 4096
qs, psis = T.prime_chain(N, [60, 60, 60, 60])
ctx, orc = T.Context(N, qs, psis), CO.Rns(N, qs, psis)
for w in (2, 7):
    D = T.ndigits(qs, w)
    keyh = rnd(rng, qs, (D, 2), N)
    key = ctx.ntt_fwd(ctx.to_device(keyh))
    cth = rnd(rng, qs, (1, 3), N)
    got = H(ctx.keyswitch(key, ctx.to_device(cth), w))
    w1, w2 = orc.keyswitch_accum(orc.keyswitch_digits(cth[0, 2], w), keyh, cth[0, 0], cth[0, 1])
    assert np.array_equal(got[0, 0], w1) and np.array_equal(got[0, 1], w2)
print("digits from limbs inside the transform + split accumulate ok", flush=True)
ranks = []
for r in range(2):
    c = T.Context(N, qs, psis)
    lo, hi = S.shard_range(4, r, 2)
    ranks.append(dict(ctx=c, lo=lo, shard=T.Context(N, qs[lo:hi], psis[lo:hi]), krows=S.key_rows_for_shard(key, lo, hi),
                      x=T.PeerExchange(c, r, 2, 2 * 4 * N), stream=torch.cuda.Stream()))
for rk in ranks:
    rk["x"].attach_local([o["x"] for o in ranks])
ctd = ctx.to_device(cth)
for rk in ranks:
    rk["ctx"].keyswitch_shard(rk["shard"], rk["lo"], rk["krows"], ctd, 7)       # sizes the scratch before the concurrent calls
torch.cuda.synchronize()
outs = []
for rk in ranks:
    with torch.cuda.stream(rk["stream"]):
        outs.append(rk["ctx"].keyswitch_shard_push(rk["shard"], rk["lo"], rk["krows"], ctd, 7, rk["x"]))
torch.cuda.synchronize()
for rk, o in zip(ranks, outs):
    assert not rk["x"].timed_out() and np.array_equal(H(o), got)
for rk in ranks:
    rk["x"].close()
print("peer-push epilogue ok", flush=True)
