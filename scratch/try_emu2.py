import sys, ctypes as C, numpy as np
sys.path.insert(0,'/root/repo')
from oracle import toyfhe_oracle as O
from oracle import c_oracle as CO
emu = C.CDLL('/root/repo/tests/emu/libemu.so')
u64p = C.POINTER(C.c_uint64)
def P(a): return a.ctypes.data_as(u64p)
rng = np.random.default_rng(0)
N = 1 << 14
qs, psis = O.prime_chain(N, (60,))
q, psi = qs[0], psis[0]
r = CO.Rns(N, qs, psis)
a = rng.integers(0, q, size=(1,N), dtype=np.uint64)
want = r.nntt(a)
got = np.zeros_like(a)
emu.emu_ntt2(0, C.c_uint64(q), C.c_uint64(psi), C.c_uint32(0), P(a), P(got))
print("fwd ok:", np.array_equal(want, got))
back = np.zeros_like(a)
emu.emu_ntt2(1, C.c_uint64(q), C.c_uint64(psi), C.c_uint32(0), P(want), P(back))
print("inv ok:", np.array_equal(back, a))
print("bank conflict degree:", emu.emu_bank_conflicts2())
