import sys, random
sys.path.insert(0, '/root/repo')
from oracle import toyfhe_oracle as O

def brev(x, bits):
    r = 0
    for b in range(bits):
        if x >> b & 1: r |= 1 << (bits-1-b)
    return r

def fwd_merged(c, q, psi):
    N = len(c); lg = N.bit_length()-1
    tw = [pow(psi, brev(k, lg), q) for k in range(N)]
    a = list(c); m = 1; t = N//2
    while m < N:
        for i in range(m):
            S = tw[m+i]
            for j in range(i*2*t, i*2*t+t):
                U = a[j]; V = a[j+t]*S % q
                a[j] = (U+V) % q; a[j+t] = (U-V) % q
        m *= 2; t //= 2
    return a  # position p holds chat[brev(p)]

def inv_merged(a, q, psi):
    N = len(a); lg = N.bit_length()-1
    ipsi = pow(psi, q-2, q)
    tw = [pow(ipsi, brev(k, lg), q) for k in range(N)]
    a = list(a); m = N//2; t = 1
    while m >= 1:
        for i in range(m):
            S = tw[m+i]
            for j in range(i*2*t, i*2*t+t):
                U = a[j]; V = a[j+t]
                a[j] = (U+V) % q; a[j+t] = (U-V)*S % q
        m //= 2; t *= 2
    ninv = pow(N, q-2, q)
    return [x*ninv % q for x in a]

for N in (4, 16, 64):
    qs, psis = O.prime_chain(N, (40,))
    q, psi = qs[0], psis[0]
    c = [random.randrange(q) for _ in range(N)]
    lg = N.bit_length()-1
    ref = O.nntt(c, q, psi)
    got = fwd_merged(c, q, psi)
    assert all(got[p] == ref[brev(p, lg)] for p in range(N)), N
    back = inv_merged(got, q, psi)
    assert back == c
print("merged CT/GS formulation OK")
