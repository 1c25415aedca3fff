// Host-side table construction and number theory for a NegacyclicRing context.
// Replaces, for the engine, what the reference recomputes on every transform
// (psi^i by power and the CTPlan twiddles, pow2_cyc_rings.jl:298-301,315-317) and
// the GaloisFields / Primes calls of the ring constructors
// (pow2_cyc_rings.jl:40, crt.jl:282-295).
#pragma once
#include <vector>

#include "modarith.cuh"

static inline u32 h_brev(u32 x, int bits) {
    u32 r = 0;
    for (int i = 0; i < bits; i++)
        if (x >> i & 1) r |= 1u << (bits - 1 - i);
    return r;
}

struct HostTables {
    std::vector<tw_t> fwd;  // fwd[k] = psi^brev(k)   (merged CT twiddles, index m+i)
    std::vector<tw_t> inv;  // inv[k] = psi^-brev(k)  (merged GS twiddles)
    tw_t ninv;              // N^-1
    tw_t ninv_w1;           // N^-1 * inv[1]
    PrimeConst pc;
};

static inline void build_tables(u64 N, u64 q, u64 psi, HostTables& ht) {
    int lg = 0;
    while ((1ull << lg) < N) lg++;
    std::vector<u64> pw(N), ipw(N);
    u64 ipsi = h_invmod(psi, q);
    u64 t = 1, ti = 1;
    for (u64 i = 0; i < N; i++) {
        pw[i] = t;
        ipw[i] = ti;
        t = h_mulmod(t, psi, q);
        ti = h_mulmod(ti, ipsi, q);
    }
    ht.fwd.resize(N);
    ht.inv.resize(N);
    for (u64 k = 0; k < N; k++) {
        u32 r = h_brev((u32)k, lg);
        ht.fwd[k] = h_tw(pw[r], q);
        ht.inv[k] = h_tw(ipw[r], q);
    }
    u64 ninv = h_invmod(N % q, q);
    ht.ninv = h_tw(ninv, q);
    ht.ninv_w1 = h_tw(N > 1 ? h_mulmod(ninv, ht.inv[1].w, q) : ninv, q);
    ht.pc = h_prime_const(q);
}

// Thread-order copy of the pass-3 twiddles of the row-resident kernels (ntt_core.cuh
// pass3_base): rows of N = 2^logN words are processed as 2^s0 sub-blocks of 2^(10+R)
// positions, R = min(logN,14)-10; thread t = (warp w, lane l) of sub-block blk handles,
// in group g, the positions a = brev5(l), b = brev5(w*G+g) and uses at level u (1..R)
// block j the natural-table entry 2^(s0+9+u) + blk*2^(9+u) + ((a*32+b) << (u-1)) + j.
static inline void permute_pass3(const tw_t* src, tw_t* dst, int logN) {
    const u64 N = 1ull << logN;
    for (u64 i = 0; i < N; i++) dst[i] = src[0];
    if (logN <= 10) return;
    const int R = (logN > 14 ? 14 : logN) - 10, s0 = logN - 10 - R;
    const u32 Nb = 1u << (10 + R), T = Nb / 32, RS = 1u << R, G = 32 / RS;
    for (u32 blk = 0; blk < (1u << s0); blk++)
        for (u32 g = 0; g < G; g++)
            for (int u = 1; u <= R; u++)
                for (u32 j = 0; j < (1u << (u - 1)); j++)
                    for (u32 t = 0; t < T; t++) {
                        const u32 a = h_brev(t & 31, 5), b = h_brev((t >> 5) * G + g, 5);
                        const u64 from = (1ull << (s0 + 9 + u)) + ((u64)blk << (9 + u)) + ((u64)(a * 32 + b) << (u - 1)) + j;
                        const u64 to = (u64)blk * Nb + (u64)(g * (RS - 1) + (1u << (u - 1)) - 1 + j) * T + t;
                        dst[to] = src[from];
                    }
}

// deterministic Miller-Rabin for 64-bit integers
static inline bool h_is_prime(u64 n) {
    if (n < 2) return false;
    static const u64 bases[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    for (u64 p : bases) {
        if (n % p == 0) return n == p;
    }
    u64 d = n - 1;
    int s = 0;
    while ((d & 1) == 0) {
        d >>= 1;
        s++;
    }
    for (u64 a : bases) {
        u64 x = h_powmod(a, d, n);
        if (x == 1 || x == n - 1) continue;
        bool comp = true;
        for (int i = 0; i < s - 1; i++) {
            x = h_mulmod(x, x, n);
            if (x == n - 1) {
                comp = false;
                break;
            }
        }
        if (comp) return false;
    }
    return true;
}

// smallest integer of multiplicative order exactly n (n a power of two) mod prime q
// (GaloisFields.minimal_primitive_root as used at pow2_cyc_rings.jl:40, crt.jl:142-144)
static inline bool h_minimal_primitive_root(u64 q, u64 n, u64* out) {
    if (n < 2 || (n & (n - 1)) || (q - 1) % n != 0) return false;
    u64 e = (q - 1) / n, r = 0;
    for (u64 a = 2; a < q; a++) {
        r = h_powmod(a, e, q);
        if (h_powmod(r, n / 2, q) == q - 1) break;
        r = 0;
        if (a > 1000) return false;
    }
    if (!r) return false;
    u64 best = r, r2 = h_mulmod(r, r, q), x = r;
    for (u64 i = 1; i < n / 2; i++) {
        x = h_mulmod(x, r2, q);
        if (x < best) best = x;
    }
    *out = best;
    return true;
}
