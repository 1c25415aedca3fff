// Third-generation ladder for row-resident sub-blocks of N = 2^(10+R) positions, R = 2..4, on primes
// q = 2^b + e with 32 <= b <= 60 and e < 2^28 (the reference's nextprime(2^logq + 1) chains, crt.jl:282-295;
// b may differ between the primes of one ring, e.g. the 60/40-bit CKKS chains of examples/encrypted_mnist).
// Same transform as ntt_core.cuh (pow2_cyc_rings.jl:295-303: c^[k] = sum_j c[j] psi^(j(2k+1)), natural order in
// and out) and the same N/32 threads x 32 residues, levels 5+5+R, but built around what bounds the kernel on
// sm_100a: the FMA-heavy pipe (IMAD 2, IMAD.WIDE/IMAD.HI 4 cycles per warp instruction) and the equally
// half-rate ALU pipe (tools/bfly_bench4.cu, tools/pipe_probe5.cu):
//   * approximate-quotient Shoup product shoup_lazy4 (T in [0,4q)): 1 IMAD.WIDE + 2 IMAD.HI instead of 4 IMAD.WIDE
//     for the quotient, and q's shape saves one more IMAD.WIDE in the tail;
//   * values grow by 4q per level; X is brought back to (0,2q) at levels 4 of pass 1, 1 and 4 of passes 2 and 3
//     by  x + (q - floor(x/2^b) q)  with the constant read from a 16-entry shared-memory table and the addition
//     FUSED into the butterfly's own 3-input adds (X' = x + c2[k] + T, Y' = x + c3[k] - T, c3 = c2 + 4q): a reduction
//     costs one shift, one address and one 128-bit shared load, nothing on the FMA-heavy pipe;
//   * the row sits in shared memory SKEWED (slot(a,i) = T a + skew(a) + i, skew(a) = 2(a>>2) [+ RS a if RS < 16])
//     instead of XOR-swizzled: every access of the three passes is then base register + immediate (no per-access
//     address arithmetic), the 64-bit accesses of passes 1-2 and the 128-bit loads of pass 3 stay
//     bank-conflict-free for every R, and the TMA bulk copies (one per row of T positions) keep 16-byte alignment.
// __host__ __device__ like ntt_core.cuh so tests/emu runs the same index logic on the CPU.
#pragma once
#include "ntt_core.cuh"

namespace v3 {
template <int R>
struct Lay {
    typedef NttGeo<R> Geo;
    // pass 2 reads runs of RS consecutive words, 16/RS runs (consecutive a) per half-warp: for RS < 16 row a is
    // shifted by RS a words so the runs fall on distinct bank groups; 2(a>>2) separates the rows a quarter-warp
    // of pass 3 touches (a = 4 brev3(j) + const).  The skew is non-decreasing in a: rows never overlap.
    static constexpr u32 STEP = Geo::RS < 16 ? Geo::RS : 0;
    static constexpr u32 ROW_WORDS = Geo::N + 32 * STEP + 16;      // skewed row buffer
    static constexpr u32 ROW_BYTES = ROW_WORDS * 8;
    // levels of pass 3 whose X operands are reduced (bit u-1): bound 10 -> (reduce) 6 -> 10 -> 14 [-> (reduce) 6]
    static constexpr u32 P3MASK = R >= 4 ? 0x09 : 0x01;
};
template <int R>
TFB_HD u32 skew(const u32 a) { return 2 * (a >> 2) + Lay<R>::STEP * a; }
template <int R>
TFB_HD u32 slot(const u32 a, const u32 idx) { return a * NttGeo<R>::T + skew<R>(a) + idx; }

struct __align__(16) redent_t {
    u64 c2, c3;   // c2 = q - k q, c3 = c2 + 4q  (mod 2^64), k = 0..15
};
struct Red3 {
    u64 q, q4;
    u32 ne, shb;          // 2^32 - e, b - 32
    const redent_t* tab;  // this prime's 16 entries (forward ladder: c2 and c3 = c2 + 4q in one 128-bit load)
    const u64* tab8;      // inverse ladder: c2 alone, 8-byte entries -- the 16 entries cover 128 bytes, so lanes with
                          // different k never collide on a bank (16-byte entries put k and k+8 on the same banks)
};
TFB_HD void fill_redtab(redent_t* tab, const u64 q) {
    for (u32 k = 0; k < 16; k++) {
        tab[k].c2 = q - (u64)k * q;
        tab[k].c3 = tab[k].c2 + 4 * q;
    }
}
// b = floor(log2 q) (PrimeParams::sh)
TFB_HD Red3 make_red3(const u64 q, const u32 b, const redent_t* tab) {
    Red3 r;
    r.q = q;
    r.q4 = 4 * q;
    r.ne = 0u - (u32)(q - (1ull << b));
    r.shb = b - 32;
    r.tab = tab;
    r.tab8 = nullptr;
    return r;
}
TFB_HD void fill_redtab8(u64* tab8, const u64 q) {
    for (u32 k = 0; k < 16; k++) tab8[k] = q - (u64)k * q;
}
// host-side eligibility of one prime (api.cu decides per context)
static inline bool prime_ok(const u64 q) {
    u32 b = 0;
    while (b < 63 && (q >> (b + 1))) b++;
    return b >= 32 && b <= 60 && q - (1ull << b) < (1ull << 28);
}
// floor(x / 2^b) for x < 16 * 2^b
TFB_HD u32 top4(const u64 x, const Red3& rp) { return (u32)(x >> 32) >> rp.shb; }

#ifndef __CUDA_ARCH__
static unsigned long long g_emu_overflow3 = 0;   // tests/emu: lazy-range violations (must stay 0)
#endif

// CT butterfly, X in [0,15q) if RED else X + 4q < 2^64;  Y any 64-bit value
template <bool RED>
TFB_HD void bfly3(u64& X, u64& Y, const tw_t w, const Red3& rp) {
    const u64 t = shoup_lazy4(Y, w.w, w.wp, rp.q, rp.ne, rp.shb);
    const u64 x = X;
    if (RED) {
#ifndef __CUDA_ARCH__
        if (top4(x, rp) > 15) { g_emu_overflow3++; return; }
#endif
        const redent_t c = rp.tab[top4(x, rp)];
#ifndef __CUDA_ARCH__
        const u64 xr = x + c.c2;
        if (t >= rp.q4 || xr >= 2 * rp.q || xr == 0) g_emu_overflow3++;
#endif
        X = x + c.c2 + t;
        Y = x + c.c3 - t;
    } else {
#ifndef __CUDA_ARCH__
        if (t >= rp.q4 || (((u128)x + rp.q4) >> 64) != 0) g_emu_overflow3++;
#endif
        X = x + t;
        Y = x - t + rp.q4;
    }
}
// LV levels over 2^LV registers; level u (1-based) reduces its X operands iff bit u-1 of REDMASK is set
template <int LV, u32 REDMASK>
TFB_HD void levels3(u64* x, const tw_t* __restrict__ tw, const u32* tb, const Red3& rp, const u32 js = 1) {
#pragma unroll
    for (int u = 1; u <= LV; u++) {
        const int half = (1 << LV) >> u;
#pragma unroll
        for (int j = 0; j < (1 << (u - 1)); j++) {
            const tw_t w = tw[tb[u - 1] + j * js];
#pragma unroll
            for (int k = 0; k < half; k++) {
                if ((REDMASK >> (u - 1)) & 1) bfly3<true>(x[j * 2 * half + k], x[j * 2 * half + k + half], w, rp);
                else bfly3<false>(x[j * 2 * half + k], x[j * 2 * half + k + half], w, rp);
            }
        }
    }
}
// any v < 16 * 2^b -> canonical
TFB_HD u64 canon3(const u64 v, const Red3& rp) { return csub(v + rp.tab[top4(v, rp)].c2, rp.q); }
TFB_HD u64 canon3i(const u64 v, const Red3& rp) { return csub(v + rp.tab8[top4(v, rp)], rp.q); }   // inverse kernels (8-byte table)

// pass 1 (levels 1..5): thread t holds a = 0..31 at index t; canonical input, bound 1 -> 13 -> (reduce) 6 -> 10
template <int R>
TFB_HD void pass1(u64* x, u64* smem, const tw_t* __restrict__ tw, const Red3& rp, const u32 t, const u32 s0, const u32 blk) {
#pragma unroll
    for (int a = 0; a < 32; a++) x[a] = smem[slot<R>(a, t)];
    u32 tb[5];
#pragma unroll
    for (int s = 1; s <= 5; s++) tb[s - 1] = (1u << (s0 + s - 1)) + (blk << (s - 1));
    levels3<5, 0x08>(x, tw, tb, rp);
#pragma unroll
    for (int a = 0; a < 32; a++) smem[slot<R>(a, t)] = x[a];
}
// pass 1 on a CRT keyswitch digit (rlwe_she.jl:326-329): the row in shared memory holds residues modulo q_k of the
// ciphertext's last component; the digit polynomial under this row's target prime is the CENTRED residue
// (signedmod.jl:12-19: c > q_k / 2 ? c - q_k : c) re-embedded modulo the target prime -- formed here, in registers, while
// the row is loaded, so the digit rows are never written to and re-read from HBM.  `br_hi` = floor(2^128 / p) high word of
// the target prime (one-word Barrett, needed only when q_k / 2 >= p: uniform per row).
template <int R>
TFB_HD void pass1_crt(u64* x, u64* smem, const tw_t* __restrict__ tw, const Red3& rp, const u32 t, const u64 qk, const u64 br_hi) {
    const u64 hq = qk >> 1, p = rp.q;
    if (hq < p) {
        // |centred residue| <= q_k / 2 < p: the embedding is c itself or c - q_k + p (never zero, always below p)
        const u64 d = p - qk;                          // modulo 2^64
#pragma unroll
        for (int a = 0; a < 32; a++) {
            const u64 c = smem[slot<R>(a, t)];
            x[a] = c > hq ? c + d : c;
        }
    } else {
        // a wide residue under a narrower prime (the 60-bit prime's digit under the 40-bit primes of a CKKS chain)
#pragma unroll 4
        for (int a = 0; a < 32; a++) {
            const u64 c = smem[slot<R>(a, t)];
            const bool neg = c > hq;
            const u64 v = neg ? qk - c : c;
            u64 r = v - mulhi64(v, br_hi) * p;          // [0, 3p)
            r = csub(r, 2 * p);
            r = csub(r, p);
            x[a] = (neg && r) ? p - r : r;
        }
    }
    u32 tb[5];
#pragma unroll
    for (int s = 1; s <= 5; s++) tb[s - 1] = 1u << (s - 1);
    levels3<5, 0x08>(x, tw, tb, rp);
#pragma unroll
    for (int a = 0; a < 32; a++) smem[slot<R>(a, t)] = x[a];
}
// Base-2^w keyswitch digits (rlwe_she.jl:331-337) formed while the row is loaded: the row in shared memory is ONE 64-bit limb
// of the binary integers X_n (limb-major [limbs][N], ks_limbs_kernel), the digit is bits off .. off+w-1 of it; a digit that
// straddles two limbs takes its upper bits from the next limb row with ordinary (L2) loads -- `hi` is null otherwise.
// 2^w <= every target prime (the caller checks), so the digit is its own canonical residue.
template <int R>
TFB_HD void pass1_pow2(u64* x, u64* smem, const tw_t* __restrict__ tw, const Red3& rp, const u32 t, const u32 off, const u64 mask,
                       const u64* __restrict__ hi) {
    if (hi == nullptr) {
#pragma unroll
        for (int a = 0; a < 32; a++) x[a] = (smem[slot<R>(a, t)] >> off) & mask;
    } else {
#pragma unroll
        for (int a = 0; a < 32; a++) x[a] = ((smem[slot<R>(a, t)] >> off) | (hi[a * NttGeo<R>::T + t] << (64 - off))) & mask;
    }
    u32 tb[5];
#pragma unroll
    for (int s = 1; s <= 5; s++) tb[s - 1] = 1u << (s - 1);
    levels3<5, 0x08>(x, tw, tb, rp);
#pragma unroll
    for (int a = 0; a < 32; a++) smem[slot<R>(a, t)] = x[a];
}
// pass 2 (levels 6..10): thread (a2 = t >> R, c2 = t mod RS) holds b = 0..31; bound 10 -> (reduce) 6 -> 10 -> 14 -> (reduce) 6 -> 10
template <int R>
TFB_HD void pass2(u64* x, u64* smem, const tw_t* __restrict__ tw, const Red3& rp, const u32 t, const u32 s0, const u32 blk) {
    typedef NttGeo<R> Geo;
    const u32 a2 = t >> R, c2 = t & (Geo::RS - 1);
    u64* base = smem + slot<R>(a2, c2);
#pragma unroll
    for (int b = 0; b < 32; b++) x[b] = base[b * Geo::RS];
    u32 tb[5];
#pragma unroll
    for (int u = 1; u <= 5; u++) tb[u - 1] = (1u << (s0 + 4 + u)) + (blk << (4 + u)) + (a2 << (u - 1));
    levels3<5, 0x09>(x, tw, tb, rp);
#pragma unroll
    for (int b = 0; b < 32; b++) base[b * Geo::RS] = x[b];
}
// pass 3 (levels 11..10+R): thread (warp w, lane l) holds, for a3 = brev5(l), the G groups b3 = brev5(G w + g), all c
template <int R>
TFB_HD void pass3_load(u64* x, const u64* smem, const u32 t) {
    typedef NttGeo<R> Geo;
    const u32 w = t >> 5, lane = t & 31;
    const u64* base = smem + slot<R>(brev_bits(lane, 5), brev_bits(Geo::G * w, 5) * Geo::RS);
#pragma unroll
    for (int g = 0; g < (int)Geo::G; g++) {
        const u32 off = brev_bits((u32)g, 5) * Geo::RS;   // brev5(G w + g) = brev5(G w) + brev5(g)
#ifdef __CUDA_ARCH__
#pragma unroll
        for (int c = 0; c < (int)Geo::RS; c += 2) {
            const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(base + off + c);
            x[g * Geo::RS + c] = v.x;
            x[g * Geo::RS + c + 1] = v.y;
        }
#else
        for (int c = 0; c < (int)Geo::RS; c++) x[g * Geo::RS + c] = base[off + c];
#endif
    }
}
// canonical, natural-order coalesced stores
template <int R, bool S0ZERO>
TFB_HD void pass3_compute_store(u64* x, u64* __restrict__ orow, const tw_t* __restrict__ twc, const Red3& rp,
                                const u32 t, const u32 s0, const u32 blk) {
    typedef NttGeo<R> Geo;
    const u32 w = t >> 5, lane = t & 31;
    const u32 oblk = S0ZERO ? 0 : brev_bits(blk, (int)s0);
#ifdef __CUDA_ARCH__
    // The G groups run the SAME code on different registers.  With 8 groups (N = 2^12) the loop is kept rolled -- the
    // group's residues are rotated into x[0..RS) -- which shrinks the instruction footprint (+2.5 % there); with 2 or 4
    // groups the unrolled form is faster (measured: -2 % / -3.5 % rolled at N = 2^14 / 2^13, profiles/r01_ntt_sizes.txt).
    constexpr bool ROLL = Geo::G >= 8;
#else
    constexpr bool ROLL = false;
#endif
    if (ROLL) {
#pragma unroll 1
    for (int g = 0; g < (int)Geo::G; g++) {
        const u32 k2 = Geo::G * w + g;
        u32 tb[R];
#pragma unroll
        for (int u = 1; u <= R; u++) tb[u - 1] = pass3_base<R>(blk, (u32)g, u, t);
        levels3<R, Lay<R>::P3MASK>(x, twc, tb, rp, Geo::T);
#pragma unroll
        for (int c = 0; c < (int)Geo::RS; c++) {
            const u32 kl = (brev_bits((u32)c, R) << 10) | (k2 << 5) | lane;
            if (S0ZERO) orow[kl] = canon3(x[c], rp);
            else orow[((u64)kl << s0) + oblk] = canon3(x[c], rp);
        }
        // rotate the remaining groups down by one
#pragma unroll
        for (int c = 0; c < (int)(Geo::G - 1) * (int)Geo::RS; c++) x[c] = x[c + Geo::RS];
    }
    } else {
#pragma unroll
    for (int g = 0; g < (int)Geo::G; g++) {
        const u32 k2 = Geo::G * w + g;
        u32 tb[R];
#pragma unroll
        for (int u = 1; u <= R; u++) tb[u - 1] = pass3_base<R>(blk, (u32)g, u, t);
        levels3<R, Lay<R>::P3MASK>(x + g * Geo::RS, twc, tb, rp, Geo::T);
#pragma unroll
        for (int c = 0; c < (int)Geo::RS; c++) {
            const u32 kl = (brev_bits((u32)c, R) << 10) | (k2 << 5) | lane;
            if (S0ZERO) orow[kl] = canon3(x[g * Geo::RS + c], rp);
            else orow[((u64)kl << s0) + oblk] = canon3(x[g * Geo::RS + c], rp);
        }
    }
    }
}

// ------------------------------------------------------------------ rows of 2^(14+s0) positions, s0 >= 1: last global level on load
// Global level s0 pairs position p of sub-block 2m (X) with position p of sub-block 2m+1 (Y), twiddle index
// 2^(s0-1) + m.  The CTA that owns sub-block blk = 2m + rank reads BOTH (canonical) operands straight from global memory
// and forms X' = X + wY (rank 0) or Y' = X - wY + 4q (rank 1) itself: one extra Shoup product per position instead of a
// separate read-modify-write pass over the row in HBM.  Results are below 5q; the five levels of pass 1 then reduce at
// level 3 (bound 5 -> 9 -> 13 -> (reduce) 6 -> 10 -> 14; pass 2 reduces at its level 1).
template <int R>
TFB_HD void pass1_cross_global(u64* x, const u64* __restrict__ even, const u64* __restrict__ odd, const u32 rank, const tw_t wc,
                               const Red3& rp, const u32 t) {
#pragma unroll
    for (int a = 0; a < 32; a++) {
        const u64 X = even[a * NttGeo<R>::T + t], Y = odd[a * NttGeo<R>::T + t];
        const u64 tt = shoup_lazy4(Y, wc.w, wc.wp, rp.q, rp.ne, rp.shb);
#ifndef __CUDA_ARCH__
        if (tt >= rp.q4 || X >= rp.q) g_emu_overflow3++;
#endif
        x[a] = rank ? X - tt + rp.q4 : X + tt;
    }
}
TFB_HD void pass1_cross_levels(u64* x, const tw_t* __restrict__ tw, const Red3& rp, const u32 s0, const u32 blk) {
    u32 tb[5];
#pragma unroll
    for (int s = 1; s <= 5; s++) tb[s - 1] = (1u << (s0 + s - 1)) + (blk << (s - 1));
    levels3<5, 0x04>(x, tw, tb, rp);
}
template <int R>
TFB_HD void pass1_store(const u64* x, u64* smem, const u32 t) {
#pragma unroll
    for (int a = 0; a < 32; a++) smem[slot<R>(a, t)] = x[a];
}

// ------------------------------------------------------------------ inverse (pow2_cyc_rings.jl:308-318)
// GS butterfly  X' = X + Y, Y' = (X - Y) w.  Products come back in [0,3q) (shoup_lazy4) and sums double, so a value
// has to be brought back before it passes 16q ~ 2^64.  Bounds in units of q:
//   plain level (RED = false): inputs < 4:  X' = X + Y < 2 max(in),  D = X - Y + 4q < 8,    Y' < 3
//   reducing level (RED = true): inputs < 6: D = X - Y + 8q < 14, and X' = X + Y + tab8[k] with k = floor estimate of
//     (X + Y) / 2^b from the high words only (it can be one short, which leaves X' in (0,3q) instead of (0,2q)); the
//     constant is added inside the 3-input sum.
// From canonical input the schedule plain, plain, reduce, plain, reduce, ... keeps every value below 6q; round 1 reduced
// at EVERY level after the first (one table lookup, one shift, two adds per butterfly on 12 of 13 ladder levels -- now
// on 6 of 13), see the per-pass masks below.
template <bool RED>
TFB_HD void gs_bfly3(u64& X, u64& Y, const tw_t w, const Red3& rp) {
    const u64 x = X, y = Y;
    const u64 off = RED ? 2 * rp.q4 : rp.q4;
    const u64 d = x - y + off;
#ifndef __CUDA_ARCH__
    if (y >= off || (((u128)x + off) >> 64) != 0 || (((u128)x + y) >> 64) != 0) g_emu_overflow3++;
#endif
    if (RED) {
        const u32 k = ((u32)(x >> 32) + (u32)(y >> 32)) >> rp.shb;
#ifndef __CUDA_ARCH__
        if (k > 15) { g_emu_overflow3++; return; }
#endif
        X = x + y + rp.tab8[k];
#ifndef __CUDA_ARCH__
        if (X >= 3 * rp.q) g_emu_overflow3++;
#endif
    } else {
        X = x + y;
#ifndef __CUDA_ARCH__
        if (X >= 6 * rp.q) g_emu_overflow3++;     // a plain level's output must be admissible for a reducing level
#endif
    }
    Y = shoup_lazy4(d, w.w, w.wp, rp.q, rp.ne, rp.shb);
}
// levels LV..FIRST of the inverse ladder; the e-th level executed (e = 0 for level LV) reduces iff bit e of REDMASK is set
template <int LV, int FIRST, u32 REDMASK>
TFB_HD void gs_levels3(u64* x, const tw_t* __restrict__ tw, const u32* tb, const Red3& rp, const u32 js = 1) {
#pragma unroll
    for (int u = LV; u >= FIRST; u--) {
        const int half = (1 << LV) >> u;
#pragma unroll
        for (int j = 0; j < (1 << (u - 1)); j++) {
            const tw_t w = tw[tb[u - 1] + j * js];
#pragma unroll
            for (int k = 0; k < half; k++) {
                if ((REDMASK >> (LV - u)) & 1) gs_bfly3<true>(x[j * 2 * half + k], x[j * 2 * half + k + half], w, rp);
                else gs_bfly3<false>(x[j * 2 * half + k], x[j * 2 * half + k + half], w, rp);
            }
        }
    }
}
// Reduction schedules (bit e = the e-th level executed in the pass).  Pass 3 starts from canonical input: plain, plain,
// reduce, plain (bounds 1 -> 3 -> 6 -> 3 -> 6; fewer levels for smaller rows).  Pass 2 takes anything below 6q: reduce,
// plain, reduce, plain, reduce (-> 3).  Pass 1 takes values below 3q: plain, reduce, plain, reduce (-> 3), so the final
// level's X - Y + 4q is valid; the five sub-block levels of longer rows: plain, reduce, plain, reduce, plain (-> 6 < 16q
// for the canonicalising table).
constexpr u32 INV_P3MASK = 0x4, INV_P2MASK = 0x15, INV_P1MASK = 0xA, INV_P1ALLMASK = 0xA;
// pass 3 (levels 10+R..11): natural-order canonical input from the flat copy in shared memory
template <int R>
TFB_HD void inv_pass3_load(u64* x, const u64* smem, const u32 t) {
    typedef NttGeo<R> Geo;
    const u32 w = t >> 5, lane = t & 31;
#pragma unroll
    for (int g = 0; g < (int)Geo::G; g++)
#pragma unroll
        for (int c = 0; c < (int)Geo::RS; c++) x[g * Geo::RS + c] = smem[(brev_bits((u32)c, R) << 10) | ((Geo::G * w + g) << 5) | lane];
}
template <int R>
TFB_HD void inv_pass3_compute_store(u64* x, u64* smem, const tw_t* __restrict__ itwc, const Red3& rp, const u32 t, const u32 blk = 0) {
    typedef NttGeo<R> Geo;
    const u32 w = t >> 5, lane = t & 31;
    u64* base = smem + slot<R>(brev_bits(lane, 5), brev_bits(Geo::G * w, 5) * Geo::RS);
#pragma unroll
    for (int g = 0; g < (int)Geo::G; g++) {
        const u32 off = brev_bits((u32)g, 5) * Geo::RS;
        u32 tb[R];
#pragma unroll
        for (int u = 1; u <= R; u++) tb[u - 1] = pass3_base<R>(blk, (u32)g, u, t);
        gs_levels3<R, 1, INV_P3MASK>(x + g * Geo::RS, itwc, tb, rp, Geo::T);
#ifdef __CUDA_ARCH__
#pragma unroll
        for (int c = 0; c < (int)Geo::RS; c += 2)
            *reinterpret_cast<ulonglong2*>(base + off + c) = make_ulonglong2(x[g * Geo::RS + c], x[g * Geo::RS + c + 1]);
#else
        for (int c = 0; c < (int)Geo::RS; c++) base[off + c] = x[g * Geo::RS + c];
#endif
    }
}
// pass 2 (levels 10..6)
template <int R>
TFB_HD void inv_pass2(u64* x, u64* smem, const tw_t* __restrict__ itw, const Red3& rp, const u32 t, const u32 s0 = 0, const u32 blk = 0) {
    typedef NttGeo<R> Geo;
    const u32 a2 = t >> R, c2 = t & (Geo::RS - 1);
    u64* base = smem + slot<R>(a2, c2);
#pragma unroll
    for (int b = 0; b < 32; b++) x[b] = base[b * Geo::RS];
    u32 tb[5];
#pragma unroll
    for (int u = 1; u <= 5; u++) tb[u - 1] = (1u << (s0 + 4 + u)) + (blk << (4 + u)) + (a2 << (u - 1));
    gs_levels3<5, 1, INV_P2MASK>(x, itw, tb, rp);
#pragma unroll
    for (int b = 0; b < 32; b++) base[b * Geo::RS] = x[b];
}
// pass 1 (levels 5..1, N^-1 folded into level 1: tn = N^-1, twn = N^-1 psi^-brev(1)); canonical natural-order output
template <int R>
TFB_HD void inv_pass1_load(u64* x, const u64* smem, const u32 t) {
#pragma unroll
    for (int a = 0; a < 32; a++) x[a] = smem[slot<R>(a, t)];
}
template <int R>
TFB_HD void inv_pass1_compute_store(u64* x, u64* __restrict__ orow, const tw_t* __restrict__ itw, const Red3& rp, const u32 t,
                                    const tw_t tn, const tw_t twn) {
    typedef NttGeo<R> Geo;
    u32 tb[5];
#pragma unroll
    for (int s = 1; s <= 5; s++) tb[s - 1] = 1u << (s - 1);
    gs_levels3<5, 2, INV_P1MASK>(x, itw, tb, rp);
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const u64 U = x[k], V = x[k + 16];
#ifndef __CUDA_ARCH__
        if (V >= rp.q4 || (((u128)U + V) >> 64) != 0 || (((u128)U + rp.q4) >> 64) != 0) g_emu_overflow3++;
#endif
        x[k] = shoup_lazy4(U + V, tn.w, tn.wp, rp.q, rp.ne, rp.shb);
        x[k + 16] = shoup_lazy4(U - V + rp.q4, twn.w, twn.wp, rp.q, rp.ne, rp.shb);
    }
#pragma unroll
    for (int a = 0; a < 32; a++) orow[a * Geo::T + t] = canon3i(x[a], rp);
}
// levels 5..1 of a sub-block of a longer row (no N^-1: the global inverse stages follow)
TFB_HD void inv_pass1_levels_all(u64* x, const tw_t* __restrict__ itw, const Red3& rp, const u32 s0, const u32 blk) {
    u32 tb[5];
#pragma unroll
    for (int s = 1; s <= 5; s++) tb[s - 1] = (1u << (s0 + s - 1)) + (blk << (s - 1));
    gs_levels3<5, 1, INV_P1ALLMASK>(x, itw, tb, rp);
}
}  // namespace v3
