// Third-generation forward ladder for N = 2^14 sub-blocks on primes q = 2^60 + e, e < 2^28 (the
// reference's nextprime(2^60 + 1) chains, crt.jl:282-295).  Same transform as ntt_core.cuh
// (pow2_cyc_rings.jl:295-303: c^[k] = sum_j c[j] psi^(j(2k+1)), natural order in and out) and the same
// 512 threads x 32 residues, levels 5+5+4, but built around what bounds the kernel on sm_100a: the
// FMA-heavy pipe (IMAD 2, IMAD.WIDE/IMAD.HI 4 cycles per warp instruction) and the equally half-rate ALU pipe
// (tools/bfly_bench4.cu, tools/pipe_probe5.cu):
//   * approximate-quotient Shoup product shoup_lazy4 (T in [0,4q)): 1 IMAD.WIDE + 2 IMAD.HI instead of 4 IMAD.WIDE
//     for the quotient, and q's shape saves one more IMAD.WIDE in the tail;
//   * values grow by 4q per level; X is brought back to (0,2q) at levels 4 of pass 1, 1 and 4 of passes 2 and 3
//     by  x + (q - floor(x/2^60) q)  with the constant read from a 16-entry shared-memory table and the addition
//     FUSED into the butterfly's own 3-input adds (X' = x + c2[k] + T, Y' = x + c3[k] - T, c3 = c2 + 4q): a reduction
//     costs one shift, one address and one 128-bit shared load, nothing on the FMA-heavy pipe;
//   * the row sits in shared memory SKEWED by 2 words every 4 rows of 512 (slot(a,i) = 512a + 2(a>>2) + i) instead
//     of XOR-swizzled: every access of the three passes is then base register + immediate (no per-access address
//     arithmetic), 64-bit accesses of passes 1-2 and the 128-bit loads of pass 3 stay bank-conflict-free, and the
//     TMA bulk copies (one per 4 KiB row of 512) keep their 16-byte alignment.
// __host__ __device__ like ntt_core.cuh so tests/emu runs the same index logic on the CPU.
#pragma once
#include "ntt_core.cuh"

namespace v3 {
constexpr int R = 4;
typedef NttGeo<R> Geo;                      // N = 2^14, T = 512, RS = 16, G = 2
constexpr u32 ROW_WORDS = Geo::N + 16;      // skewed row buffer
constexpr u32 ROW_BYTES = ROW_WORDS * 8;
TFB_HD u32 skew(const u32 a) { return 2 * (a >> 2); }
TFB_HD u32 slot(const u32 a, const u32 idx) { return a * Geo::T + skew(a) + idx; }

struct __align__(16) redent_t {
    u64 c2, c3;   // c2 = q - k q, c3 = c2 + 4q  (mod 2^64), k = 0..15
};
struct Red3 {
    u64 q, q4;
    u32 ne, shb;          // 2^32 - e, b - 32
    const redent_t* tab;  // this prime's 16 entries
};
TFB_HD void fill_redtab(redent_t* tab, const u64 q) {
    for (u32 k = 0; k < 16; k++) {
        tab[k].c2 = q - (u64)k * q;
        tab[k].c3 = tab[k].c2 + 4 * q;
    }
}
TFB_HD u32 floor_log2_u64(const u64 q) {
#ifdef __CUDA_ARCH__
    return 63u - (u32)__clzll((long long)q);
#else
    return 63u - (u32)__builtin_clzll(q);
#endif
}
TFB_HD Red3 make_red3(const u64 q, const redent_t* tab) {
    Red3 r;
    r.q = q;
    r.q4 = 4 * q;
    r.ne = 0u - (u32)(q - (1ull << 60)); r.shb = floor_log2_u64(q) - 32;
    r.tab = tab;
    return r;
}

#ifndef __CUDA_ARCH__
static unsigned long long g_emu_overflow3 = 0;   // tests/emu: lazy-range violations (must stay 0)
#endif

// CT butterfly, X in [0,16q) if RED else X + 4q < 2^64;  Y any 64-bit value
template <bool RED>
TFB_HD void bfly3(u64& X, u64& Y, const tw_t w, const Red3& rp) {
    const u64 t = shoup_lazy4<28>(Y, w.w, w.wp, rp.q, rp.ne, rp.shb);
    const u64 x = X;
    if (RED) {
        const redent_t c = rp.tab[x >> 60];
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
// any v < 16q -> canonical
TFB_HD u64 canon3(const u64 v, const Red3& rp) { return csub(v + rp.tab[v >> 60].c2, rp.q); }

// pass 1 (levels 1..5): thread t holds a = 0..31 at index t; canonical input, bound 1 -> 13 -> (reduce) 6 -> 10
TFB_HD void pass1(u64* x, u64* smem, const tw_t* __restrict__ tw, const Red3& rp, const u32 t, const u32 s0, const u32 blk) {
#pragma unroll
    for (int a = 0; a < 32; a++) x[a] = smem[slot(a, t)];
    u32 tb[5];
#pragma unroll
    for (int s = 1; s <= 5; s++) tb[s - 1] = (1u << (s0 + s - 1)) + (blk << (s - 1));
    levels3<5, 0x08>(x, tw, tb, rp);
#pragma unroll
    for (int a = 0; a < 32; a++) smem[slot(a, t)] = x[a];
}
// pass 2 (levels 6..10): thread (a2 = t >> 4, c2 = t & 15) holds b = 0..31; bound 10 -> (reduce) 6 -> 10 -> 14 -> (reduce) 6 -> 10
TFB_HD void pass2(u64* x, u64* smem, const tw_t* __restrict__ tw, const Red3& rp, const u32 t, const u32 s0, const u32 blk) {
    const u32 a2 = t >> R, c2 = t & (Geo::RS - 1);
    u64* base = smem + slot(a2, c2);
#pragma unroll
    for (int b = 0; b < 32; b++) x[b] = base[b * Geo::RS];
    u32 tb[5];
#pragma unroll
    for (int u = 1; u <= 5; u++) tb[u - 1] = (1u << (s0 + 4 + u)) + (blk << (4 + u)) + (a2 << (u - 1));
    levels3<5, 0x09>(x, tw, tb, rp);
#pragma unroll
    for (int b = 0; b < 32; b++) base[b * Geo::RS] = x[b];
}
// pass 3 (levels 11..14): thread (warp w, lane l) holds, for a3 = brev5(l), the groups b3 = brev5(2w + g), g = 0,1, all c
TFB_HD void pass3_load(u64* x, const u64* smem, const u32 t) {
    const u32 w = t >> 5, lane = t & 31;
    const u64* base = smem + slot(brev_bits(lane, 5), brev_bits(2 * w, 5) * Geo::RS);
#pragma unroll
    for (int g = 0; g < 2; g++) {
#ifdef __CUDA_ARCH__
#pragma unroll
        for (int c = 0; c < 16; c += 2) {
            const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(base + g * 16 * Geo::RS + c);
            x[g * 16 + c] = v.x;
            x[g * 16 + c + 1] = v.y;
        }
#else
        for (int c = 0; c < 16; c++) x[g * 16 + c] = base[g * 16 * Geo::RS + c];
#endif
    }
}
// bound 10 -> (reduce) 6 -> 10 -> 14 -> (reduce) 6 -> canonical; natural-order coalesced stores
template <bool S0ZERO>
TFB_HD void pass3_compute_store(u64* x, u64* __restrict__ orow, const tw_t* __restrict__ twc, const Red3& rp,
                                const u32 t, const u32 s0, const u32 blk) {
    const u32 w = t >> 5, lane = t & 31;
    const u32 oblk = S0ZERO ? 0 : brev_bits(blk, (int)s0);
#pragma unroll
    for (int g = 0; g < 2; g++) {
        const u32 k2 = 2 * w + g;
        u32 tb[R];
#pragma unroll
        for (int u = 1; u <= R; u++) tb[u - 1] = pass3_base<R>(blk, (u32)g, u, t);
        levels3<R, 0x09>(x + g * 16, twc, tb, rp, Geo::T);
#pragma unroll
        for (int c = 0; c < 16; c++) {
            const u32 kl = (brev_bits((u32)c, R) << 10) | (k2 << 5) | lane;
            if (S0ZERO) orow[kl] = canon3(x[g * 16 + c], rp);
            else orow[((u64)kl << s0) + oblk] = canon3(x[g * 16 + c], rp);
        }
    }
}

// ------------------------------------------------------------------ inverse (pow2_cyc_rings.jl:308-318)
// GS butterfly  X' = X + Y, Y' = (X - Y) w.  Products come back in [0,4q) (shoup_lazy4), so every level after the
// first brings the sum back with the same table: k is estimated from the high words only (it can be one short of
// floor((X+Y)/2^60), which leaves X' in (0,3q) instead of (0,2q)), and the constant is added inside the 3-input
// sum.  Inputs of a reducing level are < 4q, of the first level canonical; X - Y + 4q < 8q.
template <bool RED>
TFB_HD void gs_bfly3(u64& X, u64& Y, const tw_t w, const Red3& rp) {
    const u64 x = X, y = Y;
    const u64 d = x - y + rp.q4;
#ifndef __CUDA_ARCH__
    if (y >= rp.q4 || (((u128)x + rp.q4) >> 64) != 0 || (((u128)x + y) >> 64) != 0) g_emu_overflow3++;
#endif
    if (RED) {
        const u32 k = ((u32)(x >> 32) + (u32)(y >> 32)) >> 28;
        X = x + y + rp.tab[k].c2;
#ifndef __CUDA_ARCH__
        if (X >= 3 * rp.q || k > 15) g_emu_overflow3++;
#endif
    } else {
        X = x + y;
    }
    Y = shoup_lazy4<28>(d, w.w, w.wp, rp.q, rp.ne, rp.shb);
}
// levels LV..FIRST of the inverse ladder; the level executed first reduces iff RED_TOP, all later ones always
template <int LV, int FIRST, bool RED_TOP>
TFB_HD void gs_levels3(u64* x, const tw_t* __restrict__ tw, const u32* tb, const Red3& rp, const u32 js = 1) {
#pragma unroll
    for (int u = LV; u >= FIRST; u--) {
        const int half = (1 << LV) >> u;
#pragma unroll
        for (int j = 0; j < (1 << (u - 1)); j++) {
            const tw_t w = tw[tb[u - 1] + j * js];
#pragma unroll
            for (int k = 0; k < half; k++) {
                if (u == LV && !RED_TOP) gs_bfly3<false>(x[j * 2 * half + k], x[j * 2 * half + k + half], w, rp);
                else gs_bfly3<true>(x[j * 2 * half + k], x[j * 2 * half + k + half], w, rp);
            }
        }
    }
}
// pass 3 (levels 14..11): natural-order canonical input from the flat copy in shared memory
TFB_HD void inv_pass3_load(u64* x, const u64* smem, const u32 t) {
    const u32 w = t >> 5, lane = t & 31;
#pragma unroll
    for (int g = 0; g < 2; g++)
#pragma unroll
        for (int c = 0; c < 16; c++) x[g * 16 + c] = smem[(brev_bits((u32)c, R) << 10) | ((2 * w + g) << 5) | lane];
}
TFB_HD void inv_pass3_compute_store(u64* x, u64* smem, const tw_t* __restrict__ itwc, const Red3& rp, const u32 t) {
    const u32 w = t >> 5, lane = t & 31;
    u64* base = smem + slot(brev_bits(lane, 5), brev_bits(2 * w, 5) * Geo::RS);
#pragma unroll
    for (int g = 0; g < 2; g++) {
        u32 tb[R];
#pragma unroll
        for (int u = 1; u <= R; u++) tb[u - 1] = pass3_base<R>(0, (u32)g, u, t);
        gs_levels3<R, 1, false>(x + g * 16, itwc, tb, rp, Geo::T);
#ifdef __CUDA_ARCH__
#pragma unroll
        for (int c = 0; c < 16; c += 2)
            *reinterpret_cast<ulonglong2*>(base + g * 16 * Geo::RS + c) = make_ulonglong2(x[g * 16 + c], x[g * 16 + c + 1]);
#else
        for (int c = 0; c < 16; c++) base[g * 16 * Geo::RS + c] = x[g * 16 + c];
#endif
    }
}
// pass 2 (levels 10..6)
TFB_HD void inv_pass2(u64* x, u64* smem, const tw_t* __restrict__ itw, const Red3& rp, const u32 t) {
    const u32 a2 = t >> R, c2 = t & (Geo::RS - 1);
    u64* base = smem + slot(a2, c2);
#pragma unroll
    for (int b = 0; b < 32; b++) x[b] = base[b * Geo::RS];
    u32 tb[5];
#pragma unroll
    for (int u = 1; u <= 5; u++) tb[u - 1] = (1u << (4 + u)) + (a2 << (u - 1));
    gs_levels3<5, 1, true>(x, itw, tb, rp);
#pragma unroll
    for (int b = 0; b < 32; b++) base[b * Geo::RS] = x[b];
}
// pass 1 (levels 5..1, N^-1 folded into level 1: tn = N^-1, twn = N^-1 psi^-brev(1)); canonical natural-order output
TFB_HD void inv_pass1_load(u64* x, const u64* smem, const u32 t) {
#pragma unroll
    for (int a = 0; a < 32; a++) x[a] = smem[slot(a, t)];
}
TFB_HD void inv_pass1_compute_store(u64* x, u64* __restrict__ orow, const tw_t* __restrict__ itw, const Red3& rp, const u32 t,
                                    const tw_t tn, const tw_t twn) {
    u32 tb[5];
#pragma unroll
    for (int s = 1; s <= 5; s++) tb[s - 1] = 1u << (s - 1);
    gs_levels3<5, 2, true>(x, itw, tb, rp);
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const u64 U = x[k], V = x[k + 16];
#ifndef __CUDA_ARCH__
        if (V >= rp.q4 || (((u128)U + V) >> 64) != 0 || (((u128)U + rp.q4) >> 64) != 0) g_emu_overflow3++;
#endif
        x[k] = shoup_lazy4<28>(U + V, tn.w, tn.wp, rp.q, rp.ne, rp.shb);
        x[k + 16] = shoup_lazy4<28>(U - V + rp.q4, twn.w, twn.wp, rp.q, rp.ne, rp.shb);
    }
#pragma unroll
    for (int a = 0; a < 32; a++) orow[a * Geo::T + t] = canon3(x[a], rp);
}
}  // namespace v3
