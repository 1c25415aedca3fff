// Second-generation row-resident negacyclic NTT for N = 2^14 (and 2^14 sub-blocks
// of longer rows): 1024 threads x 16 residues per prime-row.
//
// Same contract as ntt_core.cuh (pow2_cyc_rings.jl:295-318, natural order in and
// out); different mapping, chosen from the round-1 ncu capture of the 512x32 kernel
// (IPC 0.46 with 4 warps/scheduler, 127 KB of straight-line SASS, exposed row load):
//   * 16 residues per thread -> 64 registers -> 32 warps per SM;
//   * position p = (a:4 | b:4 | c:4 | d:2); three in-place radix-16 passes (over a,
//     b, c) run through ONE looped code body (I-cache resident), then a 2-level pass
//     over d whose stores are already in natural order and coalesced;
//   * the row lives in shared memory (flat as TMA delivers it, XOR-swizzled from
//     pass b on so every access pattern is bank-conflict free);
//   * unified addressing of a middle pass: addr(r) = (r>>2)*S4 + w[r&3].
//
// All functions are __host__ __device__ so tests/emu can run them thread by thread.
#pragma once
#include "ntt_core.cuh"

namespace v2 {

constexpr int LOGN = 14;
constexpr u32 N = 1u << LOGN;
constexpr u32 T = N / 16;  // 1024 threads

struct PassCfg {
    u32 S4;      // element stride of (r >> 2)
    u32 wl[4];   // load offsets for r & 3
    u32 ws[4];   // store offsets for r & 3
    u32 tb[4];   // twiddle index base of levels 1..4
};

// swizzled slot of position p:  low 4 bits ^= top4(p) ^ (bits[7:6](p) << 2)
TFB_HD u32 swz2(u32 p) { return p ^ ((p >> 10) ^ (((p >> 6) & 3u) << 2)); }

// middle pass k (0: field a = bits 13..10, 1: b = 9..6, 2: c = 5..2) of thread t
TFB_HD PassCfg make_cfg(const int k, const u32 t, const u32 s0, const u32 blk) {
    PassCfg c;
    u32 thi;
    if (k == 0) {
        thi = 0;
        c.S4 = 4096;
#pragma unroll
        for (int i = 0; i < 4; i++) c.wl[i] = c.ws[i] = t + (u32)i * 1024u;
    } else if (k == 1) {
        const u32 tlo = t & 63u, a = t >> 6;
        thi = a;
        const u32 P0 = a << 10;
        c.S4 = 256;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            c.wl[i] = (P0 | tlo) + (u32)i * 64u;                                        // flat
            c.ws[i] = (P0 | (tlo & ~15u)) + (u32)i * 64u + ((tlo & 15u) ^ a ^ ((u32)i << 2));  // swizzled
        }
    } else {
        const u32 d = t & 3u;
        thi = t >> 2;  // (a:4 | b:4)
        const u32 m = (thi >> 4) ^ ((thi & 3u) << 2);
        c.S4 = 16;
#pragma unroll
        for (int i = 0; i < 4; i++) c.wl[i] = c.ws[i] = (thi << 6) + ((((u32)i << 2) | d) ^ m);
    }
#pragma unroll
    for (int u = 1; u <= 4; u++) c.tb[u - 1] = (1u << (s0 + 4 * k + u - 1)) + (blk << (4 * k + u - 1)) + (thi << (u - 1));
    return c;
}

// ------------------------------------------------------------------ forward
TFB_HD void fwd_mid_load(u64* x, const u64* smem, const PassCfg& c) {
#pragma unroll
    for (int r = 0; r < 16; r++) x[r] = smem[(u32)(r >> 2) * c.S4 + c.wl[r & 3]];
}
TFB_HD void fwd_mid_store(const u64* x, u64* smem, const PassCfg& c) {
#pragma unroll
    for (int r = 0; r < 16; r++) smem[(u32)(r >> 2) * c.S4 + c.ws[r & 3]] = x[r];
}
// MODE 1 bounds: every radix-16 pass reduces X at its first level (-> 2q) and leaves
// [0,10q); the 2-level output pass adds 4q more (<= 14q) before the final reduction.
template <int MODE>
TFB_HD void fwd_mid_compute(u64* x, const tw_t* __restrict__ tw, const PassCfg& c, const RedParams& rp) {
    ct_levels_m<4, MODE, true>(x, tw, c.tb, rp);
}

// last pass: group g of thread t covers natural indices kr = g*1024 + t (+ e-part)
TFB_HD u32 last_rest(const u32 t, const int g) { return brev_bits((u32)g * T + t, 12); }
TFB_HD u32 last_slot(const u32 rest, const int e) {
    const u32 m = (rest >> 8) ^ (((rest >> 4) & 3u) << 2);
    return ((rest >> 2) << 4) + (((((rest & 3u) << 2) | (u32)e)) ^ m);
}
TFB_HD void fwd_last_load(u64* x, const u64* smem, const u32 t) {
#pragma unroll
    for (int g = 0; g < 4; g++) {
        const u32 rest = last_rest(t, g);
#pragma unroll
        for (int e = 0; e < 4; e++) x[g * 4 + e] = smem[last_slot(rest, e)];
    }
}
template <int MODE>
TFB_HD void fwd_last_compute_store(u64* x, u64* __restrict__ out, const tw_t* __restrict__ tw, const RedParams& rp,
                                   const u32 t, const u32 s0, const u32 blk) {
    const u32 oblk = brev_bits(blk, (int)s0);
#pragma unroll
    for (int g = 0; g < 4; g++) {
        const u32 rest = last_rest(t, g);
        const tw_t w1 = tw[(1u << (s0 + 12)) + (blk << 12) + rest];
        const u32 b2 = (1u << (s0 + 13)) + (blk << 13) + (rest << 1);
        const tw_t w20 = tw[b2], w21 = tw[b2 + 1];
        u64* y = x + g * 4;
        ct_bfly_m<MODE, false>(y[0], y[2], w1, rp);
        ct_bfly_m<MODE, false>(y[1], y[3], w1, rp);
        ct_bfly_m<MODE, false>(y[0], y[1], w20, rp);
        ct_bfly_m<MODE, false>(y[2], y[3], w21, rp);
        const u32 kr = (u32)g * T + t;
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const u32 kl = (brev_bits((u32)e, 2) << 12) | kr;
            out[((u64)kl << s0) + oblk] = canon<MODE>(y[e], rp);
        }
    }
}

// ------------------------------------------------------------------ inverse
// first inverse pass: natural-order source `src` (flat shared copy of the row, or
// global memory when the row is a strided sub-block), GS levels 2,1 over d
TFB_HD void inv_first_load(u64* x, const u64* __restrict__ src, const u32 t, const u32 s0, const u32 blk) {
    const u32 oblk = brev_bits(blk, (int)s0);
#pragma unroll
    for (int g = 0; g < 4; g++) {
        const u32 kr = (u32)g * T + t;
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const u32 kl = (brev_bits((u32)e, 2) << 12) | kr;
            x[g * 4 + e] = src[((u64)kl << s0) + oblk];
        }
    }
}
TFB_HD void inv_first_compute_store(u64* x, u64* smem, const tw_t* __restrict__ itw, const u64 q, const u32 t,
                                    const u32 s0, const u32 blk) {
    const u64 q2 = 2 * q;
#pragma unroll
    for (int g = 0; g < 4; g++) {
        const u32 rest = last_rest(t, g);
        const tw_t w1 = itw[(1u << (s0 + 12)) + (blk << 12) + rest];
        const u32 b2 = (1u << (s0 + 13)) + (blk << 13) + (rest << 1);
        const tw_t w20 = itw[b2], w21 = itw[b2 + 1];
        u64* y = x + g * 4;
        gs_bfly(y[0], y[1], w20, q, q2);
        gs_bfly(y[2], y[3], w21, q, q2);
        gs_bfly(y[0], y[2], w1, q, q2);
        gs_bfly(y[1], y[3], w1, q, q2);
#pragma unroll
        for (int e = 0; e < 4; e++) smem[last_slot(rest, e)] = y[e];
    }
}
// inverse middle pass: loads with the forward pass's STORE layout, stores with its
// LOAD layout (the mirror image)
TFB_HD void inv_mid_load(u64* x, const u64* smem, const PassCfg& c) {
#pragma unroll
    for (int r = 0; r < 16; r++) x[r] = smem[(u32)(r >> 2) * c.S4 + c.ws[r & 3]];
}
TFB_HD void inv_mid_store(const u64* x, u64* smem, const PassCfg& c) {
#pragma unroll
    for (int r = 0; r < 16; r++) smem[(u32)(r >> 2) * c.S4 + c.wl[r & 3]] = x[r];
}
TFB_HD void inv_mid_compute(u64* x, const tw_t* __restrict__ itw, const PassCfg& c, const u64 q) {
    gs_levels<4, 1>(x, itw, c.tb, q, 2 * q);
}
// final inverse pass (field a): levels 4..2 then level 1 with N^-1 folded in (s0 == 0),
// canonical coalesced store to global
TFB_HD void inv_final_compute_store(u64* x, u64* __restrict__ out, const tw_t* __restrict__ itw, const PassCfg& c,
                                    const u64 q, const u32 t, const u32 s0, const tw_t tn, const tw_t twn) {
    const u64 q2 = 2 * q;
    if (s0 == 0) {
        gs_levels<4, 2>(x, itw, c.tb, q, q2);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const u64 U = x[k], V = x[k + 8];
            x[k] = shoup_lazy(U + V, tn.w, tn.wp, q);
            x[k + 8] = shoup_lazy(U - V + q2, twn.w, twn.wp, q);
        }
    } else {
        gs_levels<4, 1>(x, itw, c.tb, q, q2);
    }
#pragma unroll
    for (int r = 0; r < 16; r++) out[(u32)r * 1024u + t] = csub(x[r], q);
}

}  // namespace v2
