// Per-thread bodies of the row-resident negacyclic NTT kernels (N = 2^(10+R),
// R = 0..4: one CTA of N/32 threads per prime-row, 32 residues per thread in
// registers, two shared-memory exchanges).
//
// Computes exactly what pow2_cyc_rings.jl:295-303 (nntt) and :308-318 (inntt)
// define -- c^[k] = sum_j c[j] psi^(j(2k+1)), natural order in and out -- but as a
// merged (psi folded into the twiddles) Cooley-Tukey / Gentleman-Sande ladder
// with Harvey lazy butterflies; FourierTransforms.jl's CTPlan and the per-call
// twiddle rebuild (pow2_cyc_rings.jl:298-301) have no counterpart here.
//
// Position bits of a residue inside the row: p = (a : 5 | b : 5 | c : R).
//   pass 1 (stages 1..5)      thread t=(b,c) holds a = 0..31     (stride N/32)
//   pass 2 (stages 6..10)     thread (a,c)   holds b = 0..31
//   pass 3 (stages 11..10+R)  thread (warp w, lane l) holds, for a = brev5(l),
//                             the G = 32/2^R groups b = brev5(w*G+g), all c.
// The merged CT ladder leaves c^[brev(p)] at position p; with the pass-3 mapping
// the natural index is k = brevR(c)<<10 | (w*G+g)<<5 | l, so a warp stores 32
// consecutive words: natural-order output costs no extra permutation pass.
//
// The functions are __host__ __device__ so tests can run the *same* index logic
// thread-by-thread on the CPU (tests/emu); the product only uses them from the
// kernels in ntt_kernels.cu.
#pragma once
#include "modarith.cuh"

template <int R>
struct NttGeo {
    static constexpr int LOGN = 10 + R;
    static constexpr u32 N = 1u << LOGN;
    static constexpr u32 T = N / 32;   // threads per row
    static constexpr u32 RS = 1u << R; // pass-3 group size
    static constexpr u32 G = 32 / RS;  // pass-3 groups per thread
};

TFB_HD u32 brev_bits(u32 x, int bits) {
    if (bits == 0) return 0;
#ifdef __CUDA_ARCH__
    return __brev(x) >> (32 - bits);
#else
    u32 r = 0;
    for (int i = 0; i < bits; i++)
        if (x >> i & 1) r |= 1u << (bits - 1 - i);
    return r;
#endif
}

// shared-memory slot of position (a, idx=(b,c)); the XOR keeps all three access
// patterns (lanes along t, along (a_lo,c), along a) on distinct 8-byte banks.
template <int R>
TFB_HD u32 swz(u32 a, u32 idx) {
    return a * NttGeo<R>::T + (idx ^ (a >> 1));
}

// Pass-3 twiddles are read from a second copy of the table laid out in THREAD order
// (host: permute_pass3 in tables.h): the twiddle of thread t, group g, level u, block j
// of sub-block blk sits at  blk*N + (g*(RS-1) + 2^(u-1)-1 + j)*T + t,  so a warp's
// 128-bit loads are contiguous (4 L1 wavefronts instead of 32 for the natural
// psi^brev(k) order, whose pass-3 entries are 32*2^(u-1) apart between lanes).
template <int R>
TFB_HD u32 pass3_base(const u32 blk, const u32 g, const int u, const u32 t) {
    return blk * NttGeo<R>::N + (g * (NttGeo<R>::RS - 1) + (1u << (u - 1)) - 1) * NttGeo<R>::T + t;
}

// Harvey lazy butterflies.  CT: X,Y in [0,4q) -> [0,4q).  GS: X,Y in [0,2q) -> [0,2q).
TFB_HD void ct_bfly(u64& X, u64& Y, const tw_t w, const u64 q, const u64 q2) {
    u64 x = csub(X, q2);
    u64 t = shoup_lazy(Y, w.w, w.wp, q);
    X = x + t;
    Y = x - t + q2;
}
TFB_HD void gs_bfly(u64& X, u64& Y, const tw_t w, const u64 q, const u64 q2) {
    u64 s = X + Y;
    u64 d = X - Y + q2;
    X = csub(s, q2);
    Y = shoup_lazy(d, w.w, w.wp, q);
}

// ---- range policy of the forward ladder ---------------------------------
// MODE 0 (Harvey): X is brought back to [0,2q) before every butterfly; needs q < 2^62.
// MODE 1 (lazy):   for primes q = 2^b + e with 0 <= e <= 2^b/16 and 15q < 2^64 -- what
//   nextprime(2^logq + 1) of the reference constructor yields (crt.jl:282-295) --
//   values are left to grow by 2q per level ([0,Bq), B <= 14) and X is reduced once
//   per pass with  x - (x >> b) q + q  in [q - 14e, 2q)  (5 instructions instead
//   of a compare/select subtract at every level).  Shoup products accept any 64-bit Y.
// MODE 2 (lazy, approximate quotient): primes q = 2^60 + e, e < 2^28.  Shoup products use shoup_lazy4
//   (T in [0,4q)), values grow by 4q per level and the X operands are reduced at levels 4 of pass 1,
//   1 and 4 of passes 2 and 3 with a 16-entry table  x -> (x mod 2^60) + (q - (x >> 60) e)  in (0,2q):
//   5 ALU/LSU instructions, nothing on the FMA-heavy pipe that bounds the kernel.
#ifndef __CUDA_ARCH__
static unsigned long long g_emu_overflow = 0;
#endif
struct RedParams {
    u64 q, q2, nq;  // q, 2q, -q (mod 2^64)
    u32 sh;         // b = floor(log2 q)
    u32 ne, shb;    // MODE 2: 2^32 - e, b - 32
    u64 q4;         // MODE 2: 4q
    const u64* tab; // MODE 2: tab[k] = q - k e, k = 0..15 (shared memory in the kernels)
};
TFB_HD RedParams make_red(const u64 q, const u32 sh) {
    RedParams r;
    r.q = q;
    r.q2 = 2 * q;
    r.nq = 0 - q;
    r.sh = sh;
    r.ne = 0; r.shb = 0; r.q4 = 0; r.tab = nullptr;
    return r;
}
TFB_HD RedParams make_red2(const u64 q, const u32 sh, const u64* tab) {
    RedParams r = make_red(q, sh);
    r.ne = 0u - (u32)(q - (1ull << sh));
    r.shb = sh - 32;
    r.q4 = 4 * q;
    r.tab = tab;
    return r;
}
// x in [0,16q) -> (0,2q), same residue (q = 2^60 + e)
TFB_HD u64 reduce_tab(const u64 x, const RedParams& rp) {
    return (x & 0x0fffffffffffffffull) + rp.tab[x >> 60];
}
TFB_HD u64 reduce_shift(const u64 x, const RedParams& rp) {
    const u64 k = x >> rp.sh;
    return x + rp.q + k * rp.nq;
}
template <int MODE>
TFB_HD u64 canon(const u64 v, const RedParams& rp) {
    if (MODE == 0) return csub(csub(v, rp.q2), rp.q);
    if (MODE == 2) return csub(reduce_tab(v, rp), rp.q);
    return csub(reduce_shift(v, rp), rp.q);
}
template <int MODE, bool RED>
TFB_HD void ct_bfly_m(u64& X, u64& Y, const tw_t w, const RedParams& rp) {
    u64 x = X;
    if (MODE == 2) {
        if (RED) x = reduce_tab(x, rp);
        const u64 t = shoup_lazy4<28>(Y, w.w, w.wp, rp.q, rp.ne);
#ifndef __CUDA_ARCH__
        if (t >= rp.q4 || x + t < x || (((u128)x + rp.q4 - t) >> 64) != 0) g_emu_overflow++;   // tests/emu: the lazy bounds must hold
#endif
        X = x + t;
        Y = x - t + rp.q4;
        return;
    }
    if (MODE == 0) x = csub(x, rp.q2);
    else if (RED) x = reduce_shift(x, rp);
    const u64 t = shoup_lazy(Y, w.w, w.wp, rp.q);
    X = x + t;
    Y = x - t + rp.q2;
}
// LV levels; in MODE 1 the X operands are reduced at level 1 iff RED_FIRST
template <int LV, int MODE, bool RED_FIRST>
TFB_HD void ct_levels_m(u64* x, const tw_t* __restrict__ tw, const u32* tb, const RedParams& rp, const u32 js = 1) {
#pragma unroll
    for (int u = 1; u <= LV; u++) {
        const int half = (1 << LV) >> u;
#pragma unroll
        for (int j = 0; j < (1 << (u - 1)); j++) {
            const tw_t w = tw[tb[u - 1] + j * js];
#pragma unroll
            for (int k = 0; k < half; k++) {
                if ((u == 1 && RED_FIRST) || (MODE == 2 && u == 4)) ct_bfly_m<MODE, true>(x[j * 2 * half + k], x[j * 2 * half + k + half], w, rp);
                else ct_bfly_m<MODE, false>(x[j * 2 * half + k], x[j * 2 * half + k + half], w, rp);
            }
        }
    }
}

// LV radix-2 CT levels over CNT=2^LV consecutive registers x[off..off+CNT);
// level u (1..LV) block j uses twiddle tw[base(u) + j], base(u) = (lead << (u-1)) + ofs(u)
// where the caller folds everything into `tb[u-1]`.
template <int LV>
TFB_HD void ct_levels(u64* x, const tw_t* __restrict__ tw, const u32* tb, const u64 q, const u64 q2) {
#pragma unroll
    for (int u = 1; u <= LV; u++) {
        const int half = (1 << LV) >> u;
#pragma unroll
        for (int j = 0; j < (1 << (u - 1)); j++) {
            const tw_t w = tw[tb[u - 1] + j];
#pragma unroll
            for (int k = 0; k < half; k++) ct_bfly(x[j * 2 * half + k], x[j * 2 * half + k + half], w, q, q2);
        }
    }
}
// inverse order of the same ladder (levels LV..FIRST), GS butterflies
template <int LV, int FIRST>
TFB_HD void gs_levels(u64* x, const tw_t* __restrict__ tw, const u32* tb, const u64 q, const u64 q2, const u32 js = 1) {
#pragma unroll
    for (int u = LV; u >= FIRST; u--) {
        const int half = (1 << LV) >> u;
#pragma unroll
        for (int j = 0; j < (1 << (u - 1)); j++) {
            const tw_t w = tw[tb[u - 1] + j * js];
#pragma unroll
            for (int k = 0; k < half; k++) gs_bfly(x[j * 2 * half + k], x[j * 2 * half + k + half], w, q, q2);
        }
    }
}

// ------------------------------------------------------------------ forward
// `in` points at the sub-block (Nsub = N contiguous positions); s0 = stages
// already applied to the whole row (0 unless the row is longer than 2^14),
// blk = index of this sub-block at level s0.
template <int R, int MODE>
TFB_HD void fwd_phaseA(u64* x, const u64* __restrict__ in, u64* smem, const tw_t* __restrict__ tw,
                       const RedParams& rp, const u32 t, const u32 s0, const u32 blk) {
    typedef NttGeo<R> Geo;
#pragma unroll
    for (int a = 0; a < 32; a++) x[a] = in[a * Geo::T + t];
    u32 tb[5];
#pragma unroll
    for (int s = 1; s <= 5; s++) tb[s - 1] = (1u << (s0 + s - 1)) + (blk << (s - 1));
    ct_levels_m<5, MODE, false>(x, tw, tb, rp);   // canonical input: bound 1 -> 11
#pragma unroll
    for (int a = 0; a < 32; a++) smem[swz<R>(a, t)] = x[a];
}

template <int R, int MODE>
TFB_HD void fwd_phaseB(u64* x, u64* smem, const tw_t* __restrict__ tw, const RedParams& rp, const u32 t,
                       const u32 s0, const u32 blk) {
    typedef NttGeo<R> Geo;
    const u32 a2 = t >> R, c2 = t & (Geo::RS - 1);
#pragma unroll
    for (int b = 0; b < 32; b++) x[b] = smem[swz<R>(a2, b * Geo::RS + c2)];
    u32 tb[5];
#pragma unroll
    for (int u = 1; u <= 5; u++) tb[u - 1] = (1u << (s0 + 4 + u)) + (blk << (4 + u)) + (a2 << (u - 1));
    ct_levels_m<5, MODE, true>(x, tw, tb, rp);    // bound 11 -> reduce -> 12
#pragma unroll
    for (int b = 0; b < 32; b++) smem[swz<R>(a2, b * Geo::RS + c2)] = x[b];
}

// out points at the row base; the natural index of local position p is
// (brev(p) << s0) + brev_s0(blk)
template <int R, int MODE>
TFB_HD void fwd_phaseC(u64* x, u64* __restrict__ out, const u64* smem, const tw_t* __restrict__ twc,
                       const RedParams& rp, const u32 t, const u32 s0, const u32 blk) {
    typedef NttGeo<R> Geo;
    const u32 w = t >> 5, lane = t & 31;
    const u32 a3 = brev_bits(lane, 5);
    const u32 oblk = brev_bits(blk, (int)s0);
#pragma unroll
    for (int g = 0; g < (int)Geo::G; g++) {
        const u32 k2 = w * Geo::G + g;
        const u32 b3 = brev_bits(k2, 5);
#pragma unroll
        for (int c = 0; c < (int)Geo::RS; c++) x[g * Geo::RS + c] = smem[swz<R>(a3, b3 * Geo::RS + c)];
        if (R > 0) {
            u32 tb[R > 0 ? R : 1];
#pragma unroll
            for (int u = 1; u <= R; u++) tb[u - 1] = pass3_base<R>(blk, (u32)g, u, t);
            ct_levels_m<R, MODE, true>(x + g * Geo::RS, twc, tb, rp, Geo::T);  // bound 12 -> reduce -> <= 10
        }
#pragma unroll
        for (int c = 0; c < (int)Geo::RS; c++) {
            const u32 kl = (brev_bits((u32)c, R) << 10) | (k2 << 5) | lane;
            out[((u64)kl << s0) + oblk] = canon<MODE>(x[g * Geo::RS + c], rp);
        }
    }
}

// ------------------------------------------------------------------ inverse
// itw = inverse table (psi^-brev(idx)); scale: N^-1 folded into the last level
// when s0 == 0 (tn = Shoup pair of N^-1, twn = Shoup pair of N^-1 * itw[1]).
template <int R>
TFB_HD void inv_phaseC(u64* x, const u64* __restrict__ in, u64* smem, const tw_t* __restrict__ itwc,
                       const u64 q, const u32 t, const u32 s0, const u32 blk) {
    typedef NttGeo<R> Geo;
    const u64 q2 = 2 * q;
    const u32 w = t >> 5, lane = t & 31;
    const u32 a3 = brev_bits(lane, 5);
    const u32 oblk = brev_bits(blk, (int)s0);
#pragma unroll
    for (int g = 0; g < (int)Geo::G; g++) {
        const u32 k2 = w * Geo::G + g;
        const u32 b3 = brev_bits(k2, 5);
#pragma unroll
        for (int c = 0; c < (int)Geo::RS; c++) {
            const u32 kl = (brev_bits((u32)c, R) << 10) | (k2 << 5) | lane;
            x[g * Geo::RS + c] = in[((u64)kl << s0) + oblk];
        }
        if (R > 0) {
            u32 tb[R > 0 ? R : 1];
#pragma unroll
            for (int u = 1; u <= R; u++) tb[u - 1] = pass3_base<R>(blk, (u32)g, u, t);
            gs_levels<R, 1>(x + g * Geo::RS, itwc, tb, q, q2, Geo::T);
        }
#pragma unroll
        for (int c = 0; c < (int)Geo::RS; c++) smem[swz<R>(a3, b3 * Geo::RS + c)] = x[g * Geo::RS + c];
    }
}

template <int R>
TFB_HD void inv_phaseB(u64* x, u64* smem, const tw_t* __restrict__ itw, const u64 q, const u32 t,
                       const u32 s0, const u32 blk) {
    typedef NttGeo<R> Geo;
    const u64 q2 = 2 * q;
    const u32 a2 = t >> R, c2 = t & (Geo::RS - 1);
#pragma unroll
    for (int b = 0; b < 32; b++) x[b] = smem[swz<R>(a2, b * Geo::RS + c2)];
    u32 tb[5];
#pragma unroll
    for (int u = 1; u <= 5; u++) tb[u - 1] = (1u << (s0 + 4 + u)) + (blk << (4 + u)) + (a2 << (u - 1));
    gs_levels<5, 1>(x, itw, tb, q, q2);
#pragma unroll
    for (int b = 0; b < 32; b++) smem[swz<R>(a2, b * Geo::RS + c2)] = x[b];
}

template <int R>
TFB_HD void inv_phaseA(u64* x, u64* __restrict__ out, const u64* smem, const tw_t* __restrict__ itw,
                       const u64 q, const u32 t, const u32 s0, const u32 blk, const tw_t tn, const tw_t twn) {
    typedef NttGeo<R> Geo;
    const u64 q2 = 2 * q;
#pragma unroll
    for (int a = 0; a < 32; a++) x[a] = smem[swz<R>(a, t)];
    u32 tb[5];
#pragma unroll
    for (int s = 1; s <= 5; s++) tb[s - 1] = (1u << (s0 + s - 1)) + (blk << (s - 1));
    if (s0 == 0) {
        gs_levels<5, 2>(x, itw, tb, q, q2);
        // last level with N^-1 folded in (both outputs pass through a Shoup product)
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const u64 U = x[k], V = x[k + 16];
            x[k] = shoup_lazy(U + V, tn.w, tn.wp, q);
            x[k + 16] = shoup_lazy(U - V + q2, twn.w, twn.wp, q);
        }
    } else {
        gs_levels<5, 1>(x, itw, tb, q, q2);
    }
#pragma unroll
    for (int a = 0; a < 32; a++) out[a * Geo::T + t] = csub(x[a], q);
}
