// Persistent row kernels of the third-generation ladder (ntt_core3.cuh), N = 2^(10+R) sub-blocks, R = 2..4.
//   * CTAs of T = N/32 threads walk over their rows with stride gridDim.x; 512/T CTAs are resident per SM
//     (one at N = 2^14, two at 2^13, four at 2^12), so that at the smaller sizes the shared-memory phases of
//     one row overlap the butterflies of another;
//   * the next row arrives as 32 bulk copies (cp.async.bulk, TMA, SASS UBLKCP) issued by the lanes of warp 0
//     as soon as the current row's last shared-memory reads are done, each to its skewed slot;
//   * the per-prime reduction tables (16 entries of 16 bytes) are built once per CTA in shared memory.
// Included by ntt_kernels3.cu (R = 4) and ntt_kernels4.cu (R = 2, 3) so the instantiations compile in parallel.
#pragma once
#include "engine.h"
#include "ntt_core3.cuh"

namespace v3k {

__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64* bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(u64* bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, u32 bytes, u64* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP_V3K:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_V3K;\n\t"
        "bra WAIT_LOOP_V3K;\n\t"
        "DONE_V3K:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// DRAM -> L2 ahead of time: the row a CTA will need NEXT is prefetched while the current one is being transformed, so
// the bulk copy issued after pass 3's loads finds it in L2 (measured with the quotient change: -5 % kernel time,
// tools/ntt_lab.cu ABL bits 128 + 2048)
__device__ __forceinline__ void l2_prefetch(const void* src, u32 bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
// one row of N positions as 32 bulk copies of T words, copy a to its skewed slot
template <int R>
__device__ __forceinline__ void tma_load_row_skewed(u64* smem, const u64* src, u64* bar, const u32 lane) {
    typedef NttGeo<R> Geo;
    if (lane == 0) mbar_expect_tx(bar, Geo::N * 8);
    __syncwarp();
    fence_proxy_async();
    tma_load_1d(smem + v3::slot<R>(lane, 0), src + lane * Geo::T, Geo::T * 8, bar);
}
__device__ __forceinline__ void build_redtab(v3::redent_t* redtab, const PrimeParams* __restrict__ pp, const u32 L, const u32 t,
                                             const u32 nthreads) {
#pragma unroll 1
    for (u32 i = t; i < L * 16; i += nthreads) {
        const u64 q = pp[i >> 4].pc.q;
        redtab[i].c2 = q - (u64)(i & 15) * q;
        redtab[i].c3 = redtab[i].c2 + 4 * q;
    }
}

__device__ __forceinline__ void build_redtab8(u64* redtab8, const PrimeParams* __restrict__ pp, const u32 L, const u32 t, const u32 nthreads) {
#pragma unroll 1
    for (u32 i = t; i < L * 16; i += nthreads) {
        const u64 q = pp[i >> 4].pc.q;
        redtab8[i] = q - (u64)(i & 15) * q;
    }
}

// Where the rows of a forward launch come from when they are not one contiguous buffer: polynomial p = unit / lj of the
// launch takes its first lq rows from the caller's operand buffers (polynomials [0, polys0) from base[0], the rest from
// base[1]) and its remaining lj - lq rows from `ext` ([polys][lj - lq][N]).  This is how tfb_ct_tensor transforms c1 and
// c2 in ONE launch (lq = lj) and how tfb_bfv_mul's joint-basis transform reads the Q rows of the expanded operands
// straight from the caller's ciphertexts (the expansion then writes only its K new rows, rns_fast.cu).  lj = 0: `in` is
// one contiguous buffer.
struct NttSrc {
    const u64* base[2];
    const u64* ext;
    u32 polys0, lq, lj;
};
// Evaluated only while no residues are live (before the first row and at the top of the row loop): the next row's address
// then waits in shared memory until the bulk copy is issued after pass 3's loads, so the gather costs the butterfly
// ladder no registers (computed at the copy's issue point it cost 40-64 bytes of spill stack, ptxas -v).
template <int R>
__device__ __forceinline__ const u64* src_row(const u64* in, const NttSrc& src, const u32 unit, const u32 in_div, const u32 s0, const u64 nrow) {
    typedef NttGeo<R> Geo;
    if (src.lj == 0) return in + (u64)((unit / in_div) >> s0) * nrow + (u64)(unit & ((1u << s0) - 1)) * Geo::N;
    const u32 p = unit / src.lj, i = unit - p * src.lj;
    if (i < src.lq) {
        const bool second = p >= src.polys0;
        return src.base[second] + ((u64)(p - (second ? src.polys0 : 0)) * src.lq + i) * Geo::N;
    }
    return src.ext + ((u64)p * (src.lj - src.lq) + (i - src.lq)) * Geo::N;
}

// CRT keyswitch digits formed inside the transform (DIG = 1, s0 = 0): unit = (b * dn + kk) * L + j is digit k0 + kk of
// ciphertext b under target prime j; its source row is residue row k0 + kk of that ciphertext's last component.
// Base-2^w digits (DIG = 2): `cend` is the limb-major binary form [batch][nl][N] of the integers (ct_stride = nl * N words);
// digit k0 + kk is bits (k0 + kk) w .. of it, so its source row is limb ((k0 + kk) w) / 64 (pass1_pow2).
struct CrtDig {
    const u64* cend;
    u64 ct_stride;
    const PrimeParams* ppq;   // primes of the ciphertext ring (source); CRT digits only
    u32 k0, dn;
    u32 w, nl;                // base-2^w digits only
};
template <int R, int DIG>
__device__ __forceinline__ const u64* crt_row(const CrtDig& cd, const u32 unit, const u32 L) {
    const u32 d = unit / L;                        // (b, kk)
    const u32 k = cd.k0 + d % cd.dn;
    return cd.cend + (u64)(d / cd.dn) * cd.ct_stride + (u64)(DIG == 2 ? (k * cd.w) >> 6 : k) * NttGeo<R>::N;
}

template <int R, bool S0ZERO, int DIG = 0>   // DIG: 0 rows from memory, 1 CRT digits, 2 base-2^w digits formed on load
__global__ void __launch_bounds__(NttGeo<R>::T, 512 / NttGeo<R>::T)
ntt_fwd_s_kernel(const u64* __restrict__ in, u64* __restrict__ out, const tw_t* __restrict__ tw_all,
                 const PrimeParams* __restrict__ pp, const u32 L, const u32 s0_, const u32 nunits, const u32 in_div,
                 const NttSrc src, const CrtDig cd = CrtDig()) {
    typedef NttGeo<R> Geo;
    extern __shared__ __align__(128) u64 smem[];
    __shared__ __align__(8) u64 bar;
    __shared__ const u64* next_src;
    __shared__ v3::redent_t redtab[TFB_MAX_L * 16];
    const u32 s0 = S0ZERO ? 0 : s0_;
    u32 t = threadIdx.x;
    const u64 nrow = (u64)Geo::N << s0;
    u32 unit = blockIdx.x;
    if (t == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    __shared__ u64 crt_q[DIG == 1 ? TFB_MAX_L : 1];   // source primes of the digits (CRT): read from shared memory, not through a dependent global load per row
    if (DIG == 1 && t < cd.dn) crt_q[t] = cd.ppq[cd.k0 + t].pc.q;
    build_redtab(redtab, pp, L, t, Geo::T);
    __syncthreads();
    // in_div > 1 (s0 = 0 only): input row = unit / in_div -- one small-integer polynomial (a keyswitch digit) is
    // transformed under in_div consecutive primes without being replicated in memory first
    if (t < 32 && unit < nunits) tma_load_row_skewed<R>(smem, DIG ? crt_row<R, DIG>(cd, unit, L) : src_row<R>(in, src, unit, in_div, s0, nrow), &bar, t);
    u32 parity = 0;
    u64 x[32];
    for (; unit < nunits; unit += gridDim.x) {
        const u64 row = unit >> s0;
        const u32 blk = unit & ((1u << s0) - 1);
        const u32 prime = (u32)(row % L);
        const tw_t* tw = tw_all + (u64)prime * nrow;
        const v3::Red3 rp = v3::make_red3(pp[prime].pc.q, pp[prime].sh, redtab + prime * 16);
        // opaque to the optimiser: keeps the per-row address arithmetic inside the loop (hoisted, it
        // is 32 loop-invariant values per thread that ptxas spills to local memory)
        asm volatile("" : "+r"(t));
        {
            const u32 nxt = unit + gridDim.x;
            if (t < 32 && nxt < nunits) {
                const u64* p = DIG ? crt_row<R, DIG>(cd, nxt, L) : src_row<R>(in, src, nxt, in_div, s0, nrow);
                l2_prefetch(p + t * Geo::T, Geo::T * 8);
                if (t == 0) next_src = p;       // read back after pass 3's loads (three barriers later)
            }
        }
        mbar_wait(&bar, parity);
        parity ^= 1;
        if (DIG == 1) v3::pass1_crt<R>(x, smem, tw, rp, t, crt_q[(unit / L) % cd.dn], pp[prime].pc.br_hi);
        else if (DIG == 2) {
            const u32 d = unit / L, bit = (cd.k0 + d % cd.dn) * cd.w, limb = bit >> 6, off = bit & 63;
            const bool two = off + cd.w > 64 && limb + 1 < cd.nl;
            v3::pass1_pow2<R>(x, smem, tw, rp, t, off, (1ull << cd.w) - 1,
                              two ? cd.cend + (u64)(d / cd.dn) * cd.ct_stride + (u64)(limb + 1) * Geo::N : nullptr);
        } else v3::pass1<R>(x, smem, tw, rp, t, s0, blk);     // reads and writes this thread's own slots
        __syncthreads();
        v3::pass2<R>(x, smem, tw, rp, t, s0, blk);
        __syncthreads();
        v3::pass3_load<R>(x, smem, t);
        __syncthreads();
        const u32 next = unit + gridDim.x;
        if (t < 32 && next < nunits) tma_load_row_skewed<R>(smem, next_src, &bar, t);
        v3::pass3_compute_store<R, S0ZERO>(x, out + row * nrow, tw_all + (u64)(L + prime) * nrow, rp, t, s0, blk);
    }
}

// Forward sub-blocks of rows of 2^(10+R+s0) positions, s0 >= 1, with the row's LAST global level applied while loading
// (v3::pass1_cross_global): `in` carries levels 1..s0-1 (canonical; s0 = 1: the untouched input row).  Both CTAs of a
// pair read both sub-blocks (the second read is an L2 hit), so the separate HBM pass of that level -- 27 % of the
// N = 2^15 forward transform -- disappears at the price of one more product per position.  OUT OF PLACE only: a CTA
// writes natural-order outputs all over the row while its partner may still be reading the inputs.
template <int R>
__global__ void __launch_bounds__(NttGeo<R>::T, 512 / NttGeo<R>::T)
ntt_fwd_x_kernel(const u64* __restrict__ in, u64* __restrict__ out, const tw_t* __restrict__ tw_all,
                 const PrimeParams* __restrict__ pp, const u32 L, const u32 s0, const u32 nunits) {
    typedef NttGeo<R> Geo;
    extern __shared__ __align__(128) u64 smem[];
    __shared__ v3::redent_t redtab[TFB_MAX_L * 16];
    u32 t = threadIdx.x;
    const u64 nrow = (u64)Geo::N << s0;
    build_redtab(redtab, pp, L, t, Geo::T);
    __syncthreads();
    u64 x[32];
    for (u32 unit = blockIdx.x; unit < nunits; unit += gridDim.x) {
        const u64 row = unit >> s0;
        const u32 blk = unit & ((1u << s0) - 1);
        const u32 prime = (u32)(row % L);
        const tw_t* tw = tw_all + (u64)prime * nrow;
        const v3::Red3 rp = v3::make_red3(pp[prime].pc.q, pp[prime].sh, redtab + prime * 16);
        asm volatile("" : "+r"(t));   // see ntt_fwd_s_kernel
        const u64* even = in + row * nrow + (u64)(blk & ~1u) * Geo::N;
        {   // the pair this CTA handles next: DRAM -> L2 while this one is transformed
            const u32 nxt = unit + gridDim.x;
            if (t < 64 && nxt < nunits)
                l2_prefetch(in + (u64)(nxt >> s0) * nrow + (u64)((nxt & ((1u << s0) - 1)) & ~1u) * Geo::N + (u64)t * Geo::T, Geo::T * 8);
        }
        v3::pass1_cross_global<R>(x, even, even + Geo::N, blk & 1, tw[(1u << (s0 - 1)) + (blk >> 1)], rp, t);
        v3::pass1_cross_levels(x, tw, rp, s0, blk);
        v3::pass1_store<R>(x, smem, t);
        __syncthreads();
        v3::pass2<R>(x, smem, tw, rp, t, s0, blk);
        __syncthreads();
        v3::pass3_load<R>(x, smem, t);
        __syncthreads();              // the buffer is free for the next unit's pass 1
        v3::pass3_compute_store<R, false>(x, out + row * nrow, tw_all + (u64)(L + prime) * nrow, rp, t, s0, blk);
    }
}

// inverse, s0 == 0 only (for longer rows the natural-order input of a sub-block is strided)
template <int R>
__global__ void __launch_bounds__(NttGeo<R>::T, 512 / NttGeo<R>::T)
ntt_inv_s_kernel(const u64* __restrict__ in, u64* __restrict__ out, const tw_t* __restrict__ tw_all,
                 const PrimeParams* __restrict__ pp, const u32 L, const u32 nunits) {
    typedef NttGeo<R> Geo;
    extern __shared__ __align__(128) u64 smem[];
    __shared__ __align__(8) u64 bar;
    __shared__ u64 redtab8[TFB_MAX_L * 16];
    u32 t = threadIdx.x;
    u32 unit = blockIdx.x;
    if (t == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    build_redtab8(redtab8, pp, L, t, Geo::T);
    __syncthreads();
    if (t == 0 && unit < nunits) {
        mbar_expect_tx(&bar, Geo::N * 8);
        tma_load_1d(smem, in + (u64)unit * Geo::N, Geo::N * 8, &bar);
    }
    u32 parity = 0;
    u64 x[32];
    for (; unit < nunits; unit += gridDim.x) {
        asm volatile("" : "+r"(t));   // see ntt_fwd_s_kernel
        const u32 prime = (u32)(unit % L);
        const tw_t* tw = tw_all + (u64)prime * Geo::N;
        v3::Red3 rp = v3::make_red3(pp[prime].pc.q, pp[prime].sh, nullptr);
        rp.tab8 = redtab8 + prime * 16;
        if (t < 32 && unit + gridDim.x < nunits) l2_prefetch(in + (u64)(unit + gridDim.x) * Geo::N + t * Geo::T, Geo::T * 8);
        mbar_wait(&bar, parity);
        parity ^= 1;
        v3::inv_pass3_load<R>(x, smem, t);
        __syncthreads();  // the flat copy is fully read before it is overwritten in skewed order
        v3::inv_pass3_compute_store<R>(x, smem, tw_all + (u64)(L + prime) * Geo::N, rp, t);
        __syncthreads();
        v3::inv_pass2<R>(x, smem, tw, rp, t);
        __syncthreads();
        v3::inv_pass1_load<R>(x, smem, t);
        __syncthreads();
        const u32 next = unit + gridDim.x;
        if (t == 0 && next < nunits) {
            fence_proxy_async();
            mbar_expect_tx(&bar, Geo::N * 8);
            tma_load_1d(smem, in + (u64)next * Geo::N, Geo::N * 8, &bar);
        }
        v3::inv_pass1_compute_store<R>(x, out + (u64)unit * Geo::N, tw, rp, t, pp[prime].ninv, pp[prime].ninv_w1);
    }
}

// inverse over the sub-blocks of longer rows (s0 > 0): sub-block blk's natural-order input is strided
// (n = (kl << s0) + brev_s0(blk)), so it is read with ordinary loads instead of a bulk copy; the sub-block's 14 levels
// follow, and its canonical results go to positions blk*N .. of the row for the global inverse stages.
template <int R>
__global__ void __launch_bounds__(NttGeo<R>::T, 512 / NttGeo<R>::T)
ntt_inv_sub_kernel(const u64* __restrict__ in, u64* __restrict__ out, const tw_t* __restrict__ tw_all,
                   const PrimeParams* __restrict__ pp, const u32 L, const u32 s0, const u32 nunits) {
    typedef NttGeo<R> Geo;
    extern __shared__ __align__(128) u64 smem[];
    __shared__ u64 redtab8[TFB_MAX_L * 16];
    u32 t = threadIdx.x;
    const u64 nrow = (u64)Geo::N << s0;
    build_redtab8(redtab8, pp, L, t, Geo::T);
    __syncthreads();
    u64 x[32];
    for (u32 unit = blockIdx.x; unit < nunits; unit += gridDim.x) {
        asm volatile("" : "+r"(t));   // see ntt_fwd_s_kernel
        const u64 row = unit >> s0;
        const u32 blk = unit & ((1u << s0) - 1);
        const u32 prime = (u32)(row % L);
        const tw_t* tw = tw_all + (u64)prime * nrow;
        v3::Red3 rp = v3::make_red3(pp[prime].pc.q, pp[prime].sh, nullptr);
        rp.tab8 = redtab8 + prime * 16;
        {
            const u32 w = t >> 5, lane = t & 31;
            const u64* irow = in + row * nrow + brev_bits(blk, (int)s0);
#pragma unroll
            for (int g = 0; g < (int)Geo::G; g++)
#pragma unroll
                for (int c = 0; c < (int)Geo::RS; c++)
                    x[g * Geo::RS + c] = irow[(u64)((brev_bits((u32)c, R) << 10) | ((Geo::G * w + g) << 5) | lane) << s0];
        }
        v3::inv_pass3_compute_store<R>(x, smem, tw_all + (u64)(L + prime) * nrow, rp, t, blk);
        __syncthreads();
        v3::inv_pass2<R>(x, smem, tw, rp, t, s0, blk);
        __syncthreads();
        v3::inv_pass1_load<R>(x, smem, t);
        __syncthreads();              // the buffer is free for the next unit's pass 3
        v3::inv_pass1_levels_all(x, tw, rp, s0, blk);
        u64* orow = out + row * nrow + (u64)blk * Geo::N;
#pragma unroll
        for (int a = 0; a < 32; a++) orow[a * Geo::T + t] = v3::canon3i(x[a], rp);
    }
}

template <int R>
int setup_s() {
    const int smem = (int)v3::Lay<R>::ROW_BYTES;
    TFB_CUDA(cudaFuncSetAttribute(ntt_fwd_s_kernel<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    TFB_CUDA(cudaFuncSetAttribute(ntt_fwd_s_kernel<R, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    TFB_CUDA(cudaFuncSetAttribute(ntt_fwd_s_kernel<R, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    TFB_CUDA(cudaFuncSetAttribute(ntt_inv_s_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    return TFB_OK;
}
// rows of exactly N = 2^(10+R) positions
template <int R>
int launch_s(tfb_ctx* c, const u64* in, u64* out, u64 rows, bool inverse, cudaStream_t st, u32 in_div = 1, const NttSrc* src = nullptr) {
    typedef NttGeo<R> Geo;
    if (rows > 0x7fffffffull) { tfb_set_error("too many rows for one launch"); return TFB_EINVAL; }
    const u64 slots = (u64)(c->num_sms > 0 ? c->num_sms : 148) * (512 / Geo::T);
    const unsigned grid = (unsigned)(rows < slots ? rows : slots);
    if (inverse) {
        ProfScope ps(PC_NTT_INV, st);
        ntt_inv_s_kernel<R><<<grid, Geo::T, v3::Lay<R>::ROW_BYTES, st>>>(in, out, c->d_inv, c->d_pp, c->L, (u32)rows);
    } else {
        ProfScope ps(PC_NTT_FWD, st);
        NttSrc none = {};
        ntt_fwd_s_kernel<R, true><<<grid, Geo::T, v3::Lay<R>::ROW_BYTES, st>>>(in, out, c->d_fwd, c->d_pp, c->L, 0, (u32)rows, in_div, src ? *src : none);
    }
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}

// CRT keyswitch digits k0 .. k0+dn-1 of `cend` ([batch] ciphertexts, ct_stride words apart, residues modulo the primes of
// `c`) in the NTT domain of ring r: dig [batch][dn][r->L][N], rows of exactly N = 2^(10+R) positions
template <int R>
int launch_crt(tfb_ctx* c, tfb_ctx* r, const u64* cend, u64 ct_stride, u64* dig, u32 k0, u32 dn, u64 batch, cudaStream_t st) {
    typedef NttGeo<R> Geo;
    const u64 rows = batch * dn * r->L;
    if (rows > 0x7fffffffull) { tfb_set_error("too many rows for one launch"); return TFB_EINVAL; }
    const u64 slots = (u64)(r->num_sms > 0 ? r->num_sms : 148) * (512 / Geo::T);
    const unsigned grid = (unsigned)(rows < slots ? rows : slots);
    CrtDig cd;
    cd.cend = cend; cd.ct_stride = ct_stride; cd.ppq = c->d_pp; cd.k0 = k0; cd.dn = dn; cd.w = 0; cd.nl = 0;
    ProfScope ps(PC_NTT_FWD, st);
    ntt_fwd_s_kernel<R, true, 1><<<grid, Geo::T, v3::Lay<R>::ROW_BYTES, st>>>(nullptr, dig, r->d_fwd, r->d_pp, r->L, 0, (u32)rows, 1, NttSrc(), cd);
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}

// Base-2^w keyswitch digits k0 .. k0+dn-1 of the integers whose binary limbs are `limbs` ([batch][nl][N], ks_limbs_kernel)
// in the NTT domain of ring r: dig [batch][dn][r->L][N].  The caller guarantees 0 < w < 64 and 2^w <= every prime of r.
template <int R>
int launch_pow2(tfb_ctx* r, const u64* limbs, u32 nl, u32 w, u64* dig, u32 k0, u32 dn, u64 batch, cudaStream_t st) {
    typedef NttGeo<R> Geo;
    const u64 rows = batch * dn * r->L;
    if (rows > 0x7fffffffull) { tfb_set_error("too many rows for one launch"); return TFB_EINVAL; }
    const u64 slots = (u64)(r->num_sms > 0 ? r->num_sms : 148) * (512 / Geo::T);
    const unsigned grid = (unsigned)(rows < slots ? rows : slots);
    CrtDig cd;
    cd.cend = limbs; cd.ct_stride = (u64)nl * Geo::N; cd.ppq = nullptr; cd.k0 = k0; cd.dn = dn; cd.w = w; cd.nl = nl;
    ProfScope ps(PC_NTT_FWD, st);
    ntt_fwd_s_kernel<R, true, 2><<<grid, Geo::T, v3::Lay<R>::ROW_BYTES, st>>>(nullptr, dig, r->d_fwd, r->d_pp, r->L, 0, (u32)rows, 1, NttSrc(), cd);
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}

}  // namespace v3k
