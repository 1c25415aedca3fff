// Persistent row kernels for N = 2^14 sub-blocks: the 512-thread x 32-residue
// ladder of ntt_core.cuh (levels 5+5+4, two swizzled shared-memory exchanges), but
//   * one CTA per SM walks over its rows with stride gridDim.x (no CTA relaunch
//     gap between rows), and
//   * the next row's 128 KiB are fetched by ONE cp.async.bulk (TMA, UBLKCP) into the
//     row buffer as soon as the current row's last shared-memory reads are done, so
//     the global-load latency and transfer overlap the final pass (4 of 14 levels)
//     and its stores.
// Same arithmetic and index maps as ntt_kernels.cu (pow2_cyc_rings.jl:295-318
// definitions); only the data movement differs.
#include "engine.h"
#include "ntt_core.cuh"
#include "ntt_v3_kernels.cuh"

namespace {

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
        "WAIT_LOOP3:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE3;\n\t"
        "bra WAIT_LOOP3;\n\t"
        "DONE3:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

constexpr int R = 4;
typedef NttGeo<R> Geo;
constexpr u32 ROW_BYTES = Geo::N * sizeof(u64);
// the row is fetched as 4 bulk copies so the first pass can start on partial data?  No:
// level 1 pairs position p with p + N/2, every thread needs the whole row.  One copy.

template <int MODE>
__global__ void __launch_bounds__(Geo::T, 1)
ntt_fwd14p_kernel(const u64* __restrict__ in, u64* __restrict__ out, const tw_t* __restrict__ tw_all,
                  const PrimeParams* __restrict__ pp, const u32 L, const u32 s0, const u32 nunits) {
    extern __shared__ __align__(128) u64 smem[];
    __shared__ __align__(8) u64 bar;
    u32 t = threadIdx.x;
    const u64 nrow = (u64)Geo::N << s0;
    u32 unit = blockIdx.x;
    if (t == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (t == 0 && unit < nunits) {
        mbar_expect_tx(&bar, ROW_BYTES);
        tma_load_1d(smem, in + (u64)(unit >> s0) * nrow + (u64)(unit & ((1u << s0) - 1)) * Geo::N, ROW_BYTES, &bar);
    }
    u32 parity = 0;
    u64 x[32];
    for (; unit < nunits; unit += gridDim.x) {
        const u64 row = unit >> s0;
        const u32 blk = unit & ((1u << s0) - 1);
        const u32 prime = (u32)(row % L);
        const tw_t* tw = tw_all + (u64)prime * nrow;
        const RedParams rp = make_red(pp[prime].pc.q, pp[prime].sh);
        // opaque to the optimiser: keeps the per-row address arithmetic inside the loop (hoisted, it
        // is 32 loop-invariant values per thread that ptxas spills to local memory)
        asm volatile("" : "+r"(t));
        mbar_wait(&bar, parity);
        parity ^= 1;
        // ---- pass 1 (levels 1..5) from the flat natural-order copy
#pragma unroll
        for (int a = 0; a < 32; a++) x[a] = smem[a * Geo::T + t];
        __syncwarp();  // the swizzled slots written below were read by lanes of this warp only
        {
            u32 tb[5];
#pragma unroll
            for (int s = 1; s <= 5; s++) tb[s - 1] = (1u << (s0 + s - 1)) + (blk << (s - 1));
            ct_levels_m<5, MODE, false>(x, tw, tb, rp);
        }
#pragma unroll
        for (int a = 0; a < 32; a++) smem[swz<R>(a, t)] = x[a];
        __syncthreads();
        // ---- pass 2 (levels 6..10)
        fwd_phaseB<R, MODE>(x, smem, tw, rp, t, s0, blk);
        __syncthreads();
        // ---- pass 3 (levels 11..14): read everything, release the buffer, then compute
        const u32 w = t >> 5, lane = t & 31;
        const u32 a3 = brev_bits(lane, 5);
#pragma unroll
        for (int g = 0; g < (int)Geo::G; g++) {
            const u32 b3 = brev_bits(w * Geo::G + g, 5);
#pragma unroll
            for (int c = 0; c < (int)Geo::RS; c++) x[g * Geo::RS + c] = smem[swz<R>(a3, b3 * Geo::RS + c)];
        }
        __syncthreads();
        const u32 next = unit + gridDim.x;
        if (t == 0 && next < nunits) {
            fence_proxy_async();
            mbar_expect_tx(&bar, ROW_BYTES);
            tma_load_1d(smem, in + (u64)(next >> s0) * nrow + (u64)(next & ((1u << s0) - 1)) * Geo::N, ROW_BYTES, &bar);
        }
        u64* orow = out + row * nrow;
        const u32 oblk = brev_bits(blk, (int)s0);
#pragma unroll
        for (int g = 0; g < (int)Geo::G; g++) {
            const u32 k2 = w * Geo::G + g;
            const u32 b3 = brev_bits(k2, 5);
            u32 tb[R];
#pragma unroll
            for (int u = 1; u <= R; u++) tb[u - 1] = pass3_base<R>(blk, (u32)g, u, t);
            ct_levels_m<R, MODE, true>(x + g * Geo::RS, tw_all + (u64)(L + prime) * nrow, tb, rp, Geo::T);
#pragma unroll
            for (int c = 0; c < (int)Geo::RS; c++) {
                const u32 kl = (brev_bits((u32)c, R) << 10) | (k2 << 5) | lane;
                orow[((u64)kl << s0) + oblk] = canon<MODE>(x[g * Geo::RS + c], rp);
            }
        }
    }
}

// inverse, s0 == 0 only (for longer rows the natural-order input of a sub-block is strided)
__global__ void __launch_bounds__(Geo::T, 1)
ntt_inv14p_kernel(const u64* __restrict__ in, u64* __restrict__ out, const tw_t* __restrict__ tw_all,
                  const PrimeParams* __restrict__ pp, const u32 L, const u32 nunits) {
    extern __shared__ __align__(128) u64 smem[];
    __shared__ __align__(8) u64 bar;
    u32 t = threadIdx.x;
    u32 unit = blockIdx.x;
    if (t == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (t == 0 && unit < nunits) {
        mbar_expect_tx(&bar, ROW_BYTES);
        tma_load_1d(smem, in + (u64)unit * Geo::N, ROW_BYTES, &bar);
    }
    u32 parity = 0;
    u64 x[32];
    for (; unit < nunits; unit += gridDim.x) {
        asm volatile("" : "+r"(t));   // see ntt_fwd14p_kernel
        const u32 w = t >> 5, lane = t & 31;
        const u32 a3 = brev_bits(lane, 5);
        const u32 prime = (u32)(unit % L);
        const tw_t* tw = tw_all + (u64)prime * Geo::N;
        const u64 q = pp[prime].pc.q, q2 = 2 * q;
        mbar_wait(&bar, parity);
        parity ^= 1;
        // ---- pass 3 (levels 14..11) from the flat natural-order copy
#pragma unroll
        for (int g = 0; g < (int)Geo::G; g++) {
            const u32 k2 = w * Geo::G + g;
#pragma unroll
            for (int c = 0; c < (int)Geo::RS; c++) {
                const u32 kl = (brev_bits((u32)c, R) << 10) | (k2 << 5) | lane;
                x[g * Geo::RS + c] = smem[kl];
            }
        }
        __syncthreads();  // the flat copy is fully read before it is overwritten in swizzled order
#pragma unroll
        for (int g = 0; g < (int)Geo::G; g++) {
            const u32 b3 = brev_bits(w * Geo::G + g, 5);
            u32 tb[R];
#pragma unroll
            for (int u = 1; u <= R; u++) tb[u - 1] = pass3_base<R>(0, (u32)g, u, t);
            gs_levels<R, 1>(x + g * Geo::RS, tw_all + (u64)(L + prime) * Geo::N, tb, q, q2, Geo::T);
#pragma unroll
            for (int c = 0; c < (int)Geo::RS; c++) smem[swz<R>(a3, b3 * Geo::RS + c)] = x[g * Geo::RS + c];
        }
        __syncthreads();
        // ---- pass 2 (levels 10..6)
        inv_phaseB<R>(x, smem, tw, q, t, 0, 0);
        __syncthreads();
        // ---- pass 1 (levels 5..1, N^-1 folded into the last): read, release the buffer, compute
#pragma unroll
        for (int a = 0; a < 32; a++) x[a] = smem[swz<R>(a, t)];
        __syncthreads();
        const u32 next = unit + gridDim.x;
        if (t == 0 && next < nunits) {
            fence_proxy_async();
            mbar_expect_tx(&bar, ROW_BYTES);
            tma_load_1d(smem, in + (u64)next * Geo::N, ROW_BYTES, &bar);
        }
        {
            u32 tb[5];
#pragma unroll
            for (int s = 1; s <= 5; s++) tb[s - 1] = 1u << (s - 1);
            gs_levels<5, 2>(x, tw, tb, q, q2);
            const tw_t tn = pp[prime].ninv, twn = pp[prime].ninv_w1;   // loaded late: 8 registers less across passes 3 and 2
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const u64 U = x[k], V = x[k + 16];
                x[k] = shoup_lazy(U + V, tn.w, tn.wp, q);
                x[k + 16] = shoup_lazy(U - V + q2, twn.w, twn.wp, q);
            }
        }
        u64* orow = out + (u64)unit * Geo::N;
#pragma unroll
        for (int a = 0; a < 32; a++) orow[a * Geo::T + t] = csub(x[a], q);
    }
}

}  // namespace

int ntt3_setup_device() {
    TFB_CUDA(cudaFuncSetAttribute(ntt_fwd14p_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ROW_BYTES));
    TFB_CUDA(cudaFuncSetAttribute(ntt_fwd14p_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ROW_BYTES));
    TFB_CUDA(cudaFuncSetAttribute(v3k::ntt_fwd_s_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v3::Lay<4>::ROW_BYTES));
    TFB_CUDA(cudaFuncSetAttribute(ntt_inv14p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ROW_BYTES));
    TFB_CUDA(cudaFuncSetAttribute(v3k::ntt_fwd_x_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v3::Lay<4>::ROW_BYTES));
    TFB_CUDA(cudaFuncSetAttribute(v3k::ntt_inv_sub_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v3::Lay<4>::ROW_BYTES));
    int rc = v3k::setup_s<4>();
    if (rc) return rc;
    return ntt4_setup_device();
}

// rows of length 2^(14+s0).  Returns -1 when the persistent kernels do not apply
// (in place: the prefetch of the next row must not race with this row's stores only
// when rows are distinct, which holds in place too -- every row is read completely
// before it is written; inverse with s0 > 0: strided input).
int launch_ntt14p(tfb_ctx* c, const u64* in, u64* out, u64 rows, bool inverse, u32 s0, cudaStream_t st) {
    if (inverse && s0 != 0) return -1;
    const u64 units = rows << s0;
    if (units > 0x7fffffffull) { tfb_set_error("too many rows for one launch"); return TFB_EINVAL; }
    const int nsm = c->num_sms > 0 ? c->num_sms : 148;
    const unsigned grid = (unsigned)(units < (u64)nsm ? units : (u64)nsm);
    if (inverse) {
        ProfScope ps(PC_NTT_INV, st);
        if (c->v3_ok && !g_ntt_force_harvey && g_ntt_max_mode >= 2)
            v3k::ntt_inv_s_kernel<4><<<grid, Geo::T, v3::Lay<4>::ROW_BYTES, st>>>(in, out, c->d_inv, c->d_pp, c->L, (u32)units);
        else
            ntt_inv14p_kernel<<<grid, Geo::T, ROW_BYTES, st>>>(in, out, c->d_inv, c->d_pp, c->L, (u32)units);
    } else {
        ProfScope ps(PC_NTT_FWD, st);
        if (c->v3_ok && !g_ntt_force_harvey && g_ntt_max_mode >= 2) {
            v3k::NttSrc none = {};
            if (s0 == 0) v3k::ntt_fwd_s_kernel<4, true><<<grid, Geo::T, v3::Lay<4>::ROW_BYTES, st>>>(in, out, c->d_fwd, c->d_pp, c->L, 0, (u32)units, 1, none);
            else v3k::ntt_fwd_s_kernel<4, false><<<grid, Geo::T, v3::Lay<4>::ROW_BYTES, st>>>(in, out, c->d_fwd, c->d_pp, c->L, s0, (u32)units, 1, none);
        }
        else if (c->ntt_mode >= 1 && !g_ntt_force_harvey && g_ntt_max_mode >= 1)
            ntt_fwd14p_kernel<1><<<grid, Geo::T, ROW_BYTES, st>>>(in, out, c->d_fwd, c->d_pp, c->L, s0, (u32)units);
        else
            ntt_fwd14p_kernel<0><<<grid, Geo::T, ROW_BYTES, st>>>(in, out, c->d_fwd, c->d_pp, c->L, s0, (u32)units);
    }
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}

// forward NTT of `polys` small-integer polynomials (in [polys][N], values below every prime) under all L primes:
// out [polys][L][N].  Returns -1 when the third-generation kernel does not apply (the caller replicates the rows).
int launch_ntt_bcast(tfb_ctx* c, const u64* in, u64* out, u64 polys, cudaStream_t st) {
    if (!c->v3_ok || g_ntt_force_harvey || g_ntt_max_mode < 2 || g_ntt_version != 3) return -1;
    const u64 rows = polys * c->L;
    if (c->logN == 14) {
        if (rows > 0x7fffffffull) { tfb_set_error("too many rows for one launch"); return TFB_EINVAL; }
        const int nsm = c->num_sms > 0 ? c->num_sms : 148;
        const unsigned grid = (unsigned)(rows < (u64)nsm ? rows : (u64)nsm);
        ProfScope ps(PC_NTT_FWD, st);
        v3k::NttSrc none = {};
        v3k::ntt_fwd_s_kernel<4, true><<<grid, Geo::T, v3::Lay<4>::ROW_BYTES, st>>>(in, out, c->d_fwd, c->d_pp, c->L, 0, (u32)rows, c->L, none);
        TFB_CUDA(cudaGetLastError());
        return TFB_OK;
    }
    return launch_ntt_s_bcast(c, in, out, polys, st);
}

// inverse sub-block pass for rows of 2^(14+s0) positions, s0 >= 1 (the global inverse stages follow); -1: not applicable
int launch_ntt_inv_sub(tfb_ctx* c, const u64* in, u64* out, u64 rows, u32 s0, cudaStream_t st) {
    if (!c->v3_ok || g_ntt_force_harvey || g_ntt_max_mode < 2 || s0 < 1) return -1;
    const u64 units = rows << s0;
    if (units > 0x7fffffffull) { tfb_set_error("too many rows for one launch"); return TFB_EINVAL; }
    const u64 nsm = (u64)(c->num_sms > 0 ? c->num_sms : 148);
    const unsigned grid = (unsigned)(units < nsm ? units : nsm);
    ProfScope ps(PC_NTT_INV, st);
    v3k::ntt_inv_sub_kernel<4><<<grid, Geo::T, v3::Lay<4>::ROW_BYTES, st>>>(in, out, c->d_inv, c->d_pp, c->L, s0, (u32)units);
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}

// Forward transform of rows gathered from up to three buffers (v3k::NttSrc) into one contiguous output, N = 2^12..2^14 on
// third-generation primes.  Returns -1 when that kernel family does not apply (the caller copies and transforms).
int launch_ntt_gather(tfb_ctx* c, const u64* base0, const u64* base1, u32 polys0, u32 lq, const u64* ext, u64* out, u64 polys, cudaStream_t st) {
    if (!c->v3_ok || g_ntt_force_harvey || g_ntt_max_mode < 2 || g_ntt_version != 3 || c->logN < 12 || c->logN > 14) return -1;
    v3k::NttSrc src;
    src.base[0] = base0; src.base[1] = base1; src.ext = ext; src.polys0 = polys0; src.lq = lq; src.lj = c->L;
    const u64 rows = polys * c->L;
    if (c->logN == 14) {
        if (rows > 0x7fffffffull) { tfb_set_error("too many rows for one launch"); return TFB_EINVAL; }
        const int nsm = c->num_sms > 0 ? c->num_sms : 148;
        const unsigned grid = (unsigned)(rows < (u64)nsm ? rows : (u64)nsm);
        ProfScope ps(PC_NTT_FWD, st);
        v3k::ntt_fwd_s_kernel<4, true><<<grid, Geo::T, v3::Lay<4>::ROW_BYTES, st>>>(nullptr, out, c->d_fwd, c->d_pp, c->L, 0, (u32)rows, 1, src);
        TFB_CUDA(cudaGetLastError());
        return TFB_OK;
    }
    return launch_ntt_s_gather(c, &src, out, rows, st);
}

// forward sub-blocks of rows of 2^(14+s0) positions with the last global level applied on load (`in` carries levels
// 1..s0-1); out of place only.  -1: not applicable (the caller runs that level as a global pass).
int launch_ntt_fwd_cross(tfb_ctx* c, const u64* in, u64* out, u64 rows, u32 s0, cudaStream_t st) {
    if (!c->v3_ok || g_ntt_force_harvey || g_ntt_max_mode < 2 || s0 < 1 || in == out) return -1;
    const u64 units = rows << s0;
    if (units > 0x7fffffffull) { tfb_set_error("too many rows for one launch"); return TFB_EINVAL; }
    const u64 nsm = (u64)(c->num_sms > 0 ? c->num_sms : 148);
    const unsigned grid = (unsigned)(units < nsm ? units : nsm);
    ProfScope ps(PC_NTT_FWD, st);
    v3k::ntt_fwd_x_kernel<4><<<grid, Geo::T, v3::Lay<4>::ROW_BYTES, st>>>(in, out, c->d_fwd, c->d_pp, c->L, s0, (u32)units);
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}

// CRT keyswitch digits formed inside the forward transform's load phase (N = 2^12 .. 2^14; ntt_core3.cuh pass1_crt).
// -1: not applicable (the caller extracts the digit rows and transforms them).
int launch_ntt_crt(tfb_ctx* c, tfb_ctx* r, const u64* cend, u64 ct_stride, u64* dig, u32 k0, u32 dn, u64 batch, cudaStream_t st) {
    if (!r->v3_ok || g_ntt_force_harvey || g_ntt_max_mode < 2 || g_ntt_version != 3 || r->logN < 12 || r->logN > 14 || c->N != r->N) return -1;
    if (k0 + dn > c->L) { tfb_set_error("keyswitch digits: digit range out of bounds"); return TFB_EINVAL; }
    if (r->logN == 14) return v3k::launch_crt<4>(c, r, cend, ct_stride, dig, k0, dn, batch, st);
    return launch_ntt_s_crt(c, r, cend, ct_stride, dig, k0, dn, batch, st);
}

// Base-2^w keyswitch digits formed inside the forward transform's load phase from the binary limbs of the integers
// (N = 2^12 .. 2^14; ntt_core3.cuh pass1_pow2).  -1: not applicable.
int launch_ntt_pow2(tfb_ctx* r, const u64* limbs, u32 nl, u32 w, u64* dig, u32 k0, u32 dn, u64 batch, cudaStream_t st) {
    if (!r->v3_ok || g_ntt_force_harvey || g_ntt_max_mode < 2 || g_ntt_version != 3 || r->logN < 12 || r->logN > 14) return -1;
    if (w == 0 || w > 63) return -1;
    for (u32 i = 0; i < r->L; i++)
        if ((1ull << w) > r->q[i]) return -1;
    if (r->logN == 14) return v3k::launch_pow2<4>(r, limbs, nl, w, dig, k0, dn, batch, st);
    return launch_ntt_s_pow2(r, limbs, nl, w, dig, k0, dn, batch, st);
}
