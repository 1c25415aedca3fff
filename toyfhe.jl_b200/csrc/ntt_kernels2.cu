// Persistent, TMA-prefetched row kernels for N = 2^14 sub-blocks (ntt_core2.cuh).
// One CTA of 1024 threads per SM; each CTA walks over its units (row, sub-block)
// with stride gridDim.x.  The next unit's 128 KiB are fetched by one
// cp.async.bulk (UBLKCP) into the row buffer as soon as the current unit's last
// shared-memory reads are done, and land while the final pass computes/stores.
#include "engine.h"
#include "ntt_core2.cuh"

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
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

constexpr u32 ROW_BYTES = v2::N * sizeof(u64);

template <int MODE>
__global__ void __launch_bounds__(1024, 1)
ntt_fwd14_kernel(const u64* __restrict__ in, u64* __restrict__ out, const tw_t* __restrict__ tw_all,
                 const PrimeParams* __restrict__ pp, const u32 L, const u32 s0, const u32 nunits) {
    using namespace v2;
    extern __shared__ __align__(128) u64 smem[];
    __shared__ __align__(8) u64 bar;
    const u32 t = threadIdx.x;
    const u64 nrow = (u64)N << s0;
    u32 unit = blockIdx.x;
    if (t == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (t == 0 && unit < nunits) {
        mbar_expect_tx(&bar, ROW_BYTES);
        tma_load_1d(smem, in + (u64)(unit >> s0) * nrow + (u64)(unit & ((1u << s0) - 1)) * N, ROW_BYTES, &bar);
    }
    u32 parity = 0;
    u64 x[16];
    for (; unit < nunits; unit += gridDim.x) {
        const u64 row = unit >> s0;
        const u32 blk = unit & ((1u << s0) - 1);
        const u32 prime = (u32)(row % L);
        const tw_t* tw = tw_all + (u64)prime * nrow;
        const RedParams rp = make_red(pp[prime].pc.q, pp[prime].sh);
        mbar_wait(&bar, parity);
        parity ^= 1;
#pragma unroll 1
        for (int k = 0; k < 3; k++) {
            const PassCfg c = make_cfg(k, t, s0, blk);
            fwd_mid_load(x, smem, c);
            fwd_mid_compute<MODE>(x, tw, c, rp);
            __syncwarp();  // pass 1 writes the swizzled slots of lanes of the same warp
            fwd_mid_store(x, smem, c);
            __syncthreads();
        }
        fwd_last_load(x, smem, t);
        __syncthreads();  // every read of the row buffer is done: refill it for the next unit
        const u32 next = unit + gridDim.x;
        if (t == 0 && next < nunits) {
            fence_proxy_async();
            mbar_expect_tx(&bar, ROW_BYTES);
            tma_load_1d(smem, in + (u64)(next >> s0) * nrow + (u64)(next & ((1u << s0) - 1)) * N, ROW_BYTES, &bar);
        }
        fwd_last_compute_store<MODE>(x, out + row * nrow, tw, rp, t, s0, blk);
    }
}

// inverse: TMA prefetch only when the row is contiguous in natural order (s0 == 0)
__global__ void __launch_bounds__(1024, 1)
ntt_inv14_kernel(const u64* __restrict__ in, u64* __restrict__ out, const tw_t* __restrict__ tw_all,
                 const PrimeParams* __restrict__ pp, const u32 L, const u32 s0, const u32 nunits) {
    using namespace v2;
    extern __shared__ __align__(128) u64 smem[];
    __shared__ __align__(8) u64 bar;
    const u32 t = threadIdx.x;
    const u64 nrow = (u64)N << s0;
    const bool tma = s0 == 0;
    u32 unit = blockIdx.x;
    if (t == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (tma && t == 0 && unit < nunits) {
        mbar_expect_tx(&bar, ROW_BYTES);
        tma_load_1d(smem, in + (u64)unit * nrow, ROW_BYTES, &bar);
    }
    u32 parity = 0;
    u64 x[16];
    for (; unit < nunits; unit += gridDim.x) {
        const u64 row = unit >> s0;
        const u32 blk = unit & ((1u << s0) - 1);
        const u32 prime = (u32)(row % L);
        const tw_t* tw = tw_all + (u64)prime * nrow;
        const PrimeParams P = pp[prime];
        const u64 q = P.pc.q;
        if (tma) {
            mbar_wait(&bar, parity);
            parity ^= 1;
            inv_first_load(x, smem, t, 0, 0);
            __syncthreads();  // flat natural-order copy fully read before it is overwritten in place
        } else {
            inv_first_load(x, in + row * nrow, t, s0, blk);
        }
        inv_first_compute_store(x, smem, tw, q, t, s0, blk);
        __syncthreads();
#pragma unroll 1
        for (int k = 2; k >= 1; k--) {
            const PassCfg c = make_cfg(k, t, s0, blk);
            inv_mid_load(x, smem, c);
            inv_mid_compute(x, tw, c, q);
            __syncwarp();
            inv_mid_store(x, smem, c);
            __syncthreads();
        }
        const PassCfg c0 = make_cfg(0, t, s0, blk);
        inv_mid_load(x, smem, c0);
        __syncthreads();
        const u32 next = unit + gridDim.x;
        if (tma && t == 0 && next < nunits) {
            fence_proxy_async();
            mbar_expect_tx(&bar, ROW_BYTES);
            tma_load_1d(smem, in + (u64)next * nrow, ROW_BYTES, &bar);
        }
        inv_final_compute_store(x, out + row * nrow + (u64)blk * N, tw, c0, q, t, s0, P.ninv, P.ninv_w1);
    }
}

}  // namespace

int ntt2_setup_device() {
    TFB_CUDA(cudaFuncSetAttribute(ntt_fwd14_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ROW_BYTES));
    TFB_CUDA(cudaFuncSetAttribute(ntt_fwd14_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ROW_BYTES));
    TFB_CUDA(cudaFuncSetAttribute(ntt_inv14_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ROW_BYTES));
    return TFB_OK;
}

// rows of length 2^(14+s0); in/out may alias only when s0 == 0
int launch_ntt14(tfb_ctx* c, const u64* in, u64* out, u64 rows, bool inverse, u32 s0, cudaStream_t st) {
    const u64 units = rows << s0;
    if (units > 0x7fffffffull) { tfb_set_error("too many rows for one launch"); return TFB_EINVAL; }
    int nsm = c->num_sms > 0 ? c->num_sms : 148;
    const unsigned grid = (unsigned)(units < (u64)nsm ? units : (u64)nsm);
    if (inverse) {
        ProfScope ps(PC_NTT_INV, st);
        ntt_inv14_kernel<<<grid, 1024, ROW_BYTES, st>>>(in, out, c->d_inv, c->d_pp, c->L, s0, (u32)units);
    } else {
        ProfScope ps(PC_NTT_FWD, st);
        if (c->ntt_mode >= 1 && !g_ntt_force_harvey && g_ntt_max_mode >= 1)
            ntt_fwd14_kernel<1><<<grid, 1024, ROW_BYTES, st>>>(in, out, c->d_fwd, c->d_pp, c->L, s0, (u32)units);
        else
            ntt_fwd14_kernel<0><<<grid, 1024, ROW_BYTES, st>>>(in, out, c->d_fwd, c->d_pp, c->L, s0, (u32)units);
    }
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}
