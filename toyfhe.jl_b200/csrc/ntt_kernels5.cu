// Rows longer than one shared-memory buffer: N = 2^15 (the CKKS config of BASELINE.json; 2^16 with one global
// level in front).  A row of 2^15 positions is a PAIR of 2^14 sub-blocks held by a thread-block cluster of two
// CTAs (two SMs): each CTA keeps its sub-block resident exactly like ntt_fwd_s_kernel<4>, and the one butterfly
// level that couples the two sub-blocks reads the partner's shared memory through distributed shared memory
// (mapa + generic loads) instead of taking a separate pass through HBM:
//   forward (pow2_cyc_rings.jl:295-303):  cross level first -- X' = X + wY on rank 0, Y' = X - wY on rank 1,
//            each CTA forming the product wY itself (one extra Shoup product per position instead of a global
//            read-modify-write pass over the row) -- then the 14 in-CTA levels with twiddles of (s0, blk);
//   inverse (pow2_cyc_rings.jl:308-318):  each CTA bulk-loads one contiguous half of the natural-order row, the
//            first pass gathers this sub-block's (interleaved) elements from both halves, 14 in-CTA levels, then
//            the partners exchange their results through shared memory for the last level with N^-1 folded in.
// Cluster barriers (barrier.cluster arrive/wait, split so the wait sits behind independent work) order the
// remote reads against the owner's overwrites; everything else is CTA-local as in ntt_v3_kernels.cuh.
#include "ntt_v3_kernels.cuh"

namespace {
using namespace v3k;
constexpr int R = 4;
typedef NttGeo<R> Geo;

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ u32 cluster_rank() {
    u32 r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// generic address of the same shared-memory location in CTA `rank` of this cluster
__device__ __forceinline__ const u64* map_peer(const u64* p, const u32 rank) {
    u64 out;
    asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"((u64)p), "r"(rank));
    return (const u64*)out;
}

// units = pairs; pair u of row (u >> (s0-1)) is sub-blocks 2m, 2m+1 with m = u mod 2^(s0-1); levels 1..s0-1 of the
// row have already been applied (ntt_fwd_stage_kernel) when s0 > 1
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Geo::T, 1)
ntt_fwd_pair_kernel(const u64* __restrict__ in, u64* __restrict__ out, const tw_t* __restrict__ tw_all,
                    const PrimeParams* __restrict__ pp, const u32 L, const u32 s0, const u32 npairs) {
    extern __shared__ __align__(128) u64 smem[];
    __shared__ __align__(8) u64 bar;
    __shared__ v3::redent_t redtab[TFB_MAX_L * 16];
    u32 t = threadIdx.x;
    const u32 rank = cluster_rank();
    const u64* peer = map_peer(smem, rank ^ 1);
    const u64 nrow = (u64)Geo::N << s0;
    const u32 nclusters = gridDim.x >> 1;
    u32 unit = blockIdx.x >> 1;
    if (t == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    build_redtab(redtab, pp, L, t, Geo::T);
    __syncthreads();
    if (t < 32 && unit < npairs)
        tma_load_row_skewed<R>(smem, in + (u64)(unit >> (s0 - 1)) * nrow + (u64)(2 * (unit & ((1u << (s0 - 1)) - 1)) + rank) * Geo::N, &bar, t);
    u32 parity = 0;
    u64 x[32];
    for (; unit < npairs; unit += nclusters) {
        const u64 row = unit >> (s0 - 1);
        const u32 m = unit & ((1u << (s0 - 1)) - 1);
        const u32 blk = 2 * m + rank;
        const u32 prime = (u32)(row % L);
        const tw_t* tw = tw_all + (u64)prime * nrow;
        const v3::Red3 rp = v3::make_red3(pp[prime].pc.q, pp[prime].sh, redtab + prime * 16);
        asm volatile("" : "+r"(t));   // see ntt_fwd_s_kernel
        mbar_wait(&bar, parity);
        parity ^= 1;
        cluster_arrive();             // my sub-block has landed ...
        cluster_wait();               // ... and so has the partner's
        v3::pass1_cross_load<R>(x, smem, peer, rank, tw[(1u << (s0 - 1)) + m], rp, t);
        cluster_arrive();             // done reading the partner's buffer
        v3::pass1_cross_levels(x, tw, rp, s0, blk);
        cluster_wait();               // the partner is done reading mine: it may be overwritten
        v3::pass1_store<R>(x, smem, t);
        __syncthreads();
        v3::pass2<R>(x, smem, tw, rp, t, s0, blk);
        __syncthreads();
        v3::pass3_load<R>(x, smem, t);
        __syncthreads();
        const u32 next = unit + nclusters;
        if (t < 32 && next < npairs)
            tma_load_row_skewed<R>(smem, in + (u64)(next >> (s0 - 1)) * nrow + (u64)(2 * (next & ((1u << (s0 - 1)) - 1)) + rank) * Geo::N, &bar, t);
        v3::pass3_compute_store<R, false>(x, out + row * nrow, tw_all + (u64)(L + prime) * nrow, rp, t, s0, blk);
    }
}

// rows of exactly 2^15 positions (s0 = 1): unit = row
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Geo::T, 1)
ntt_inv_pair_kernel(const u64* __restrict__ in, u64* __restrict__ out, const tw_t* __restrict__ tw_all,
                    const PrimeParams* __restrict__ pp, const u32 L, const u32 nrows) {
    extern __shared__ __align__(128) u64 smem[];
    __shared__ __align__(8) u64 bar;
    __shared__ u64 redtab8[TFB_MAX_L * 16];
    u32 t = threadIdx.x;
    const u32 rank = cluster_rank();
    const u64* peer = map_peer(smem, rank ^ 1);
    constexpr u64 nrow = (u64)Geo::N * 2;
    const u32 nclusters = gridDim.x >> 1;
    u32 unit = blockIdx.x >> 1;
    if (t == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    build_redtab8(redtab8, pp, L, t, Geo::T);
    __syncthreads();
    if (t == 0 && unit < nrows) {
        mbar_expect_tx(&bar, Geo::N * 8);
        tma_load_1d(smem, in + (u64)unit * nrow + (u64)rank * Geo::N, Geo::N * 8, &bar);
    }
    u32 parity = 0;
    u64 x[32];
    for (; unit < nrows; unit += nclusters) {
        asm volatile("" : "+r"(t));   // see ntt_fwd_s_kernel
        const u32 prime = (u32)(unit % L);
        const tw_t* tw = tw_all + (u64)prime * nrow;
        v3::Red3 rp = v3::make_red3(pp[prime].pc.q, pp[prime].sh, nullptr);
        rp.tab8 = redtab8 + prime * 16;
        mbar_wait(&bar, parity);
        parity ^= 1;
        cluster_arrive();
        cluster_wait();               // both halves of the row are resident
        v3::inv_pass3_load_pair<R>(x, smem, peer, rank, t);
        cluster_arrive();             // done reading the partner's half
        __syncthreads();              // my own half is fully read by this CTA ...
        cluster_wait();               // ... and by the partner: overwrite it in skewed order
        v3::inv_pass3_compute_store<R>(x, smem, tw_all + (u64)(L + prime) * nrow, rp, t, rank);
        __syncthreads();
        v3::inv_pass2<R>(x, smem, tw, rp, t, 1, rank);
        __syncthreads();
        v3::inv_pass1_load<R>(x, smem, t);
        __syncthreads();
        v3::inv_pass1_levels_all(x, tw, rp, 1, rank);
        v3::pass1_store<R>(x, smem, t);        // publish this sub-block's results for the partner
        cluster_arrive();
        cluster_wait();
        v3::inv_cross_combine<R>(x, peer, rank, rp, t);
        cluster_arrive();             // done reading the partner's results
        cluster_wait();               // the partner is done reading mine: the buffer is free for the next row
        const u32 next = unit + nclusters;
        if (t == 0 && next < nrows) {
            fence_proxy_async();
            mbar_expect_tx(&bar, Geo::N * 8);
            tma_load_1d(smem, in + (u64)next * nrow + (u64)rank * Geo::N, Geo::N * 8, &bar);
        }
        v3::inv_cross_finish<R>(x, out + (u64)unit * nrow + (u64)rank * Geo::N, rank ? pp[prime].ninv_w1 : pp[prime].ninv, rp, t);
    }
}

int g_pair_clusters = 0;   // co-resident clusters of two (cudaOccupancyMaxActiveClusters)
}  // namespace

int ntt5_setup_device() {
    const int smem = (int)v3::Lay<R>::ROW_BYTES;
    TFB_CUDA(cudaFuncSetAttribute(ntt_fwd_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    TFB_CUDA(cudaFuncSetAttribute(ntt_inv_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    TFB_CUDA(cudaFuncSetAttribute(v3k::ntt_inv_sub_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * 74, 1, 1);
    cfg.blockDim = dim3(Geo::T, 1, 1);
    cfg.dynamicSmemBytes = (size_t)smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, ntt_fwd_pair_kernel, &cfg) != cudaSuccess || n <= 0) {
        (void)cudaGetLastError();
        n = 0;   // no cluster support: the callers keep the global-stage path
    }
    g_pair_clusters = n;
    return TFB_OK;
}

// rows of 2^(14+s0) positions, s0 >= 1; `in` already carries levels 1..s0-1 (forward).  Returns -1 when the pair
// kernels do not apply (inverse with s0 > 1: the sub-blocks' natural-order inputs interleave four ways).
int launch_ntt_pair(tfb_ctx* c, const u64* in, u64* out, u64 rows, bool inverse, u32 s0, cudaStream_t st) {
    if (!c->v3_ok || g_ntt_force_harvey || g_ntt_max_mode < 2 || g_pair_clusters <= 0 || s0 < 1) return -1;
    if (inverse && s0 != 1) return -1;
    const u64 pairs = rows << (s0 - 1);
    if (pairs > 0x7fffffffull) { tfb_set_error("too many rows for one launch"); return TFB_EINVAL; }
    const unsigned clusters = (unsigned)(pairs < (u64)g_pair_clusters ? pairs : (u64)g_pair_clusters);
    if (inverse) {
        ProfScope ps(PC_NTT_INV, st);
        ntt_inv_pair_kernel<<<2 * clusters, Geo::T, v3::Lay<R>::ROW_BYTES, st>>>(in, out, c->d_inv, c->d_pp, c->L, (u32)pairs);
    } else {
        ProfScope ps(PC_NTT_FWD, st);
        ntt_fwd_pair_kernel<<<2 * clusters, Geo::T, v3::Lay<R>::ROW_BYTES, st>>>(in, out, c->d_fwd, c->d_pp, c->L, s0, (u32)pairs);
    }
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}

// inverse sub-block pass for rows of 2^(14+s0) positions, s0 >= 1 (the global inverse stages follow); -1: not applicable
int launch_ntt_inv_sub(tfb_ctx* c, const u64* in, u64* out, u64 rows, u32 s0, cudaStream_t st) {
    if (!c->v3_ok || g_ntt_force_harvey || g_ntt_max_mode < 2 || s0 < 1) return -1;
    const u64 units = rows << s0;
    if (units > 0x7fffffffull) { tfb_set_error("too many rows for one launch"); return TFB_EINVAL; }
    const u64 nsm = (u64)(c->num_sms > 0 ? c->num_sms : 148);
    const unsigned grid = (unsigned)(units < nsm ? units : nsm);
    ProfScope ps(PC_NTT_INV, st);
    v3k::ntt_inv_sub_kernel<R><<<grid, Geo::T, v3::Lay<R>::ROW_BYTES, st>>>(in, out, c->d_inv, c->d_pp, c->L, s0, (u32)units);
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}
