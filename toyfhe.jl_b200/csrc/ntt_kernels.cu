// Negacyclic NTT kernels for sm_100a: the engine's replacement for
// NTT.nntt / NTT.inntt (pow2_cyc_rings.jl:295-318) and their per-prime RNS
// dispatch (crt.jl:247-267).  Rows are laid out [..][L][N] (residue-major, the
// StructArray layout); row r belongs to prime r % L.
//
//  * N = 2^10 .. 2^14 : one CTA per row, whole row resident on chip
//                        (32 residues per thread in registers, two swizzled
//                        shared-memory exchanges) -- ntt_core.cuh.
//  * N = 2^15, 2^16   : the top s0 = logN-14 levels run as global-memory stage
//                        kernels, the rest as 2^s0 row-resident sub-blocks.
//  * N < 2^10         : small shared-memory kernel (test-sized rings of the
//                        reference's own unit tests: N = 16, 32, ...).
#include <atomic>

#include "engine.h"
#include "ntt_core.cuh"

int g_ntt_version = 3;  // 1 = one CTA per row (512x32, ntt_core.cuh), 3 = persistent 512x32 + TMA (ntt_core3.cuh); the 1024x16 generation and the
                        // cluster-pair kernels of round 1 were measured slower and removed (profiles/r01_ntt_sizes*.txt keep their numbers)
bool g_ntt_force_harvey = false;
int g_ntt_max_mode = 2;
bool g_ntt_cross = true;   // N > 2^14 forward, out of place: last global level applied on load (testing hook tfb_debug_ntt_cross)
static std::atomic<unsigned long long> g_launches{0};
unsigned long long tfb_launch_count() { return g_launches.load(); }
void tfb_count_launch(int n) { g_launches.fetch_add((unsigned long long)n); }

// ------------------------------------------------------------ row-resident
template <int R, int MODE>
__global__ void __launch_bounds__(NttGeo<R>::T, 1)
ntt_fwd_row_kernel(const u64* __restrict__ in, u64* __restrict__ out, const tw_t* __restrict__ tw_all,
                   const PrimeParams* __restrict__ pp, const u32 L, const u32 s0) {
    typedef NttGeo<R> Geo;
    extern __shared__ __align__(16) u64 smem[];
    const u32 t = threadIdx.x;
    const u64 row = (u64)blockIdx.x >> s0;
    const u32 blk = blockIdx.x & ((1u << s0) - 1);
    const u32 prime = (u32)(row % L);
    const u64 nrow = (u64)Geo::N << s0;
    const tw_t* tw = tw_all + (u64)prime * nrow;
    const RedParams rp = make_red(pp[prime].pc.q, pp[prime].sh);
    u64 x[32];
    fwd_phaseA<R, MODE>(x, in + row * nrow + (u64)blk * Geo::N, smem, tw, rp, t, s0, blk);
    __syncthreads();
    fwd_phaseB<R, MODE>(x, smem, tw, rp, t, s0, blk);
    __syncthreads();
    fwd_phaseC<R, MODE>(x, out + row * nrow, smem, tw_all + (u64)(L + prime) * nrow, rp, t, s0, blk);
}

template <int R>
__global__ void __launch_bounds__(NttGeo<R>::T, 1)
ntt_inv_row_kernel(const u64* __restrict__ in, u64* __restrict__ out, const tw_t* __restrict__ tw_all,
                   const PrimeParams* __restrict__ pp, const u32 L, const u32 s0) {
    typedef NttGeo<R> Geo;
    extern __shared__ __align__(16) u64 smem[];
    const u32 t = threadIdx.x;
    const u64 row = (u64)blockIdx.x >> s0;
    const u32 blk = blockIdx.x & ((1u << s0) - 1);
    const u32 prime = (u32)(row % L);
    const u64 nrow = (u64)Geo::N << s0;
    const tw_t* tw = tw_all + (u64)prime * nrow;
    const PrimeParams P = pp[prime];
    const u64 q = P.pc.q;
    u64 x[32];
    inv_phaseC<R>(x, in + row * nrow, smem, tw_all + (u64)(L + prime) * nrow, q, t, s0, blk);
    __syncthreads();
    inv_phaseB<R>(x, smem, tw, q, t, s0, blk);
    __syncthreads();
    inv_phaseA<R>(x, out + row * nrow + (u64)blk * Geo::N, smem, tw, q, t, s0, blk, P.ninv, P.ninv_w1);
}

// ------------------------------------------------- global stages (N > 2^14)
// forward level s (1-based) over whole rows, canonical in -> canonical out.  Each thread handles TWO adjacent
// butterflies of one block (half >= 2^14 here), so every access is 128 bits wide.
__global__ void ntt_fwd_stage_kernel(const u64* __restrict__ in, u64* __restrict__ out,
                                     const tw_t* __restrict__ tw_all, const PrimeParams* __restrict__ pp,
                                     const u32 L, const u32 logN, const u32 s, const u64 total) {
    const u64 gid = ((u64)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (gid >= total) return;
    const u64 row = gid >> (logN - 1);
    const u32 i = (u32)(gid & ((1ull << (logN - 1)) - 1));
    const u32 prime = (u32)(row % L);
    const u64 q = pp[prime].pc.q;
    const u32 half = 1u << (logN - s);
    const u32 j = i >> (logN - s), k = i & (half - 1);
    const u64 p0 = (row << logN) + ((u64)j << (logN - s + 1)) + k;
    const tw_t w = tw_all[((u64)prime << logN) + (1u << (s - 1)) + j];
    const ulonglong2 X = *reinterpret_cast<const ulonglong2*>(in + p0), Y = *reinterpret_cast<const ulonglong2*>(in + p0 + half);
    const u64 T0 = shoup_full(Y.x, w.w, w.wp, q), T1 = shoup_full(Y.y, w.w, w.wp, q);
    *reinterpret_cast<ulonglong2*>(out + p0) = make_ulonglong2(add_mod(X.x, T0, q), add_mod(X.y, T1, q));
    *reinterpret_cast<ulonglong2*>(out + p0 + half) = make_ulonglong2(sub_mod(X.x, T0, q), sub_mod(X.y, T1, q));
}
// inverse level s; level 1 folds in N^-1
__global__ void ntt_inv_stage_kernel(const u64* __restrict__ in, u64* __restrict__ out,
                                     const tw_t* __restrict__ tw_all, const PrimeParams* __restrict__ pp,
                                     const u32 L, const u32 logN, const u32 s, const u64 total) {
    const u64 gid = ((u64)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (gid >= total) return;
    const u64 row = gid >> (logN - 1);
    const u32 i = (u32)(gid & ((1ull << (logN - 1)) - 1));
    const u32 prime = (u32)(row % L);
    const PrimeParams P = pp[prime];
    const u64 q = P.pc.q;
    const u32 half = 1u << (logN - s);
    const u32 j = i >> (logN - s), k = i & (half - 1);
    const u64 p0 = (row << logN) + ((u64)j << (logN - s + 1)) + k;
    const tw_t w = tw_all[((u64)prime << logN) + (1u << (s - 1)) + j];
    const ulonglong2 X = *reinterpret_cast<const ulonglong2*>(in + p0), Y = *reinterpret_cast<const ulonglong2*>(in + p0 + half);
    u64 S0 = add_mod(X.x, Y.x, q), S1 = add_mod(X.y, Y.y, q), D0 = sub_mod(X.x, Y.x, q), D1 = sub_mod(X.y, Y.y, q);
    if (s == 1) {
        S0 = shoup_full(S0, P.ninv.w, P.ninv.wp, q);
        S1 = shoup_full(S1, P.ninv.w, P.ninv.wp, q);
        D0 = shoup_full(D0, P.ninv_w1.w, P.ninv_w1.wp, q);
        D1 = shoup_full(D1, P.ninv_w1.w, P.ninv_w1.wp, q);
    } else {
        D0 = shoup_full(D0, w.w, w.wp, q);
        D1 = shoup_full(D1, w.w, w.wp, q);
    }
    *reinterpret_cast<ulonglong2*>(out + p0) = make_ulonglong2(S0, S1);
    *reinterpret_cast<ulonglong2*>(out + p0 + half) = make_ulonglong2(D0, D1);
}

// ------------------------------------------------------ small rows (N < 2^10)
__global__ void ntt_small_kernel(const u64* __restrict__ in, u64* __restrict__ out,
                                 const tw_t* __restrict__ tw_all, const PrimeParams* __restrict__ pp,
                                 const u32 L, const u32 logN, const int inverse) {
    extern __shared__ __align__(16) u64 smem[];
    const u32 N = 1u << logN;
    const u64 row = blockIdx.x;
    const u32 prime = (u32)(row % L);
    const PrimeParams P = pp[prime];
    const u64 q = P.pc.q, q2 = 2 * q;
    const tw_t* tw = tw_all + ((u64)prime << logN);
    const u64* src = in + (row << logN);
    u64* dst = out + (row << logN);
    if (!inverse) {
        for (u32 i = threadIdx.x; i < N; i += blockDim.x) smem[i] = src[i];
        __syncthreads();
        for (u32 s = 1; s <= logN; s++) {
            const u32 half = 1u << (logN - s);
            for (u32 i = threadIdx.x; i < N / 2; i += blockDim.x) {
                const u32 j = i >> (logN - s), k = i & (half - 1);
                const u32 p0 = (j << (logN - s + 1)) + k;
                ct_bfly(smem[p0], smem[p0 + half], tw[(1u << (s - 1)) + j], q, q2);
            }
            __syncthreads();
        }
        for (u32 i = threadIdx.x; i < N; i += blockDim.x)
            dst[brev_bits(i, (int)logN)] = csub(csub(smem[i], q2), q);
    } else {
        for (u32 i = threadIdx.x; i < N; i += blockDim.x) smem[i] = src[brev_bits(i, (int)logN)];
        __syncthreads();
        for (u32 s = logN; s >= 1; s--) {
            const u32 half = 1u << (logN - s);
            for (u32 i = threadIdx.x; i < N / 2; i += blockDim.x) {
                const u32 j = i >> (logN - s), k = i & (half - 1);
                const u32 p0 = (j << (logN - s + 1)) + k;
                gs_bfly(smem[p0], smem[p0 + half], tw[(1u << (s - 1)) + j], q, q2);
            }
            __syncthreads();
        }
        for (u32 i = threadIdx.x; i < N; i += blockDim.x)
            dst[i] = shoup_full(smem[i], P.ninv.w, P.ninv.wp, q);
    }
}

// ------------------------------------------------------------------ launcher
template <int R>
static int launch_row(tfb_ctx* c, const u64* in, u64* out, u64 rows, bool inverse, u32 s0, cudaStream_t st) {
    typedef NttGeo<R> Geo;
    const size_t smem = (size_t)Geo::N * sizeof(u64);
    const u64 blocks = rows << s0;
    if (blocks > 0x7fffffffull) { tfb_set_error("too many rows for one launch"); return TFB_EINVAL; }
    if (inverse)
        { ProfScope ps(PC_NTT_INV, st); ntt_inv_row_kernel<R><<<(unsigned)blocks, Geo::T, smem, st>>>(in, out, c->d_inv, c->d_pp, c->L, s0); }
    else if (c->ntt_mode >= 1 && !g_ntt_force_harvey && g_ntt_max_mode >= 1)
        { ProfScope ps(PC_NTT_FWD, st); ntt_fwd_row_kernel<R, 1><<<(unsigned)blocks, Geo::T, smem, st>>>(in, out, c->d_fwd, c->d_pp, c->L, s0); }
    else
        { ProfScope ps(PC_NTT_FWD, st); ntt_fwd_row_kernel<R, 0><<<(unsigned)blocks, Geo::T, smem, st>>>(in, out, c->d_fwd, c->d_pp, c->L, s0); }
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}

template <int R>
static int setup_row() {
    const int smem = (int)(NttGeo<R>::N * sizeof(u64));
    TFB_CUDA(cudaFuncSetAttribute(ntt_inv_row_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    TFB_CUDA(cudaFuncSetAttribute(ntt_fwd_row_kernel<R, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    TFB_CUDA(cudaFuncSetAttribute(ntt_fwd_row_kernel<R, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    return TFB_OK;
}
// opt in to >48 KiB dynamic shared memory on the current device (called per context)
int ntt_setup_device() {
    int rc;
    if ((rc = setup_row<0>())) return rc;
    if ((rc = setup_row<1>())) return rc;
    if ((rc = setup_row<2>())) return rc;
    if ((rc = setup_row<3>())) return rc;
    return setup_row<4>();
}

static int launch_row_dispatch(tfb_ctx* c, int R, const u64* in, u64* out, u64 rows, bool inverse, u32 s0, cudaStream_t st) {
    switch (R) {
        case 0: return launch_row<0>(c, in, out, rows, inverse, s0, st);
        case 1: return launch_row<1>(c, in, out, rows, inverse, s0, st);
        case 2: return launch_row<2>(c, in, out, rows, inverse, s0, st);
        case 3: return launch_row<3>(c, in, out, rows, inverse, s0, st);
        case 4: return launch_row<4>(c, in, out, rows, inverse, s0, st);
    }
    tfb_set_error("internal: bad R");
    return TFB_EINVAL;
}

int launch_ntt(tfb_ctx* c, const u64* in, u64* out, u64 rows, bool inverse, cudaStream_t st) {
    if (rows == 0) return TFB_OK;
    const u32 logN = c->logN;
    if (logN < 10) {
        const u32 N = c->N;
        unsigned threads = N / 2 < 32 ? 32 : (N / 2 > 512 ? 512 : N / 2);
        if (rows > 0x7fffffffull) { tfb_set_error("too many rows for one launch"); return TFB_EINVAL; }
        { ProfScope ps(PC_NTT_OTHER, st); ntt_small_kernel<<<(unsigned)rows, threads, (size_t)N * sizeof(u64), st>>>(
            in, out, inverse ? c->d_inv : c->d_fwd, c->d_pp, c->L, logN, inverse ? 1 : 0); }
        TFB_CUDA(cudaGetLastError());
        return TFB_OK;
    }
    if (logN == 14 && g_ntt_version == 3) return launch_ntt14p(c, in, out, rows, inverse, 0, st);
    if ((logN == 12 || logN == 13) && g_ntt_version == 3) {
        const int rc = launch_ntt_s(c, in, out, rows, inverse, st);
        if (rc != -1) return rc;
    }
    if (logN <= 14) return launch_row_dispatch(c, (int)logN - 10, in, out, rows, inverse, 0, st);
    if (logN > 16) { tfb_set_error("N > 2^16 is not supported"); return TFB_EUNSUPPORTED; }
    // long rows: s0 global levels + row-resident sub-blocks, through scratch
    const u32 s0 = logN - 14;
    const size_t bytes = (size_t)rows * c->N * sizeof(u64);
    int rc = ws_reserve(c, bytes);
    if (rc) return rc;
    u64* tmp = (u64*)c->ws;
    const u64 total = rows << (logN - 1);
    const unsigned tb = 256;
    const unsigned nb = (unsigned)((total / 2 + tb - 1) / tb);   // two butterflies per thread
    if (!inverse) {
        if (g_ntt_version == 3 && in != out && g_ntt_cross) {
            // levels 1..s0-1 as global passes, level s0 inside the sub-block kernel's loads (one HBM pass less)
            const u64* src = in;
            for (u32 s = 1; s + 1 <= s0; s++) {
                { ProfScope ps(PC_NTT_OTHER, st); ntt_fwd_stage_kernel<<<nb, tb, 0, st>>>(src, tmp, c->d_fwd, c->d_pp, c->L, logN, s, total); }
                src = tmp;
            }
            TFB_CUDA(cudaGetLastError());
            rc = launch_ntt_fwd_cross(c, src, out, rows, s0, st);
            if (rc != -1) return rc;
            if (s0 > 1) {   // not applicable after all: finish the remaining global level, then the plain sub-blocks
                { ProfScope ps(PC_NTT_OTHER, st); ntt_fwd_stage_kernel<<<nb, tb, 0, st>>>(tmp, tmp, c->d_fwd, c->d_pp, c->L, logN, s0, total); }
                TFB_CUDA(cudaGetLastError());
                return launch_ntt14p(c, tmp, out, rows, false, s0, st);
            }
        }
        const u64* src = in;
        for (u32 s = 1; s <= s0; s++) {
            { ProfScope ps(PC_NTT_OTHER, st); ntt_fwd_stage_kernel<<<nb, tb, 0, st>>>(src, tmp, c->d_fwd, c->d_pp, c->L, logN, s, total); }
            src = tmp;
        }
        TFB_CUDA(cudaGetLastError());
        if (g_ntt_version == 3) return launch_ntt14p(c, tmp, out, rows, false, s0, st);
        return launch_row_dispatch(c, 4, tmp, out, rows, false, s0, st);
    }
    rc = g_ntt_version == 3 ? launch_ntt_inv_sub(c, in, tmp, rows, s0, st) : -1;
    if (rc == -1) rc = launch_row_dispatch(c, 4, in, tmp, rows, true, s0, st);
    if (rc) return rc;
    for (u32 s = s0; s >= 1; s--) {
        { ProfScope ps(PC_NTT_OTHER, st); ntt_inv_stage_kernel<<<nb, tb, 0, st>>>(tmp, s == 1 ? out : tmp, c->d_inv, c->d_pp, c->L, logN, s, total); }
    }
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}

// ------------------------------------------------- CRT keyswitch digits fused with the first global level (N = 2^15)
// rlwe_she.jl:326-329: digit k of a ciphertext component is the centred residue modulo q_k, re-embedded in every prime
// of the target ring.  For rows of 2^15 positions the forward transform starts with one global level anyway; this
// kernel forms the two embedded operands of each level-1 butterfly directly from the ciphertext row (read once per
// target prime, from L2) and writes the level's result -- the digit rows are never written out and read back.
// out [B][Dn][Lt][N]: level 1 applied; the 2^14 sub-block kernels finish the transform.
__global__ void ks_crt_stage1_kernel(const u64* __restrict__ cend, const u64 ct_stride, u64* __restrict__ out,
                                     const tw_t* __restrict__ tw_all, const PrimeParams* __restrict__ ppq,
                                     const PrimeParams* __restrict__ ppt, const u32 Lt, const u32 logN, const u32 k0, const u32 Dn,
                                     const u64 total) {
    const u64 gid = ((u64)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (gid >= total) return;
    const u32 half = 1u << (logN - 1);
    const u32 i = (u32)(gid & (half - 1));
    const u64 r = gid >> (logN - 1);          // (b, kk, j)
    const u32 j = (u32)(r % Lt);
    const u64 bk = r / Lt;
    const u64 b = bk / Dn;
    const u32 k = k0 + (u32)(bk % Dn);
    const u64 qk = ppq[k].pc.q;
    const PrimeConst pc = ppt[j].pc;
    const u64* src = cend + b * ct_stride + ((u64)k << logN) + i;
    const ulonglong2 c0 = *reinterpret_cast<const ulonglong2*>(src), c1 = *reinterpret_cast<const ulonglong2*>(src + half);
    auto embed = [&](const u64 c) {
        const bool neg = c > (qk >> 1);
        const u64 mag = neg ? qk - c : c;
        const u64 v = (qk >> 1) < pc.q ? mag : barrett_red64(mag, pc);   // |digit| <= q_k / 2: no reduction unless the target prime is smaller (uniform branch)
        return neg ? neg_mod(v, pc.q) : v;
    };
    const tw_t w = tw_all[((u64)j << logN) + 1];
    const u64 X0 = embed(c0.x), X1 = embed(c0.y);
    const u64 T0 = shoup_full(embed(c1.x), w.w, w.wp, pc.q), T1 = shoup_full(embed(c1.y), w.w, w.wp, pc.q);
    u64* dst = out + (r << logN) + i;
    *reinterpret_cast<ulonglong2*>(dst) = make_ulonglong2(add_mod(X0, T0, pc.q), add_mod(X1, T1, pc.q));
    *reinterpret_cast<ulonglong2*>(dst + half) = make_ulonglong2(sub_mod(X0, T0, pc.q), sub_mod(X1, T1, pc.q));
}
// CRT digit polynomials k0..k0+Dn-1 of `cend` in the NTT domain of ring r (N = 2^15): dig [B][Dn][r->L][N].
// Returns -1 when the fused route does not apply (the caller extracts the digits and transforms them separately).
int launch_ks_crt_ntt(tfb_ctx* c, tfb_ctx* r, const u64* cend, u64 ct_stride, u64* dig, u32 k0, u32 Dn, u64 batch, cudaStream_t st) {
    if (r->logN != 15 || c->N != r->N || g_ntt_version != 3 || !r->v3_ok || g_ntt_force_harvey || g_ntt_max_mode < 2) return -1;
    if (k0 + Dn > c->L) { tfb_set_error("keyswitch digits: digit range out of bounds"); return TFB_EINVAL; }
    const u64 rows = batch * Dn * r->L;
    int rc = ws_reserve(r, (size_t)rows * r->N * sizeof(u64));
    if (rc) return rc;
    u64* tmp = (u64*)r->ws;
    const u64 total = rows << (r->logN - 1);
    const unsigned tb = 256;
    { ProfScope ps(PC_KS_DIGITS, st); ks_crt_stage1_kernel<<<(unsigned)((total / 2 + tb - 1) / tb), tb, 0, st>>>(cend, ct_stride, tmp, r->d_fwd, c->d_pp, r->d_pp, r->L, r->logN, k0, Dn, total); }
    TFB_CUDA(cudaGetLastError());
    return launch_ntt14p(r, tmp, dig, rows, false, 1, st);
}
