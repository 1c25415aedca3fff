// Feasibility probe for a "half-row" forward kernel: a 2^14 row as TWO independent CTAs of 512 threads x 16 residues
// (64 registers per thread, 64 KiB of shared memory each, two CTAs resident per SM = 32 warps instead of 16), the row's
// first level applied while loading (both halves read by both CTAs), then 1 + 4 + 4 + 4 levels with one warp-shuffle
// level and two shared-memory exchanges.  The index maps here are NOT a transform (arbitrary twiddles, no bit reversal):
// the probe has the instruction mix, memory traffic and synchronisation of the real thing and answers one question --
// what rate does this geometry reach -- before the index maps are written.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../toyfhe.jl_b200/csrc/ntt_v3_kernels.cuh"
#include "../toyfhe.jl_b200/csrc/tables.h"

void tfb_set_error(const std::string&) {}
int tfb_cuda_fail(cudaError_t e, const char* what) { fprintf(stderr, "CUDA error %s in %s\n", cudaGetErrorString(e), what); return -1; }
ProfScope::ProfScope(int c, cudaStream_t s) : cls(c), st(s), stop(nullptr) {}
ProfScope::~ProfScope() {}

#ifndef NCTA
#define NCTA 2
#endif
#ifndef VARIANT
#define VARIANT 0   // 1: no shuffle level; 2: contiguous (non-interleaved) stores
#endif

namespace lab {
using namespace v3;
constexpr u32 N = 1u << 14, H = N / 2, T = 512;

__device__ __forceinline__ u64 sl4(u64 x, u64 w, u64 wp, u64 q, u32 ne, u32 shb) {
    u32 x0, x1, w0, w1, p0, p1, u0, u1, m1, c, t0, t1, h0, h1, lo, hi;
    u64 u, t, acc;
    asm("mov.b64 {%0,%1}, %2;" : "=r"(x0), "=r"(x1) : "l"(x));
    asm("mov.b64 {%0,%1}, %2;" : "=r"(w0), "=r"(w1) : "l"(w));
    asm("mov.b64 {%0,%1}, %2;" : "=r"(p0), "=r"(p1) : "l"(wp));
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(u) : "r"(x1), "r"(p0));
    asm("mov.b64 {%0,%1}, %2;" : "=r"(u0), "=r"(u1) : "l"(u));
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(x1), "r"(p1));
    asm("mov.b64 {%0,%1}, %2;" : "=r"(t0), "=r"(t1) : "l"(t));
    asm("{\n\t.reg .u32 d;\n\tmad.lo.cc.u32 d, %4, %5, %6;\n\tmadc.hi.cc.u32 %0, %4, %5, %7;\n\taddc.u32 %1, 0, 0;\n\t"
        "add.cc.u32 %2, %8, %0;\n\taddc.u32 %3, %9, %1;\n\t}"
        : "=&r"(m1), "=&r"(c), "=&r"(h0), "=&r"(h1) : "r"(x0), "r"(p1), "r"(u0), "r"(u1), "r"(t0), "r"(t1));
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(acc) : "r"(h0), "r"(ne));
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(x0), "r"(w0));
    asm("mov.b64 {%0,%1}, %2;" : "=r"(lo), "=r"(hi) : "l"(acc));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(x0), "r"(w1));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(x1), "r"(w0));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(h1), "r"(ne));
    hi = hi - h0 - (h0 << shb);
    asm("mov.b64 %0, {%1,%2};" : "=l"(acc) : "r"(lo), "r"(hi));
    return acc;
}
template <bool RED>
__device__ __forceinline__ void bfly(u64& X, u64& Y, const tw_t w, const Red3& rp) {
    const u64 t = sl4(Y, w.w, w.wp, rp.q, rp.ne, rp.shb);
    const u64 x = X;
    if (RED) {
        const redent_t c = rp.tab[top4(x, rp) & 15];
        X = x + c.c2 + t;
        Y = x + c.c3 - t;
    } else {
        X = x + t;
        Y = x - t + rp.q4;
    }
}
template <u32 REDMASK>
__device__ __forceinline__ void levels4(u64* x, const tw_t* __restrict__ tw, const u32 base, const u32 js, const Red3& rp) {
#pragma unroll
    for (int u = 1; u <= 4; u++) {
        const int half = 16 >> u;
#pragma unroll
        for (int j = 0; j < (1 << (u - 1)); j++) {
            const tw_t w = tw[base + ((1u << (u - 1)) - 1 + j) * js];
#pragma unroll
            for (int k = 0; k < half; k++) {
                if ((REDMASK >> (u - 1)) & 1) bfly<true>(x[j * 2 * half + k], x[j * 2 * half + k + half], w, rp);
                else bfly<false>(x[j * 2 * half + k], x[j * 2 * half + k + half], w, rp);
            }
        }
    }
}
__device__ __forceinline__ u64 shfl64(u64 v, int mask) {
    u32 lo, hi;
    asm("mov.b64 {%0,%1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
    lo = __shfl_xor_sync(0xffffffffu, lo, mask);
    hi = __shfl_xor_sync(0xffffffffu, hi, mask);
    u64 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}

__global__ void __launch_bounds__(T, NCTA)
half_kernel(const u64* __restrict__ in, u64* __restrict__ out, const tw_t* __restrict__ tw_all, const PrimeParams* __restrict__ pp,
            const u32 L, const u32 nunits) {
    extern __shared__ __align__(128) u64 smem[];   // 8192 words + padding (18-word groups)
    __shared__ redent_t redtab[16 * 16];
    u32 t = threadIdx.x;
    v3k::build_redtab(redtab, pp, L < 16 ? L : 16, t, T);
    __syncthreads();
    u64 x[16];
    for (u32 unit = blockIdx.x; unit < nunits; unit += gridDim.x) {
        const u32 row = unit >> 1, blk = unit & 1;
        const u32 prime = row % L;
        const tw_t* tw = tw_all + (u64)prime * N;
        const Red3 rp = make_red3(pp[prime].pc.q, pp[prime].sh, redtab + prime * 16);
        asm volatile("" : "+r"(t));
        const u32 lane = t & 31, warp = t >> 5;
        // ---- load + first level of the row (both halves), then 4 levels on the 16 strided positions of this thread
        {
            const u32 c = ((lane >> 4) << 8) | (warp << 4) | (lane & 15);
            const u64* lo = in + (u64)row * N + c;
            const tw_t w1 = tw[1];
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const u64 a = lo[k * T], b = lo[H + k * T];
                const u64 tt = sl4(b, w1.w, w1.wp, rp.q, rp.ne, rp.shb);
                x[k] = blk ? a - tt + rp.q4 : a + tt;
            }
            levels4<0x04>(x, tw, 2 + blk * 15, 1, rp);
#if VARIANT != 1
            // shuffle level: lanes l and l^16 hold X_k and Y_k; each computes 8 of the 16 butterflies
            const bool up = lane & 16;
            const tw_t ws = tw[64 + blk * 16 + warp];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const u64 send = up ? x[k] : x[k + 8];
                const u64 got = shfl64(send, 16);
                u64 X = up ? got : x[k], Y = up ? x[k + 8] : got;
                bfly<false>(X, Y, ws, rp);
                const u64 back = shfl64(up ? X : Y, 16);
                if (up) { x[k] = back; x[k + 8] = Y; } else { x[k] = X; x[k + 8] = back; }
            }
#endif
#pragma unroll
            for (int k = 0; k < 16; k++) smem[(k * 32 + warp * 2 + (lane >> 4)) * 18 + (lane & 15)] = x[k];
        }
        __syncthreads();
        // ---- pass 2: thread (g = t >> 4, c = t & 15) holds the 16 positions g*256 + k*16 + c
        {
            const u32 g = t >> 4, c = t & 15;
            u64* base = smem + (g * 16) * 18 + c;
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = base[k * 18];
            levels4<0x05>(x, tw, 512 + g * 15, 1, rp);
#pragma unroll
            for (int k = 0; k < 16; k++) base[k * 18] = x[k];
        }
        __syncthreads();
        // ---- pass 3: 16 consecutive positions per thread (128-bit loads), per-thread twiddles from a thread-order table
        {
            const u64* base = smem + t * 18;
#pragma unroll
            for (int k = 0; k < 16; k += 2) {
                const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(base + k);
                x[k] = v.x; x[k + 1] = v.y;
            }
            __syncthreads();
            levels4<0x05>(x, tw_all + (u64)(L + prime) * N, blk * 15 * T + t, T, rp);
            u64* orow = out + (u64)row * N;
#pragma unroll
            for (int k = 0; k < 16; k++) {
#if VARIANT == 2
                orow[blk * H + k * T + t] = canon3(x[k], rp);
#else
                orow[((k * T + t) << 1) + blk] = canon3(x[k], rp);
#endif
            }
        }
    }
}
}  // namespace lab

static u64 splitmix(u64& s) { u64 z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }

int main(int argc, char** argv) {
    const u32 N = 1u << 14, L = 8;
    const u32 rows = argc > 1 ? (u32)atoi(argv[1]) : 2048;
    const u64 qs[8] = {1152921504607338497ull, 1152921504608747521ull, 1152921504609239041ull, 1152921504612646913ull,
                       1152921504614023169ull, 1152921504614055937ull, 1152921504615628801ull, 1152921504615694337ull};
    std::vector<tw_t> fwd((size_t)2 * L * N);
    std::vector<PrimeParams> pp(L);
    for (u32 i = 0; i < L; i++) {
        u64 psi;
        h_minimal_primitive_root(qs[i], 2ull * N, &psi);
        HostTables ht;
        build_tables(N, qs[i], psi, ht);
        memcpy(&fwd[(size_t)i * N], ht.fwd.data(), (size_t)N * sizeof(tw_t));
        permute_pass3(ht.fwd.data(), &fwd[(size_t)(L + i) * N], 14);
        pp[i].pc = ht.pc; pp[i].ninv = ht.ninv; pp[i].ninv_w1 = ht.ninv_w1; pp[i].sh = 60; pp[i].pad_ = 0;
    }
    std::vector<u64> h((size_t)rows * N);
    u64 s = 7;
    for (u32 r = 0; r < rows; r++) for (u32 i = 0; i < N; i++) h[(size_t)r * N + i] = splitmix(s) % qs[r % L];
    u64 *din, *dout, *dref; tw_t* dtw; PrimeParams* dpp;
    cudaMalloc(&din, h.size() * 8); cudaMalloc(&dout, h.size() * 8); cudaMalloc(&dref, h.size() * 8);
    cudaMalloc(&dtw, fwd.size() * sizeof(tw_t)); cudaMalloc(&dpp, L * sizeof(PrimeParams));
    cudaMemcpy(din, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dtw, fwd.data(), fwd.size() * sizeof(tw_t), cudaMemcpyHostToDevice);
    cudaMemcpy(dpp, pp.data(), L * sizeof(PrimeParams), cudaMemcpyHostToDevice);
    const int smem_prod = (int)v3::Lay<4>::ROW_BYTES;
    const int smem_half = 512 * 18 * 8;
    cudaFuncSetAttribute(v3k::ntt_fwd_s_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_prod);
    cudaFuncSetAttribute(lab::half_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_half);
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lab::half_kernel, 512, smem_half);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto time = [&](auto launch) {
        for (int i = 0; i < 3; i++) launch();
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        const int reps = 20;
        for (int i = 0; i < reps; i++) launch();
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        return ms / reps;
    };
    const float base = time([&] { v3k::ntt_fwd_s_kernel<4, true><<<nsm, 512, smem_prod>>>(din, dref, dtw, dpp, L, 0, rows, 1, v3k::NttSrc{}); });
    const unsigned grid = (unsigned)(nsm * occ);
    const float lab = time([&] { lab::half_kernel<<<grid, 512, smem_half>>>(din, dout, dtw, dpp, L, 2 * rows); });
    cudaError_t err = cudaDeviceSynchronize();
    const double bytes = 2.0 * rows * N * 8;
    printf("half-row probe NCTA=%d VARIANT=%d occupancy %d CTAs/SM: production %.4f ms (%.0f GB/s)   probe %.4f ms (%.0f GB/s)  ratio %.3f [%s]\n", NCTA,
           VARIANT, occ, base, bytes / base / 1e6, lab, bytes / lab / 1e6, lab / base, cudaGetErrorString(err));
    return 0;
}
