// NTT kernel lab: the production forward kernel (v3k::ntt_fwd_s_kernel<4,true>, N = 2^14) next to ABLATED copies of
// itself, each built as its own binary (-DABL=<mask>) and timed with CUDA events on 2048 rows (8 primes x 256
// polynomials: 256 MiB in, 256 MiB out).  Ablations change the RESULT (they remove work) -- they exist to attribute the
// kernel's time to its phases; variants with bit 128 and above set are real candidates and are checked bit for bit
// against the production kernel.
//   1   pass-1 butterflies skipped          2  pass-2 butterflies skipped        4  pass-3 butterflies skipped
//   8   pass-3 twiddles: every thread reads the same 15 entries (L1 hits instead of 245 KiB per row from L2)
//   16  no canonicalisation before the stores                                    32 row loaded once (no TMA per row)
//   64  no global stores                                                          128 quotient with a fused carry chain (q4)
//   256 pass-3 twiddles loaded with L1::no_allocate                               512 barriers removed (races; timing only)
//   1024 CTAs sharing an SM start DEPHASE_NS apart (LAB_R < 4: 2 or 4 CTAs per SM)
//   2048 next row prefetched into L2 (cp.async.bulk.prefetch.L2) at the start of the current one
// build: tools/ntt_lab.sh   run: for each binary, prints one line
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../toyfhe.jl_b200/csrc/ntt_v3_kernels.cuh"
#include "../toyfhe.jl_b200/csrc/tables.h"

#ifndef ABL
#define ABL 0
#endif
#ifndef LAB_R
#define LAB_R 4   // rows of 2^(10+LAB_R) positions; 512 / T CTAs resident per SM
#endif
#ifndef DEPHASE_NS
#define DEPHASE_NS 4000
#endif

// stubs for the library symbols the header refers to
void tfb_set_error(const std::string&) {}
int tfb_cuda_fail(cudaError_t e, const char* what) { fprintf(stderr, "CUDA error %s in %s\n", cudaGetErrorString(e), what); return -1; }
ProfScope::ProfScope(int c, cudaStream_t s) : cls(c), st(s), stop(nullptr) {}
ProfScope::~ProfScope() {}

namespace lab {
using namespace v3;
constexpr int R = LAB_R;
typedef NttGeo<R> Geo;

__device__ __forceinline__ u64 sl4(u64 x, u64 w, u64 wp, u64 q, u32 ne, u32 shb) {
    if (!(ABL & 128)) return shoup_lazy4(x, w, wp, q, ne, shb);
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
template <bool NOALLOC>
__device__ __forceinline__ tw_t ldtw(const tw_t* p) {
    if (!NOALLOC) return *p;
    tw_t r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(r.w), "=l"(r.wp) : "l"(p));
    return r;
}
template <bool RED>
__device__ __forceinline__ void bfly(u64& X, u64& Y, const tw_t w, const Red3& rp) {
    const u64 t = sl4(Y, w.w, w.wp, rp.q, rp.ne, rp.shb);
    const u64 x = X;
    if (RED) {
        const redent_t c = rp.tab[top4(x, rp)];
        X = x + c.c2 + t;
        Y = x + c.c3 - t;
    } else {
        X = x + t;
        Y = x - t + rp.q4;
    }
}
template <int LV, u32 REDMASK, bool NOALLOC>
__device__ __forceinline__ void levels(u64* x, const tw_t* __restrict__ tw, const u32* tb, const Red3& rp, const u32 js = 1) {
#pragma unroll
    for (int u = 1; u <= LV; u++) {
        const int half = (1 << LV) >> u;
#pragma unroll
        for (int j = 0; j < (1 << (u - 1)); j++) {
            const tw_t w = ldtw<NOALLOC>(tw + tb[u - 1] + j * js);
#pragma unroll
            for (int k = 0; k < half; k++) {
                if ((REDMASK >> (u - 1)) & 1) bfly<true>(x[j * 2 * half + k], x[j * 2 * half + k + half], w, rp);
                else bfly<false>(x[j * 2 * half + k], x[j * 2 * half + k + half], w, rp);
            }
        }
    }
}

__global__ void __launch_bounds__(Geo::T, 512 / Geo::T)
fwd_kernel(const u64* __restrict__ in, u64* __restrict__ out, const tw_t* __restrict__ tw_all, const PrimeParams* __restrict__ pp,
           const u32 L, const u32 nunits, const u32 never, u32* __restrict__ sm_arrivals) {
    extern __shared__ __align__(128) u64 smem[];
    __shared__ __align__(8) u64 bar;
    __shared__ redent_t redtab[TFB_MAX_L * 16];
    u32 t = threadIdx.x;
    const u64 nrow = Geo::N;
    u32 unit = blockIdx.x;
    if (t == 0) {
        v3k::mbar_init(&bar, 1);
        v3k::fence_barrier_init();
    }
    v3k::build_redtab(redtab, pp, L, t, Geo::T);
    __syncthreads();
    if (t < 32 && unit < nunits) v3k::tma_load_row_skewed<R>(smem, in + (u64)unit * nrow, &bar, t);
    if (ABL & 1024) {   // de-phase the CTAs that share an SM: the k-th arrival waits k * DEPHASE_NS
        __shared__ u32 order;
        if (t == 0) {
            u32 smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            order = atomicAdd(&sm_arrivals[smid], 1u) % (512 / Geo::T);
        }
        __syncthreads();
        for (u32 k = 0; k < order; k++) __nanosleep(DEPHASE_NS);
    }
    u32 parity = 0;
    u64 x[32];
    bool first = true;
    for (; unit < nunits; unit += gridDim.x) {
        const u32 prime = unit % L;
        const tw_t* tw = tw_all + (u64)prime * nrow;
        const Red3 rp = make_red3(pp[prime].pc.q, pp[prime].sh, redtab + prime * 16);
        asm volatile("" : "+r"(t));
        if ((ABL & 2048) && t < 32 && unit + gridDim.x < nunits)   // the row after this one: DRAM -> L2 now, L2 -> shared memory (TMA) after pass 3's loads
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(in + (u64)(unit + gridDim.x) * nrow + t * Geo::T), "r"((u32)(Geo::T * 8)) : "memory");
        if (!(ABL & 32) || first) {
            v3k::mbar_wait(&bar, parity);
            parity ^= 1;
        }
        first = false;
        // pass 1
        {
#pragma unroll
            for (int a = 0; a < 32; a++) x[a] = smem[slot<R>(a, t)];
            u32 tb[5];
#pragma unroll
            for (int s = 1; s <= 5; s++) tb[s - 1] = 1u << (s - 1);
            if (!(ABL & 1)) levels<5, 0x08, false>(x, tw, tb, rp);
#pragma unroll
            for (int a = 0; a < 32; a++) smem[slot<R>(a, t)] = x[a];
        }
        if (!(ABL & 512)) __syncthreads();
        // pass 2
        {
            const u32 a2 = t >> R, c2 = t & (Geo::RS - 1);
            u64* base = smem + slot<R>(a2, c2);
#pragma unroll
            for (int b = 0; b < 32; b++) x[b] = base[b * Geo::RS];
            u32 tb[5];
#pragma unroll
            for (int u = 1; u <= 5; u++) tb[u - 1] = (1u << (4 + u)) + (a2 << (u - 1));
            if (!(ABL & 2)) levels<5, 0x09, false>(x, tw, tb, rp);
#pragma unroll
            for (int b = 0; b < 32; b++) base[b * Geo::RS] = x[b];
        }
        if (!(ABL & 512)) __syncthreads();
        pass3_load<R>(x, smem, t);
        if (!(ABL & 512)) __syncthreads();
        const u32 next = unit + gridDim.x;
        if (!(ABL & 32) && t < 32 && next < nunits) v3k::tma_load_row_skewed<R>(smem, in + (u64)next * nrow, &bar, t);
        // pass 3
        {
            const tw_t* twc = tw_all + (u64)(L + prime) * nrow;
            u64* orow = out + (u64)unit * nrow;
            const u32 w = t >> 5, lane = t & 31;
#pragma unroll
            for (int g = 0; g < (int)Geo::G; g++) {
                const u32 k2 = Geo::G * w + g;
                u32 tb[R];
#pragma unroll
                for (int u = 1; u <= R; u++) tb[u - 1] = (ABL & 8) ? (u32)((1u << (u - 1)) - 1) * Geo::T : pass3_base<R>(0, (u32)g, u, t);
                if (!(ABL & 4)) levels<R, Lay<R>::P3MASK, (ABL & 256) != 0>(x + g * Geo::RS, twc, tb, rp, Geo::T);
#pragma unroll
                for (int c = 0; c < (int)Geo::RS; c++) {
                    const u32 kl = (brev_bits((u32)c, R) << 10) | (k2 << 5) | lane;
                    const u64 v = (ABL & 16) ? x[g * Geo::RS + c] : canon3(x[g * Geo::RS + c], rp);
                    if (!(ABL & 64) || v == (u64)never) orow[kl] = v;
                }
            }
        }
    }
}
}  // namespace lab

static u64 splitmix(u64& s) { u64 z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }

int main(int argc, char** argv) {
    const u32 N = 1u << (10 + LAB_R), L = 8;
    const u32 rows = argc > 1 ? (u32)atoi(argv[1]) : (2048u << (4 - LAB_R));
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
        permute_pass3(ht.fwd.data(), &fwd[(size_t)(L + i) * N], 10 + LAB_R);
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
    const int smem = (int)v3::Lay<LAB_R>::ROW_BYTES;
    cudaFuncSetAttribute(v3k::ntt_fwd_s_kernel<LAB_R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    u32* arrivals; cudaMalloc(&arrivals, 1024 * 4);
    cudaFuncSetAttribute(lab::fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    const u32 slots = (u32)nsm * (512 / lab::Geo::T);
    const unsigned grid = rows < slots ? rows : slots;
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
    const float base = time([&] { v3k::ntt_fwd_s_kernel<LAB_R, true><<<grid, lab::Geo::T, smem>>>(din, dref, dtw, dpp, L, 0, rows, 1, v3k::NttSrc{}); });
    const float lab = time([&] { cudaMemsetAsync(arrivals, 0, 1024 * 4); lab::fwd_kernel<<<grid, lab::Geo::T, smem>>>(din, dout, dtw, dpp, L, rows, 0xdeadbeefu, arrivals); });
    cudaError_t err = cudaDeviceSynchronize();
    const double bytes = 2.0 * rows * N * 8;
    std::vector<u64> a(h.size()), b(h.size());
    cudaMemcpy(a.data(), dref, h.size() * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(b.data(), dout, h.size() * 8, cudaMemcpyDeviceToHost);
    size_t diff = 0;
    for (size_t i = 0; i < a.size(); i++) diff += a[i] != b[i];
    printf("R=%d ABL=%4d dephase=%d ns  production %.4f ms (%.0f GB/s)   lab %.4f ms (%.0f GB/s)  ratio %.3f  mismatching words %zu of %zu  [%s]\n", LAB_R, ABL, DEPHASE_NS, base,
           bytes / base / 1e6, lab, bytes / lab / 1e6, lab / base, diff, a.size(), cudaGetErrorString(err));
    return 0;
}
