// Butterfly formulations at the PTX level (mad.wide chains), in-kernel SM clock.
#include <cstdint>
typedef uint64_t u64; typedef uint32_t u32;
struct __align__(16) tw_t { u64 w, wp; };
__device__ __forceinline__ u64 mk(u32 lo, u32 hi) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }
__device__ __forceinline__ void sp(u64 x, u32& lo, u32& hi) { asm("mov.b64 {%0,%1}, %2;" : "=r"(lo), "=r"(hi) : "l"(x)); }
__device__ __forceinline__ u64 madw(u32 a, u32 b, u64 c) { u64 r; asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 mulw(u32 a, u32 b) { u64 r; asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ u32 madl(u32 a, u32 b, u32 c) { u32 r; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }


template <int V>
__device__ __forceinline__ void bfly(u64& X, u64& Y, tw_t w, u64 q, u64 nq, u64 off) {
    u32 y0, y1, p0, p1, w0, w1, n0, n1;
    sp(Y, y0, y1); sp(w.wp, p0, p1); sp(w.w, w0, w1); sp(nq, n0, n1);
    u64 h;
    if (V == 0) {          // approx quotient: 3 WIDE + wide-add
        u64 a = mulw(y1, p0), c = mulw(y0, p1);
        u32 a0, a1, c0, c1; sp(a, a0, a1); sp(c, c0, c1);
        h = madw(y1, p1, (u64)a1);
        h = madw(c1, 1, h);
    } else if (V == 1) {   // approx quotient: 3 WIDE + alu add
        u64 a = mulw(y1, p0), c = mulw(y0, p1);
        h = madw(y1, p1, a >> 32) + (c >> 32);
    } else {               // exact
        h = __umul64hi(Y, w.wp);
    }
    u32 h0, h1; sp(h, h0, h1);
    u64 t;
    if (V <= 2) {
        u64 acc = mulw(y0, w0);
        acc = madw(h0, n0, acc);
        u32 l, hi; sp(acc, l, hi);
        hi = madl(y0, w1, hi); hi = madl(y1, w0, hi); hi = madl(h0, n1, hi); hi = madl(h1, n0, hi);
        t = mk(l, hi);
        u64 x = X;
        X = x + t;
        Y = x - t + off;
    } else if (V == 3) {   // X folded into the accumulate chain: X' = X + T for free, Y' = 2X + off - X'
        u64 x = X;
        u64 acc = madw(y0, w0, x);
        acc = madw(h0, n0, acc);
        u32 l, hi; sp(acc, l, hi);
        hi = madl(y0, w1, hi); hi = madl(y1, w0, hi); hi = madl(h0, n1, hi); hi = madl(h1, n0, hi);
        X = mk(l, hi);
        Y = (x << 1) + off - X;
    } else if (V == 4) {   // exact quotient, plain C tail (what ntt_core.cuh does today)
        t = Y * w.w - h * q;
        u64 x = X;
        X = x + t;
        Y = x - t + off;
    }
}

template <int V>
__global__ void __launch_bounds__(512, 1) k(u64* data, const tw_t* tw, u64 q, int iters, long long* cyc) {
    u64 x[32];
    const u64 off = 4 * q, nq = 0 - q;
    for (int i = 0; i < 32; i++) x[i] = data[(size_t)blockIdx.x * blockDim.x * 32 + threadIdx.x + blockDim.x * i];
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 1; u <= 5; u++) {
            const int half = 32 >> u;
#pragma unroll
            for (int j = 0; j < (1 << (u - 1)); j++) {
                tw_t w = tw[(1 << (u - 1)) + j + (it & 7) * 32];
#pragma unroll
                for (int kk = 0; kk < half; kk++) bfly<V>(x[j * 2 * half + kk], x[j * 2 * half + kk + half], w, q, nq, off);
            }
        }
    }
    long long t1 = clock64();
    for (int i = 0; i < 32; i++) data[(size_t)blockIdx.x * blockDim.x * 32 + threadIdx.x + blockDim.x * i] = x[i];
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
#include <cstdio>
#include <cuda_runtime.h>
template <int V>
void run(const char* name, int threads, u64* d, tw_t* tw, u64 q, long long* cyc) {
    const int iters = 64, blocks = 148;
    k<V><<<blocks, threads>>>(d, tw, q, 2, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<V><<<blocks, threads>>>(d, tw, q, iters, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; i++) avg += h[i]; avg /= blocks;
    double warps_per_smsp = threads / 32.0 / 4.0;
    printf("%-44s warps/SMSP=%3.1f : %6.2f SM-cycles per warp-bfly per SMSP  (%.3f ms)\n", name, warps_per_smsp,
           avg / (warps_per_smsp * iters * 80.0), ms);
}

template <int V, int E, int LV, int TH>
__global__ void __launch_bounds__(TH, 1) ke(u64* data, const tw_t* tw, u64 q, int iters, long long* cyc) {
    u64 x[E];
    const u64 off = 4 * q, nq = 0 - q;
    for (int i = 0; i < E; i++) x[i] = data[(size_t)blockIdx.x * blockDim.x * E + threadIdx.x + blockDim.x * i];
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 1; u <= LV; u++) {
            const int half = E >> u;
#pragma unroll
            for (int j = 0; j < (1 << (u - 1)); j++) {
                tw_t w = tw[(1 << (u - 1)) + j + (it & 7) * 32];
#pragma unroll
                for (int kk = 0; kk < half; kk++) bfly<V>(x[j * 2 * half + kk], x[j * 2 * half + kk + half], w, q, nq, off);
            }
        }
    }
    long long t1 = clock64();
    for (int i = 0; i < E; i++) data[(size_t)blockIdx.x * blockDim.x * E + threadIdx.x + blockDim.x * i] = x[i];
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int V, int E, int LV, int TH>
void rune(const char* name, u64* d, tw_t* tw, u64 q, long long* cyc) {
    const int iters = 64, blocks = 148;
    ke<V, E, LV, TH><<<blocks, TH>>>(d, tw, q, 2, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    ke<V, E, LV, TH><<<blocks, TH>>>(d, tw, q, iters, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; i++) avg += h[i]; avg /= blocks;
    double warps_per_smsp = TH / 32.0 / 4.0;
    printf("%-44s E=%d TH=%d warps/SMSP=%3.1f : %6.2f SM-cycles per warp-bfly per SMSP  (%.3f ms)\n", name, E, TH, warps_per_smsp,
           avg / (warps_per_smsp * iters * (E / 2 * LV)), ms);
}

__constant__ tw_t ctw[512];
template <int V, int E, int LV, int TH>
__global__ void __launch_bounds__(TH, 1) kc(u64* data, u64 q, int iters, long long* cyc) {
    u64 x[E];
    const u64 off = 4 * q, nq = 0 - q;
    for (int i = 0; i < E; i++) x[i] = data[(size_t)blockIdx.x * blockDim.x * E + threadIdx.x + blockDim.x * i];
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 1; u <= LV; u++) {
            const int half = E >> u;
#pragma unroll
            for (int j = 0; j < (1 << (u - 1)); j++) {
                tw_t w = ctw[(1 << (u - 1)) + j + (it & 7) * 32];
#pragma unroll
                for (int kk = 0; kk < half; kk++) bfly<V>(x[j * 2 * half + kk], x[j * 2 * half + kk + half], w, q, nq, off);
            }
        }
    }
    long long t1 = clock64();
    for (int i = 0; i < E; i++) data[(size_t)blockIdx.x * blockDim.x * E + threadIdx.x + blockDim.x * i] = x[i];
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int V, int E, int LV, int TH>
void runc(const char* name, u64* d, u64 q, long long* cyc) {
    const int iters = 64, blocks = 148;
    kc<V, E, LV, TH><<<blocks, TH>>>(d, q, 2, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kc<V, E, LV, TH><<<blocks, TH>>>(d, q, iters, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; i++) avg += h[i]; avg /= blocks;
    double warps_per_smsp = TH / 32.0 / 4.0;
    printf("%-44s E=%d TH=%d warps/SMSP=%3.1f : %6.2f SM-cycles per warp-bfly per SMSP  (%.3f ms)\n", name, E, TH, warps_per_smsp,
           avg / (warps_per_smsp * iters * (E / 2 * LV)), ms);
}
int main() {
    u64* d; tw_t* tw; long long* cyc;
    size_t n = (size_t)148 * 1024 * 32;
    cudaMalloc(&d, n * 8); cudaMemset(d, 1, n * 8);
    cudaMalloc(&tw, 4096 * sizeof(tw_t)); cudaMemset(tw, 3, 4096 * sizeof(tw_t));
    cudaMalloc(&cyc, 148 * 8);
    const u64 q = 1152921504607338497ull;
    for (int th = 512; th >= 256; th -= 256) {
        run<4>("exact mulhi, C tail (current)", th, d, tw, q, cyc);
        run<2>("exact mulhi, PTX acc tail", th, d, tw, q, cyc);
        run<1>("approx 3-WIDE + alu add, PTX acc tail", th, d, tw, q, cyc);
        run<0>("approx 3-WIDE + wide add, PTX acc tail", th, d, tw, q, cyc);
        run<3>("exact mulhi, X folded into acc chain", th, d, tw, q, cyc);
    }
    { static tw_t h[512]; for (int i = 0; i < 512; i++) { h[i].w = 0x0303030303030303ull; h[i].wp = 0x0303030303030303ull; } cudaMemcpyToSymbol(ctw, h, sizeof(h)); }
    runc<2, 32, 5, 512>("exact PTX tail, twiddles in constant bank", d, q, cyc);
    runc<0, 32, 5, 512>("approx PTX tail, twiddles in constant bank", d, q, cyc);
    rune<2, 16, 4, 1024>("exact PTX tail", d, tw, q, cyc);
    rune<2, 16, 4, 512>("exact PTX tail", d, tw, q, cyc);
    rune<2, 8, 3, 1024>("exact PTX tail", d, tw, q, cyc);
    rune<2, 32, 5, 512>("exact PTX tail", d, tw, q, cyc);
    rune<2, 32, 5, 384>("exact PTX tail", d, tw, q, cyc);
    rune<0, 16, 4, 1024>("approx PTX tail", d, tw, q, cyc);
    return 0;
}
