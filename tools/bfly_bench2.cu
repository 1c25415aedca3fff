// Butterfly-ladder probe v2: like bfly_bench.cu but timed with the in-kernel SM clock
// (so DVFS does not distort cycles) and with limb-level formulations of the Shoup product.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint64_t u64; typedef uint32_t u32;
struct __align__(16) tw_t { u64 w, wp; };
__device__ __forceinline__ u64 mk(u32 lo, u32 hi) { return ((u64)hi << 32) | lo; }
__device__ __forceinline__ u64 tail_acc(u64 y, u64 w, u64 h, u64 nq) {
    u32 y0 = (u32)y, y1 = (u32)(y >> 32), w0 = (u32)w, w1 = (u32)(w >> 32);
    u32 h0 = (u32)h, h1 = (u32)(h >> 32), n0 = (u32)nq, n1 = (u32)(nq >> 32);
    u64 acc = (u64)y0 * w0;
    acc += (u64)h0 * n0;
    u32 hi = (u32)(acc >> 32) + y0 * w1 + y1 * w0 + h0 * n1 + h1 * n0;
    return mk((u32)acc, hi);
}
template <int V>
__device__ __forceinline__ u64 shoup(u64 y, u64 w, u64 wp, u64 q, u64 nq) {
    if (V == 0) return y * w - __umul64hi(y, wp) * q;
    if (V == 1) return tail_acc(y, w, __umul64hi(y, wp), nq);
    u32 y0 = (u32)y, y1 = (u32)(y >> 32), p0 = (u32)wp, p1 = (u32)(wp >> 32);
    if (V == 2) {
        u64 a = (u64)y1 * p0;
        u64 b = (u64)y0 * p1 + (a >> 32);
        u64 h = (u64)y1 * p1 + (b >> 32);
        return tail_acc(y, w, h, nq);
    }
    if (V == 3) {
        u32 a, b;
        asm("mul.hi.u32 %0, %1, %2;" : "=r"(a) : "r"(y1), "r"(p0));
        asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(b) : "r"(y0), "r"(p1), "r"(a));
        u64 h = (u64)y1 * p1 + b;
        return tail_acc(y, w, h, nq);
    }
    if (V == 4) {  // only the top product for the quotient (NOT exact: pipe-floor probe, 1 WIDE + tail)
        u64 h = (u64)y1 * p1;
        return tail_acc(y, w, h, nq);
    }
    return 0;
}
template <int V>
__device__ __forceinline__ void bfly(u64& X, u64& Y, tw_t w, u64 q, u64 q2, u64 nq) {
    u64 x = X;
    u64 t = shoup<V>(Y, w.w, w.wp, q, nq);
    X = x + t;
    Y = x - t + q2;
}
template <int V>
__global__ void __launch_bounds__(512, 1) k(u64* data, const tw_t* tw, u64 q, int iters, long long* cyc) {
    u64 x[32];
    const u64 q2 = 2 * q, nq = 0 - q;
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
                for (int kk = 0; kk < half; kk++) bfly<V>(x[j * 2 * half + kk], x[j * 2 * half + kk + half], w, q, q2, nq);
            }
        }
#pragma unroll
        for (int i = 0; i < 32; i++) x[i] = x[i] >> 4;
    }
    long long t1 = clock64();
    for (int i = 0; i < 32; i++) data[(size_t)blockIdx.x * blockDim.x * 32 + threadIdx.x + blockDim.x * i] = x[i];
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
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
    printf("%-30s warps/SMSP=%3.1f : %6.2f SM-cycles per warp-bfly per SMSP  (%.3f ms, clk %.0f MHz)\n", name, warps_per_smsp,
           avg / (warps_per_smsp * iters * 80.0), ms, avg / (ms * 1e3));
}
int main() {
    u64* d; tw_t* tw; long long* cyc;
    size_t n = (size_t)148 * 1024 * 32;
    cudaMalloc(&d, n * 8); cudaMemset(d, 1, n * 8);
    cudaMalloc(&tw, 4096 * sizeof(tw_t)); cudaMemset(tw, 3, 4096 * sizeof(tw_t));
    cudaMalloc(&cyc, 148 * 8);
    const u64 q = 1152921504607338497ull;
    for (int th = 512; th >= 256; th -= 256) {
        run<0>("C mulhi (exact)", th, d, tw, q, cyc);
        run<1>("C mulhi + acc tail", th, d, tw, q, cyc);
        run<2>("3-WIDE chain quotient + acc tail", th, d, tw, q, cyc);
        run<3>("2 IMAD.HI + WIDE + acc tail", th, d, tw, q, cyc);
        run<4>("floor probe: 1 WIDE + acc tail", th, d, tw, q, cyc);
    }
    return 0;
}
