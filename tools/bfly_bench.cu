// Isolated butterfly-ladder throughput probe: 32 residues per thread in registers,
// 5 CT levels (80 butterflies) repeated ITERS times, twiddles from a tiny L1-resident
// table.  Reports cycles per warp-butterfly per SM sub-partition for several
// formulations and occupancies -- the pipe-mix ceiling of the NTT kernels.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint64_t u64; typedef uint32_t u32;
struct __align__(16) tw_t { u64 w, wp; };

__device__ __forceinline__ u64 shoup_c(u64 y, u64 w, u64 wp, u64 q) { return y * w - __umul64hi(y, wp) * q; }
// explicit 32-bit limb formulation: hi = y1*p1 + hi32(y0*p1) + hi32(y1*p0 + lo32(y0*p1)) ... exact
__device__ __forceinline__ u64 shoup_limb(u64 y, u64 w, u64 wp, u64 nq) {
    u32 y0 = (u32)y, y1 = (u32)(y >> 32), p0 = (u32)wp, p1 = (u32)(wp >> 32);
    u64 t0 = (u64)y0 * p0;
    u64 t1 = (u64)y0 * p1 + (t0 >> 32);
    u64 t2 = (u64)y1 * p0 + (u32)t1;
    u64 h = (u64)y1 * p1 + (t1 >> 32) + (t2 >> 32);
    return y * w + h * nq;
}
// approximate quotient (drops y0*p0 and the low carry): T in [0,4q)
__device__ __forceinline__ u64 shoup_approx(u64 y, u64 w, u64 wp, u64 nq) {
    u32 y0 = (u32)y, y1 = (u32)(y >> 32), p0 = (u32)wp, p1 = (u32)(wp >> 32);
    u64 t1 = (u64)y0 * p1;
    u64 t2 = (u64)y1 * p0;
    u64 h = (u64)y1 * p1 + (t1 >> 32) + (t2 >> 32);
    return y * w + h * nq;
}
template <int V>
__device__ __forceinline__ void bfly(u64& X, u64& Y, tw_t w, u64 q, u64 q2, u64 nq) {
    u64 x = X;
    if (V == 0 || V == 3) x = x >= q2 ? x - q2 : x;            // Harvey csub
    u64 t = (V == 0 || V == 1) ? shoup_c(Y, w.w, w.wp, q) : (V == 2 || V == 3) ? shoup_limb(Y, w.w, w.wp, nq) : shoup_approx(Y, w.w, w.wp, nq);
    X = x + t;
    Y = x - t + q2;
}
template <int V>
__global__ void __launch_bounds__(512, 1) k(u64* data, const tw_t* tw, u64 q, int iters) {
    u64 x[32];
    const u64 q2 = 2 * q, nq = 0 - q;
    for (int i = 0; i < 32; i++) x[i] = data[(size_t)blockIdx.x * blockDim.x * 32 + threadIdx.x + blockDim.x * i];
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
        if (V != 0 && V != 3) {   // keep lazy values bounded between iterations (outside the measured mix, 32 ops / 80 bfly)
#pragma unroll
            for (int i = 0; i < 32; i++) x[i] = x[i] >> 4;
        }
    }
    for (int i = 0; i < 32; i++) data[(size_t)blockIdx.x * blockDim.x * 32 + threadIdx.x + blockDim.x * i] = x[i];
}
template <int V>
void run(const char* name, int threads, int blocks_per_sm, u64* d, tw_t* tw, u64 q) {
    const int iters = 64, blocks = 148 * blocks_per_sm;
    k<V><<<blocks, threads>>>(d, tw, q, 2);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<V><<<blocks, threads>>>(d, tw, q, iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double warps_per_smsp = threads / 32.0 * blocks_per_sm / 4.0;
    double cyc = ms * 1e-3 * clk * 1e3;                   // SM cycles elapsed (at nominal clock)
    double per = cyc / (warps_per_smsp * iters * 80.0);   // cycles per warp-butterfly per SMSP
    printf("%-34s thr=%4d blk/SM=%d warps/SMSP=%4.1f : %6.2f cyc/warp-bfly/SMSP  (%.3f ms)\n", name, threads, blocks_per_sm, warps_per_smsp, per, ms);
}
int main() {
    u64* d; tw_t* tw;
    size_t n = (size_t)148 * 4 * 1024 * 32;
    cudaMalloc(&d, n * 8); cudaMemset(d, 1, n * 8);
    cudaMalloc(&tw, 4096 * sizeof(tw_t)); cudaMemset(tw, 3, 4096 * sizeof(tw_t));
    const u64 q = 1152921504607338497ull;
    for (int cfg = 0; cfg < 3; cfg++) {
        int threads = cfg == 0 ? 512 : cfg == 1 ? 256 : 512, bps = cfg == 0 ? 1 : cfg == 1 ? 1 : 1;
        if (cfg == 1) { threads = 256; bps = 1; }      // 2 warps/SMSP
        if (cfg == 2) { threads = 512; bps = 1; }
    }
    run<0>("C mulhi + csub (current Harvey)", 512, 1, d, tw, q);
    run<1>("C mulhi, lazy", 512, 1, d, tw, q);
    run<3>("limb mulhi + csub", 512, 1, d, tw, q);
    run<2>("limb mulhi, lazy", 512, 1, d, tw, q);
    run<4>("approx mulhi, lazy", 512, 1, d, tw, q);
    run<1>("C mulhi, lazy (2 warps/SMSP)", 256, 1, d, tw, q);
    run<1>("C mulhi, lazy (1 warp/SMSP)", 128, 1, d, tw, q);
    run<4>("approx mulhi, lazy (2 warps/SMSP)", 256, 1, d, tw, q);
    return 0;
}
