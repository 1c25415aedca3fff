#include <cstdint>
typedef uint64_t u64; typedef uint32_t u32;
struct __align__(16) tw_t { u64 w, wp; };
#ifndef V
#define V 0
#endif
__device__ __forceinline__ u64 mk(u32 lo, u32 hi) { return ((u64)hi << 32) | lo; }
// r = lo64(y*w + h*nq), accumulate form
__device__ __forceinline__ u64 tail_acc(u64 y, u64 w, u64 h, u64 nq) {
    u32 y0 = (u32)y, y1 = (u32)(y >> 32), w0 = (u32)w, w1 = (u32)(w >> 32);
    u32 h0 = (u32)h, h1 = (u32)(h >> 32), n0 = (u32)nq, n1 = (u32)(nq >> 32);
    u64 acc = (u64)y0 * w0;
    acc += (u64)h0 * n0;
    u32 hi = (u32)(acc >> 32) + y0 * w1 + y1 * w0 + h0 * n1 + h1 * n0;
    return mk((u32)acc, hi);
}
__device__ __forceinline__ u64 shoup(u64 y, u64 w, u64 wp, u64 q, u64 nq) {
#if V == 0
    return y * w - __umul64hi(y, wp) * q;
#elif V == 1
    return tail_acc(y, w, __umul64hi(y, wp), nq);
#elif V == 2
    u32 y0 = (u32)y, y1 = (u32)(y >> 32), p0 = (u32)wp, p1 = (u32)(wp >> 32);
    u64 a = (u64)y1 * p0;
    u64 b = (u64)y0 * p1 + (a >> 32);
    u64 h = (u64)y1 * p1 + (b >> 32);
    return tail_acc(y, w, h, nq);
#elif V == 3
    u32 y0 = (u32)y, y1 = (u32)(y >> 32), p0 = (u32)wp, p1 = (u32)(wp >> 32);
    u32 a, b, h0, h1;
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(a) : "r"(y1), "r"(p0));
    asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(b) : "r"(y0), "r"(p1), "r"(a));
    u64 h = (u64)y1 * p1 + b;
    return tail_acc(y, w, h, nq);
#endif
}
__device__ __forceinline__ void bfly(u64& X, u64& Y, tw_t w, u64 q, u64 q2, u64 nq) {
    u64 x = X;
    u64 t = shoup(Y, w.w, w.wp, q, nq);
    X = x + t;
    Y = x - t + q2;
}
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
                for (int kk = 0; kk < half; kk++) bfly(x[j * 2 * half + kk], x[j * 2 * half + kk + half], w, q, q2, nq);
            }
        }
    }
    for (int i = 0; i < 32; i++) data[(size_t)blockIdx.x * blockDim.x * 32 + threadIdx.x + blockDim.x * i] = x[i];
}
