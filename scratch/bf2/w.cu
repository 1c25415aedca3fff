#include <cstdint>
typedef uint64_t u64; typedef uint32_t u32;
struct __align__(16) tw_t { u64 w, wp; };
#ifndef V
#define V 0
#endif
__device__ __forceinline__ u64 mk(u32 lo, u32 hi) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }
__device__ __forceinline__ void sp(u64 x, u32& lo, u32& hi) { asm("mov.b64 {%0,%1}, %2;" : "=r"(lo), "=r"(hi) : "l"(x)); }
__device__ __forceinline__ u64 madw(u32 a, u32 b, u64 c) { u64 r; asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 mulw(u32 a, u32 b) { u64 r; asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ u32 madl(u32 a, u32 b, u32 c) { u32 r; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }

// X' = X + T, Y' = X - T + OFF, T = y*w - h*q (h approx, T in [0,4q))
__device__ __forceinline__ void bfly(u64& X, u64& Y, tw_t w, u64 q, u64 nq, u64 off) {
    u32 y0, y1, p0, p1, w0, w1, n0, n1;
    sp(Y, y0, y1); sp(w.wp, p0, p1); sp(w.w, w0, w1); sp(nq, n0, n1);
#if V == 0   // approx quotient, 3 WIDE + wide-add
    u64 a = mulw(y1, p0);
    u64 c = mulw(y0, p1);
    u32 a0, a1, c0, c1; sp(a, a0, a1); sp(c, c0, c1);
    u64 h = madw(y1, p1, (u64)a1);
    h = madw(c1, 1, h);
#elif V == 1  // approx quotient, 3 WIDE + alu add
    u64 a = mulw(y1, p0);
    u64 c = mulw(y0, p1);
    u64 h = madw(y1, p1, a >> 32) + (c >> 32);
#elif V == 2  // exact
    u64 h = __umul64hi(Y, w.wp);
#endif
    u32 h0, h1; sp(h, h0, h1);
    u64 acc = mulw(y0, w0);
    acc = madw(h0, n0, acc);
    u32 l, hi; sp(acc, l, hi);
    hi = madl(y0, w1, hi); hi = madl(y1, w0, hi); hi = madl(h0, n1, hi); hi = madl(h1, n0, hi);
    u64 t = mk(l, hi);
    u64 x = X;
    X = x + t;
    Y = x - t + off;
}
__global__ void __launch_bounds__(512, 1) k(u64* data, const tw_t* tw, u64 q, int iters) {
    u64 x[32];
    const u64 off = 4 * q, nq = 0 - q;
    for (int i = 0; i < 32; i++) x[i] = data[(size_t)blockIdx.x * blockDim.x * 32 + threadIdx.x + blockDim.x * i];
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 1; u <= 5; u++) {
            const int half = 32 >> u;
#pragma unroll
            for (int j = 0; j < (1 << (u - 1)); j++) {
                tw_t w = tw[(1 << (u - 1)) + j + (it & 7) * 32];
#pragma unroll
                for (int kk = 0; kk < half; kk++) bfly(x[j * 2 * half + kk], x[j * 2 * half + kk + half], w, q, nq, off);
            }
        }
    }
    for (int i = 0; i < 32; i++) data[(size_t)blockIdx.x * blockDim.x * 32 + threadIdx.x + blockDim.x * i] = x[i];
}
