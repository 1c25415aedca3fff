#include <cstdint>
#include <cuda_runtime.h>
typedef uint64_t u64; typedef uint32_t u32;
struct __align__(16) tw_t { u64 w, wp; };
__device__ __forceinline__ u64 mk(u32 lo, u32 hi) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }
__device__ __forceinline__ void sp(u64 x, u32& lo, u32& hi) { asm("mov.b64 {%0,%1}, %2;" : "=r"(lo), "=r"(hi) : "l"(x)); }
__device__ __forceinline__ u64 madw(u32 a, u32 b, u64 c) { u64 r; asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 mulw(u32 a, u32 b) { u64 r; asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ u32 madl(u32 a, u32 b, u32 c) { u32 r; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ u32 mulhi32(u32 a, u32 b) { u32 r; asm("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
struct QC { u64 q, nq; u32 e, ne; u32 shb; };

template <int V>
__device__ __forceinline__ u64 modmul(u64 Y, tw_t w, const QC& c) {
    u32 y0, y1, p0, p1, w0, w1; sp(Y, y0, y1); sp(w.wp, p0, p1); sp(w.w, w0, w1);
    if (V == 2) {   // production shoup_lazy4<28>
        u64 t = mulw(y1, p1);
        u32 a = mulhi32(y1, p0), b = mulhi32(y0, p1);
        u32 h0, h1; sp(t + (u64)a + (u64)b, h0, h1);
        u64 acc = mulw(y0, w0);
        acc = madw(h0, c.ne, acc);
        u32 l, hi; sp(acc, l, hi);
        hi = madl(y0, w1, hi); hi = madl(y1, w0, hi); hi = madl(h1, c.ne, hi);
        hi = hi - h0 - (h0 << 28);
        return mk(l, hi);
    } else if (V == 10) {  // explicit carry chain for h, runtime shift
        u64 t = mulw(y1, p1);
        u32 a = mulhi32(y1, p0), b = mulhi32(y0, p1);
        u32 t0, t1, h0, h1; sp(t, t0, t1);
        asm("{\n\t.reg .u32 s;\n\tadd.cc.u32 s, %2, %3;\n\taddc.u32 %1, %5, 0;\n\tadd.cc.u32 %0, s, %4;\n\taddc.u32 %1, %1, 0;\n\t}" : "=r"(h0), "=&r"(h1) : "r"(t0), "r"(a), "r"(b), "r"(t1));
        u64 acc = mulw(h0, c.ne);
        acc = madw(y0, w0, acc);
        u32 l, hi; sp(acc, l, hi);
        hi = madl(y0, w1, hi); hi = madl(y1, w0, hi); hi = madl(h1, c.ne, hi);
        hi = hi - h0 - (h0 << c.shb);
        return mk(l, hi);
    } else if (V == 11) {  // as production but runtime shift only
        u64 t = mulw(y1, p1);
        u32 a = mulhi32(y1, p0), b = mulhi32(y0, p1);
        u32 h0, h1; sp(t + (u64)a + (u64)b, h0, h1);
        u64 acc = mulw(y0, w0);
        acc = madw(h0, c.ne, acc);
        u32 l, hi; sp(acc, l, hi);
        hi = madl(y0, w1, hi); hi = madl(y1, w0, hi); hi = madl(h1, c.ne, hi);
        hi = hi - h0 - (h0 << c.shb);
        return mk(l, hi);
    } else if (V == 12) {  // h*q directly: q = (q1:q0); r = y*w - h*q, with q0 = e (q0<2^32), q1 = 2^28: hi -= h0<<28
        u64 t = mulw(y1, p1);
        u32 a = mulhi32(y1, p0), b = mulhi32(y0, p1);
        u32 h0, h1; sp(t + (u64)a + (u64)b, h0, h1);
        u64 acc = mulw(y0, w0);
        u64 he = mulw(h0, c.e);
        acc -= he;
        u32 l, hi; sp(acc, l, hi);
        hi = madl(y0, w1, hi); hi = madl(y1, w0, hi); hi = madl(h1, c.ne, hi);
        hi = hi - (h0 << c.shb);
        return mk(l, hi);
    }
    return 0;
}
template <int V>
__device__ __forceinline__ void bfly(u64& X, u64& Y, tw_t w, const QC& c, u64 off) {
    u64 t = modmul<V>(Y, w, c);
    u64 x = X;
    X = x + t;
    Y = x - t + off;
}
template <int V>
__global__ void __launch_bounds__(512, 1) k(u64* data, const tw_t* tw, QC c, int iters) {
    u64 x[32];
    const u64 off = 4 * c.q;
    for (int i = 0; i < 32; i++) x[i] = data[(size_t)blockIdx.x * blockDim.x * 32 + threadIdx.x + blockDim.x * i];
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 1; u <= 5; u++) {
            const int half = 32 >> u;
#pragma unroll
            for (int j = 0; j < (1 << (u - 1)); j++) {
                const int ti = (1 << (u - 1)) + j + (it & 7) * 32;
                tw_t w = tw[ti];
#pragma unroll
                for (int kk = 0; kk < half; kk++) bfly<V>(x[j * 2 * half + kk], x[j * 2 * half + kk + half], w, c, off);
            }
        }
    }
    for (int i = 0; i < 32; i++) data[(size_t)blockIdx.x * blockDim.x * 32 + threadIdx.x + blockDim.x * i] = x[i];
}
template __global__ void k<2>(u64*, const tw_t*, QC, int);
template __global__ void k<10>(u64*, const tw_t*, QC, int);
template __global__ void k<11>(u64*, const tw_t*, QC, int);
template __global__ void k<12>(u64*, const tw_t*, QC, int);
