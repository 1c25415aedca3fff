#include <cstdint>
typedef uint64_t u64; typedef uint32_t u32;
struct tw_t { u64 w, wp; };

// variant 0: plain C (current)
__device__ __forceinline__ u64 shoup0(u64 y, u64 w, u64 wp, u64 q) { return y * w - __umul64hi(y, wp) * q; }

// variant 1: PTX, exact mulhi via mad.wide chain, low part folded with nq = -q
__device__ __forceinline__ u64 shoup1(u64 y, u64 w, u64 wp, u64 nq) {
    u64 r;
    asm("{\n\t"
        ".reg .u32 y0,y1,p0,p1,w0,w1,n0,n1,h0,h1,c0,t1l,t1h,t2l,t2h,al,ah,dz;\n\t"
        ".reg .u64 t0,t1,t2,t3,s,s2,a;\n\t"
        "mov.b64 {y0,y1}, %1;\n\t"
        "mov.b64 {w0,w1}, %2;\n\t"
        "mov.b64 {p0,p1}, %3;\n\t"
        "mov.b64 {n0,n1}, %4;\n\t"
        "mul.wide.u32 t0, y0, p0;\n\t"
        "mov.b64 {dz,c0}, t0;\n\t"
        "cvt.u64.u32 s, c0;\n\t"
        "mad.wide.u32 t1, y0, p1, s;\n\t"
        "mov.b64 {t1l,t1h}, t1;\n\t"
        "cvt.u64.u32 s, t1l;\n\t"
        "mad.wide.u32 t2, y1, p0, s;\n\t"
        "mov.b64 {t2l,t2h}, t2;\n\t"
        "cvt.u64.u32 s, t1h;\n\t"
        "cvt.u64.u32 s2, t2h;\n\t"
        "add.u64 s, s, s2;\n\t"
        "mad.wide.u32 t3, y1, p1, s;\n\t"
        "mov.b64 {h0,h1}, t3;\n\t"
        "mul.wide.u32 a, y0, w0;\n\t"
        "mad.wide.u32 a, h0, n0, a;\n\t"
        "mov.b64 {al,ah}, a;\n\t"
        "mad.lo.u32 ah, y0, w1, ah;\n\t"
        "mad.lo.u32 ah, y1, w0, ah;\n\t"
        "mad.lo.u32 ah, h0, n1, ah;\n\t"
        "mad.lo.u32 ah, h1, n0, ah;\n\t"
        "mov.b64 %0, {al,ah};\n\t"
        "}" : "=l"(r) : "l"(y), "l"(w), "l"(wp), "l"(nq));
    return r;
}
// variant 2: C with explicit 32-bit limbs
__device__ __forceinline__ u64 shoup2(u64 y, u64 w, u64 wp, u64 nq) {
    u32 y0 = (u32)y, y1 = (u32)(y >> 32), p0 = (u32)wp, p1 = (u32)(wp >> 32);
    u64 t0 = (u64)y0 * p0;
    u64 t1 = (u64)y0 * p1 + (t0 >> 32);
    u64 t2 = (u64)y1 * p0 + (u32)t1;
    u64 h = (u64)y1 * p1 + (t1 >> 32) + (t2 >> 32);
    return y * w + h * nq;
}

template <int V>
__device__ __forceinline__ void bfly(u64& X, u64& Y, tw_t w, u64 q, u64 q2, u64 nq) {
    u64 x = X >= q2 ? X - q2 : X;
    u64 t = V == 0 ? shoup0(Y, w.w, w.wp, q) : V == 1 ? shoup1(Y, w.w, w.wp, nq) : shoup2(Y, w.w, w.wp, nq);
    X = x + t;
    Y = x - t + q2;
}
// no csub (lazy)
template <int V>
__device__ __forceinline__ void bfly_nc(u64& X, u64& Y, tw_t w, u64 q, u64 q2, u64 nq) {
    u64 x = X;
    u64 t = V == 0 ? shoup0(Y, w.w, w.wp, q) : V == 1 ? shoup1(Y, w.w, w.wp, nq) : shoup2(Y, w.w, w.wp, nq);
    X = x + t;
    Y = x - t + q2;
}

template <int V, int NC>
__global__ void k(u64* data, const tw_t* tw, u64 q) {
    u64 x[32];
    const u64 q2 = 2 * q, nq = 0 - q;
    for (int i = 0; i < 32; i++) x[i] = data[threadIdx.x + 512 * i];
#pragma unroll
    for (int u = 1; u <= 5; u++) {
        const int half = 32 >> u;
#pragma unroll
        for (int j = 0; j < (1 << (u - 1)); j++) {
            tw_t w = tw[(1 << (u - 1)) + j];
#pragma unroll
            for (int kk = 0; kk < half; kk++) {
                if (NC) bfly_nc<V>(x[j * 2 * half + kk], x[j * 2 * half + kk + half], w, q, q2, nq);
                else bfly<V>(x[j * 2 * half + kk], x[j * 2 * half + kk + half], w, q, q2, nq);
            }
        }
    }
    for (int i = 0; i < 32; i++) data[threadIdx.x + 512 * i] = x[i];
}
template __global__ void k<0,0>(u64*, const tw_t*, u64);
template __global__ void k<1,0>(u64*, const tw_t*, u64);
template __global__ void k<2,0>(u64*, const tw_t*, u64);
template __global__ void k<0,1>(u64*, const tw_t*, u64);
template __global__ void k<1,1>(u64*, const tw_t*, u64);
template __global__ void k<2,1>(u64*, const tw_t*, u64);
