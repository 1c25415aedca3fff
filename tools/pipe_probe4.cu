// pipe probe v4: cost model of the 64-bit modular-multiply building blocks on sm_100a.
// 512 threads x 1 CTA per SM (4 warps per SM sub-partition), 8 independent chains per thread,
// in-kernel SM clock.  Each OP is one "unit"; the SASS of every unit was checked with cuobjdump.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint64_t u64; typedef uint32_t u32;
#define ITER 512
__device__ __forceinline__ u64 mk(u32 lo, u32 hi) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }
__device__ __forceinline__ void sp(u64 x, u32& lo, u32& hi) { asm("mov.b64 {%0,%1}, %2;" : "=r"(lo), "=r"(hi) : "l"(x)); }
__device__ __forceinline__ u64 madw(u32 a, u32 b, u64 c) { u64 r; asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 mulw(u32 a, u32 b) { u64 r; asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ u32 madl(u32 a, u32 b, u32 c) { u32 r; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ u32 mulhi32(u32 a, u32 b) { u32 r; asm("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }

template <int OP>
__global__ void __launch_bounds__(512, 1) probe(u64* out, long long* cyc, u64 q, u64 w, u64 wp, u32 e) {
    u32 a = threadIdx.x * 2654435761u + 12345u, b = blockIdx.x * 40503u + 7u;
    u64 r[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { r[i] = (u64)(a * (i + 1) ^ b) * 0x9E3779B97F4A7C15ull; y[i] = r[i] ^ (w + i); }
    const u64 nq = 0 - q;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int rep = 0; rep < 2; rep++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                u32 r0, r1, y0, y1; sp(r[i], r0, r1); sp(y[i], y0, y1);
                if (OP == 0) r[i] = madw(r0, y0, r[i]);                 // WIDE acc in place
                if (OP == 1) r[i] = mulw(r0 ^ r1, y0);                  // WIDE, RZ addend (+1 LOP)
                if (OP == 2) r[i] = madw(r0, y0, y[i]);                 // WIDE, other addend
                if (OP == 3) r[i] = mk(madl(r0, y0, r1), r0);          // IMAD lo
                if (OP == 4) r[i] = mk(mulhi32(r0, y0), r0 + r1);      // IMAD.HI + IADD
                if (OP == 5) r[i] = __umul64hi(r[i], wp) + y[i];       // mulhi64 + add64
                if (OP == 6) r[i] = r[i] * w + y[i];                   // mullo64 + add64
                if (OP == 7) r[i] = r[i] + y[i];                        // add64
                if (OP == 8) { r[i] = r[i] - y[i] + q; }               // 3-input add64
                if (OP == 9) {                                          // Shoup (exact mulhi, acc tail)
                    u32 w0, w1, n0, n1; sp(w, w0, w1); sp(nq, n0, n1);
                    u64 h = __umul64hi(r[i], wp); u32 h0, h1; sp(h, h0, h1);
                    u64 acc = mulw(r0, w0); acc = madw(h0, n0, acc);
                    u32 l, hi; sp(acc, l, hi);
                    hi = madl(r0, w1, hi); hi = madl(r1, w0, hi); hi = madl(h0, n1, hi); hi = madl(h1, n0, hi);
                    r[i] = mk(l, hi) + y[i];
                }
                if (OP == 10) {                                         // full 128-bit product, sum of halves
                    u32 w0, w1; sp(w, w0, w1);
                    u64 t0_ = mulw(r0, w0); u32 t00, t01; sp(t0_, t00, t01);
                    u64 t1_ = madw(r1, w0, (u64)t01); u32 t10, t11; sp(t1_, t10, t11);
                    u64 t2_ = madw(r0, w1, (u64)t10); u32 t20, t21; sp(t2_, t20, t21);
                    u64 t3_ = madw(r1, w1, (u64)t11) + t21;
                    r[i] = t3_ ^ mk(t00, t20);
                }
                if (OP == 11) {                                         // Solinas fold at 2^60 (see DESIGN)
                    u32 w0, w1; sp(w, w0, w1);
                    u64 t0_ = mulw(r0, w0); u32 t00, t01; sp(t0_, t00, t01);
                    u64 t1_ = madw(r1, w0, (u64)t01); u32 t10, t11; sp(t1_, t10, t11);
                    u64 t2_ = madw(r0, w1, (u64)t10); u32 t20, t21; sp(t2_, t20, t21);
                    u64 t3_ = madw(r1, w1, (u64)t11) + t21; u32 t30, t31; sp(t3_, t30, t31);
                    // P = {t31,t30,t20,t00}; Ph = P >> 60, Pl = P & (2^60-1)
                    u32 ph0 = __funnelshift_r(t20, t30, 28), ph1 = __funnelshift_r(t30, t31, 28);
                    u64 m0 = mulw(ph0, e); u32 m00, m01; sp(m0, m00, m01);
                    u64 m1 = madw(ph1, e, (u64)m01); u32 m10, m11; sp(m1, m10, m11);
                    u32 mh = __funnelshift_r(m10, m11, 28);
                    // D = Pl - Ml (60-bit fields in 64-bit words), borrow -> bit 63
                    u64 D = mk(t00, t20 & 0x0fffffffu) - mk(m00, m10 & 0x0fffffffu);
                    u32 d0, d1; sp(D, d0, d1);
                    u32 k = mh + (d1 >> 31);
                    u64 res = madw(k, e, mk(d0, d1 & 0x0fffffffu));
                    r[i] = res + y[i];
                }
                if (OP == 12) {   // LOP3/SHF mix: 4 ALU ops
                    r0 = __funnelshift_r(r0, r1, 7) ^ y0; r1 = (r1 & y1) | r0; r[i] = mk(r0, r1);
                }
                if (OP == 13) {   // DFMA
                    double d = __longlong_as_double(r[i]); d = fma(d, 1.0000001, 0.5); r[i] = __double_as_longlong(d);
                }
            }
        }
    }
    long long t1 = clock64();
    u64 acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP>
void run(const char* name) {
    const int blocks = 148, threads = 512;
    static u64* out = nullptr; static long long* cyc = nullptr;
    if (!out) { cudaMalloc(&out, sizeof(u64) * blocks * threads); cudaMalloc(&cyc, sizeof(long long) * blocks); }
    const u64 q = 1152921504607338497ull;
    for (int k = 0; k < 2; k++) { probe<OP><<<blocks, threads>>>(out, cyc, q, 0x0123456789abcdefull % q, 0x1f3456789abcdef1ull, (u32)(q - (1ull << 60))); cudaDeviceSynchronize(); }
    static long long h[148]; cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; i++) avg += h[i]; avg /= blocks;
    printf("%-44s %7.2f SM cycles per warp-unit per SMSP\n", name, avg / ((double)ITER * 16 * 4.0));
}
int main() {
    run<0>("WIDE acc in place");
    run<1>("WIDE RZ addend (+1 LOP3)");
    run<2>("WIDE other addend");
    run<3>("IMAD lo");
    run<4>("IMAD.HI + IADD");
    run<5>("umul64hi + add64");
    run<6>("mullo64 + add64");
    run<7>("add64");
    run<8>("sub/add64 3-input");
    run<9>("Shoup exact (acc tail) + add64");
    run<10>("full 128-bit product");
    run<11>("Solinas fold @2^60 + add64");
    run<12>("4 ALU (SHF/LOP3)");
    run<13>("DFMA");
    return 0;
}
