// Butterfly formulations for q = 2^60 + e primes (crt.jl:282-295 chain), in-kernel SM clock + exactness check.
//  V0  exact mulhi (compiler __umul64hi) + generic acc tail            -- today's shoup_lazy, T in [0,2q)
//  V1  exact mulhi + special-q tail (h*q = h*e + h<<60)                 -- T in [0,2q)
//  V2  approx mulhi: 1 WIDE + 2 IMAD.HI, special-q tail                 -- T in [0,4q)
//  V3  approx mulhi: 1 WIDE + 2 DFMA.RM (FP64 pipe), special-q tail     -- T in [0,4q)
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
typedef uint64_t u64; typedef uint32_t u32; typedef unsigned __int128 u128;
struct __align__(16) tw_t { u64 w, wp; };
struct __align__(8) twd_t { double w0s, w1s, k; };
__device__ __forceinline__ u64 mk(u32 lo, u32 hi) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }
__device__ __forceinline__ void sp(u64 x, u32& lo, u32& hi) { asm("mov.b64 {%0,%1}, %2;" : "=r"(lo), "=r"(hi) : "l"(x)); }
__device__ __forceinline__ u64 madw(u32 a, u32 b, u64 c) { u64 r; asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 mulw(u32 a, u32 b) { u64 r; asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ u32 madl(u32 a, u32 b, u32 c) { u32 r; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ u32 mulhi32(u32 a, u32 b) { u32 r; asm("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }

struct QC { u64 q, nq; u32 e, ne; };   // q = 2^60 + e, ne = 2^32 - e

// special-q tail: r = lo64(Y*w - h*q), h*q = h*e + (h << 60)
__device__ __forceinline__ u64 tail_special(u32 y0, u32 y1, u32 w0, u32 w1, u64 h, const QC& c) {
    u32 h0, h1; sp(h, h0, h1);
    u64 acc = mulw(y0, w0);
    acc = madw(h0, c.ne, acc);                   // + h0*(2^32 - e)  (the 2^32*h0 part is removed below)
    u32 l, hi; sp(acc, l, hi);
    hi = madl(y0, w1, hi); hi = madl(y1, w0, hi); hi = madl(h1, c.ne, hi);
    hi = hi - h0 - (h0 << 28);
    return mk(l, hi);
}
template <int V>
__device__ __forceinline__ u64 modmul(u64 Y, tw_t w, twd_t wd, const QC& c) {
    u32 y0, y1, p0, p1, w0, w1; sp(Y, y0, y1); sp(w.wp, p0, p1); sp(w.w, w0, w1);
    if (V == 0) {
        u64 h = __umul64hi(Y, w.wp); u32 h0, h1, n0, n1; sp(h, h0, h1); sp(c.nq, n0, n1);
        u64 acc = mulw(y0, w0); acc = madw(h0, n0, acc);
        u32 l, hi; sp(acc, l, hi);
        hi = madl(y0, w1, hi); hi = madl(y1, w0, hi); hi = madl(h0, n1, hi); hi = madl(h1, n0, hi);
        return mk(l, hi);
    } else if (V == 1) {
        return tail_special(y0, y1, w0, w1, __umul64hi(Y, w.wp), c);
    } else if (V == 2) {
        u64 t = mulw(y1, p1);
        u32 a = mulhi32(y1, p0), b = mulhi32(y0, p1);
        u32 t0, t1, h0, h1; sp(t, t0, t1);
        (void)t0; (void)t1; (void)h0; (void)h1;
        return tail_special(y0, y1, w0, w1, t + (u64)a + (u64)b, c);
    } else if (V == 4) {
        return tail_special(y0, y1, w0, w1, mulw(y1, p1), c);
    } else if (V == 5) {
        return tail_special(y0, y1, w0, w1, Y ^ w.wp, c);
    } else if (V == 7) {
        return Y ^ w.w;
    } else if (V == 8) {   // tail with h = Y, only the two WIDEs of the tail
        u64 acc = mulw(y0, w0); acc = madw(y1, c.ne, acc); return acc ^ w.wp;
    } else if (V == 9) {   // four IMAD lo only
        u32 hi = madl(y0, w1, y1); hi = madl(y1, w0, hi); hi = madl(y0, c.ne, hi); hi = madl(hi, p0, p1); return mk(y0, hi);
    } else {
        u64 t = mulw(y1, p1);
        double m1 = __hiloint2double(0x43300000, (int)y1), m0 = __hiloint2double(0x43300000, (int)y0);
        double d1 = (V == 6) ? fma(m1, wd.w0s, wd.k) : __fma_rd(m1, wd.w0s, wd.k);
        double d2 = (V == 6) ? fma(m0, wd.w1s, d1) : __fma_rd(m0, wd.w1s, d1);
        u64 cc = (u64)__double_as_longlong(d2) - 0x4330000000000000ull;
        return tail_special(y0, y1, w0, w1, t + cc, c);
    }
}
template <int V>
__device__ __forceinline__ void bfly(u64& X, u64& Y, tw_t w, twd_t wd, const QC& c, u64 off) {
    u64 t = modmul<V>(Y, w, wd, c);
    u64 x = X;
    X = x + t;
    Y = x - t + off;
}
template <int V>
__global__ void __launch_bounds__(512, 1) k(u64* data, const tw_t* tw, const twd_t* twd, QC c, int iters, long long* cyc) {
    u64 x[32];
    const u64 off = 4 * c.q;
    for (int i = 0; i < 32; i++) x[i] = data[(size_t)blockIdx.x * blockDim.x * 32 + threadIdx.x + blockDim.x * i];
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 1; u <= 5; u++) {
            const int half = 32 >> u;
#pragma unroll
            for (int j = 0; j < (1 << (u - 1)); j++) {
                const int ti = (1 << (u - 1)) + j + (it & 7) * 32;
                tw_t w = tw[ti];
                twd_t wd; if (V == 3 || V == 6) wd = twd[ti]; else wd = twd_t{0, 0, 0};
#pragma unroll
                for (int kk = 0; kk < half; kk++) bfly<V>(x[j * 2 * half + kk], x[j * 2 * half + kk + half], w, wd, c, off);
            }
        }
    }
    long long t1 = clock64();
    for (int i = 0; i < 32; i++) data[(size_t)blockIdx.x * blockDim.x * 32 + threadIdx.x + blockDim.x * i] = x[i];
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// exactness: T == Y*w (mod q) and T < bound*q for random Y (any 64-bit) and the table twiddles
template <int V>
__global__ void check(const u64* ys, const tw_t* tw, const twd_t* twd, QC c, int n, int ntw, unsigned long long* bad, unsigned long long* maxk) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 Y = ys[i]; tw_t w = tw[i % ntw]; twd_t wd = twd[i % ntw];
    u64 T = modmul<V>(Y, w, wd, c);
    u64 ex = (u64)((u128)Y * w.w % c.q);
    u64 k = T / c.q;
    if (T % c.q != ex) atomicAdd(bad, 1ull);
    atomicMax(maxk, (unsigned long long)k);
}
static u64 splitmix(u64& s) { u64 z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
template <int V>
void run(const char* name, u64* d, tw_t* tw, twd_t* twd, QC c, long long* cyc, const u64* ys, int n, int ntw) {
    unsigned long long *bad, hb[2];
    cudaMalloc(&bad, 16); cudaMemset(bad, 0, 16);
    check<V><<<(n + 255) / 256, 256>>>(ys, tw, twd, c, n, ntw, bad, bad + 1);
    cudaMemcpy(hb, bad, 16, cudaMemcpyDeviceToHost); cudaFree(bad);
    const int iters = 64, blocks = 148, threads = 512;
    k<V><<<blocks, threads>>>(d, tw, twd, c, 2, cyc);
    cudaDeviceSynchronize();
    k<V><<<blocks, threads>>>(d, tw, twd, c, iters, cyc);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; i++) avg += h[i]; avg /= blocks;
    printf("%-52s %6.2f SM-cycles per warp-bfly per SMSP   check: %llu mismatches of %d, max floor(T/q) = %llu\n", name,
           avg / (4.0 * iters * 80.0), hb[0], n, hb[1]);
}
int main() {
    const u64 q = 1152921504607338497ull;
    QC c; c.q = q; c.nq = 0 - q; c.e = (u32)(q - (1ull << 60)); c.ne = 0u - c.e;
    const int ntw = 4096, n = 1 << 22;
    static tw_t htw[ntw]; static twd_t htwd[ntw];
    u64 s = 42;
    for (int i = 0; i < ntw; i++) {
        u64 w = splitmix(s) % q; if (i == 0) w = q - 1; if (i == 1) w = 1; if (i == 2) w = (1ull << 60);
        u64 wp = (u64)(((u128)w << 64) / q);
        htw[i].w = w; htw[i].wp = wp;
        u32 W0 = (u32)wp, W1 = (u32)(wp >> 32);
        htwd[i].w0s = (double)W0 / 4294967296.0; htwd[i].w1s = (double)W1 / 4294967296.0;
        htwd[i].k = 4503599627370496.0 - 1048576.0 * ((double)W0 + (double)W1);
    }
    u64* hy = new u64[n];
    for (int i = 0; i < n; i++) hy[i] = splitmix(s);
    for (int i = 0; i < 64; i++) hy[i] = ~0ull - i;
    for (int i = 64; i < 128; i++) hy[i] = (u64)i - 64;
    u64 *d, *ys; tw_t* tw; twd_t* twd; long long* cyc;
    size_t nd = (size_t)148 * 512 * 32;
    cudaMalloc(&d, nd * 8); cudaMemset(d, 1, nd * 8);
    cudaMalloc(&tw, sizeof(htw)); cudaMemcpy(tw, htw, sizeof(htw), cudaMemcpyHostToDevice);
    cudaMalloc(&twd, sizeof(htwd)); cudaMemcpy(twd, htwd, sizeof(htwd), cudaMemcpyHostToDevice);
    cudaMalloc(&ys, (size_t)n * 8); cudaMemcpy(ys, hy, (size_t)n * 8, cudaMemcpyHostToDevice);
    cudaMalloc(&cyc, 148 * 8);
    run<0>("V0 exact mulhi, generic tail (today)", d, tw, twd, c, cyc, ys, n, ntw);
    run<1>("V1 exact mulhi, special-q tail", d, tw, twd, c, cyc, ys, n, ntw);
    run<2>("V2 approx mulhi 1 WIDE + 2 IMAD.HI, special-q tail", d, tw, twd, c, cyc, ys, n, ntw);
    run<3>("V3 approx mulhi 1 WIDE + 2 DFMA.RM, special-q tail", d, tw, twd, c, cyc, ys, n, ntw);
    run<6>("V6 = V3 with round-to-nearest DFMA (timing only)", d, tw, twd, c, cyc, ys, n, ntw);
    run<4>("V4 h = y1*p1 only (1 WIDE) + tail (timing only)", d, tw, twd, c, cyc, ys, n, ntw);
    run<5>("V5 no mulhi, tail only (timing only)", d, tw, twd, c, cyc, ys, n, ntw);
    run<8>("V8 two WIDE only (timing only)", d, tw, twd, c, cyc, ys, n, ntw);
    run<9>("V9 four IMAD lo only (timing only)", d, tw, twd, c, cyc, ys, n, ntw);
    run<7>("V7 add/sub only (timing only)", d, tw, twd, c, cyc, ys, n, ntw);
    return 0;
}
