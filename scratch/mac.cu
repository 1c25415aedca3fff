#include <cstdint>
typedef uint64_t u64; typedef uint32_t u32; typedef unsigned __int128 u128;
#ifndef V
#define V 0
#endif
__device__ __forceinline__ void sp(u64 x, u32& lo, u32& hi) { asm("mov.b64 {%0,%1}, %2;" : "=r"(lo), "=r"(hi) : "l"(x)); }
__device__ __forceinline__ u64 madw(u32 a, u32 b, u64 c) { u64 r; asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c)); return r; }
struct Tab { u64 c[16]; };
__global__ void k(const u64* in, u64* out, const __grid_constant__ Tab T) {
    u64 d[16];
    for (int i = 0; i < 16; i++) d[i] = in[threadIdx.x + 64 * i];
#if V == 0
    u128 a = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) a += (u128)d[i] * T.c[i];
    out[threadIdx.x] = (u64)a; out[threadIdx.x + 64] = (u64)(a >> 64);
#elif V == 1
    // 31-bit limb columns, q < 2^61; normalise every 4 terms
    u64 C0 = 0, C1 = 0, C2 = 0;
    const u64 M = 0x7fffffffull;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const u32 d0 = (u32)d[i] & 0x7fffffffu, d1 = (u32)(d[i] >> 31);
        const u32 b0 = (u32)T.c[i] & 0x7fffffffu, b1 = (u32)(T.c[i] >> 31);
        C0 = madw(d0, b0, C0);
        C1 = madw(d0, b1, C1);
        C1 = madw(d1, b0, C1);
        C2 = madw(d1, b1, C2);
        if ((i & 3) == 3) { C1 += C0 >> 31; C0 &= M; C2 += C1 >> 31; C1 &= M; }
    }
    out[threadIdx.x] = C0 + (C1 << 31) + (C2 << 62); out[threadIdx.x + 64] = C2 >> 2;
#endif
}
