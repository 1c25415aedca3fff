// pipe probe 3: additivity of FMA-pipe (IMAD/IMAD.WIDE) and ALU-pipe instructions on sm_100a.
// 8 dependent chains per thread; unit = one step of every chain. SASS checked with tools/sass_mix.sh.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint64_t u64; typedef uint32_t u32;
#define ITER 512
__device__ __forceinline__ u32 lo32(u64 x) { u32 l, h; asm("mov.b64 {%0,%1}, %2;" : "=r"(l), "=r"(h) : "l"(x)); return l; }
template <int OP>
__global__ void __launch_bounds__(512, 1) probe(u64* out, long long* cyc, u32 b, u64 t) {
    u64 r[8]; u32 s[8]; u32 z[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { r[i] = (u64)(threadIdx.x * 2654435761u + i) * 0x9E3779B97F4A7C15ull; s[i] = threadIdx.x + i * 77; z[i] = threadIdx.x * 31 + i; }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int rep = 0; rep < 8; rep++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (OP == 0) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(r[i]) : "r"(lo32(r[i])), "r"(b));
                if (OP == 1) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(r[i]) : "r"(lo32(r[i])), "r"(b));
                if (OP == 2) asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r[i]) : "r"(lo32(r[i])), "r"(b), "l"(t));
                if (OP == 3) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(s[i]) : "r"(b), "r"(z[i]));
                if (OP == 4) { asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(r[i]) : "r"(lo32(r[i])), "r"(b)); asm volatile("xor.b32 %0, %0, %1;" : "+r"(z[i]) : "r"(b)); }
                if (OP == 5) { asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(r[i]) : "r"(lo32(r[i])), "r"(b)); asm volatile("xor.b32 %0, %0, %1;" : "+r"(z[i]) : "r"(b)); asm volatile("and.b32 %0, %0, %1;" : "+r"(s[i]) : "r"(z[(i+1)&7])); }
                if (OP == 6) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(s[i]) : "r"(b), "r"(z[i])); asm volatile("xor.b32 %0, %0, %1;" : "+r"(z[i]) : "r"(b)); }
                if (OP == 7) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(s[i]) : "r"(b), "r"(z[(i+3)&7])); asm volatile("xor.b32 %0, %0, %1;" : "+r"(z[i]) : "r"(b)); asm volatile("and.b32 %0, %0, %1;" : "+r"(z[(i+5)&7]) : "r"(z[(i+1)&7])); }
                if (OP == 8) asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(s[i]), "+r"(z[i]) : "r"(b), "r"(b));
                if (OP == 9) { asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(r[i]) : "r"(lo32(r[i])), "r"(b)); asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(s[i]), "+r"(z[i]) : "r"(b), "r"(b)); }
                if (OP == 10) asm volatile("xor.b32 %0, %0, %1;" : "+r"(z[i]) : "r"(b));
                if (OP == 11) { asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(r[i]) : "r"(lo32(r[i])), "r"(b)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(s[i]) : "r"(b), "r"(z[i])); }
            }
        }
    }
    long long t1 = clock64();
    u64 acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= r[i] ^ s[i] ^ ((u64)z[i] << 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP>
void run(const char* name) {
    const int blocks = 148, threads = 512;
    u64* out; long long* cyc;
    cudaMalloc(&out, sizeof(u64) * blocks * threads);
    cudaMalloc(&cyc, sizeof(long long) * blocks);
    probe<OP><<<blocks, threads>>>(out, cyc, 12345u, 0x123456789abcdefull);
    cudaDeviceSynchronize();
    probe<OP><<<blocks, threads>>>(out, cyc, 12345u, 0x123456789abcdefull);
    cudaDeviceSynchronize();
    static long long h[148]; cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; i++) avg += h[i]; avg /= blocks;
    printf("%-40s %6.3f SM-cycles per warp-unit per SMSP (4 warps/SMSP)\n", name, avg / ((double)ITER * 64 * 4));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>("WIDE (RZ addend)");
    run<1>("WIDE (acc in place)");
    run<2>("WIDE (other 64-bit addend)");
    run<3>("IMAD lo");
    run<10>("LOP3");
    run<8>("IADD3.cc + IADD3.X (64-bit add)");
    run<4>("WIDE + LOP3");
    run<5>("WIDE + 2 LOP3");
    run<6>("IMAD lo + LOP3");
    run<7>("IMAD lo + 2 LOP3");
    run<9>("WIDE + 64-bit add");
    run<11>("WIDE + IMAD lo");
    return 0;
}
