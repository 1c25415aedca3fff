// pipe probe v2: in-kernel clock, non-hoistable operands
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 1024
template <int OP>
__global__ void probe(uint64_t* out, long long* cyc) {
    uint32_t a = threadIdx.x * 2654435761u + 12345u, b = blockIdx.x * 40503u + 7u;
    uint64_t r[8]; uint32_t s[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { r[i] = a * (i + 1) ^ b; s[i] = a + i * b; }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int rep = 0; rep < 4; rep++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(r[i]) : "r"(s[i]), "r"(b));          // wide acc, mult operands fixed per chain
            if (OP == 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(s[i]) : "r"(b), "r"(a));              // dependent mult
            if (OP == 2) asm volatile("add.u32 %0, %0, %1;" : "+r"(s[i]) : "r"(a));
            if (OP == 3) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(r[i]) : "r"(s[i]), "r"(b)); asm volatile("add.u32 %0, %0, %1;" : "+r"(s[i]) : "r"(a)); }
            if (OP == 4) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(s[i]) : "r"(b), "r"(a)); asm volatile("add.u32 %0, %0, %1;" : "+r"(s[(i+4)&7]) : "r"(a)); }
            if (OP == 5) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(r[i]) : "r"(s[i]), "r"(b)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(s[i]) : "r"(b), "r"(a)); }
            if (OP == 6) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(r[i]) : "r"(s[i]), "r"(b)); asm volatile("add.u32 %0, %0, %1;" : "+r"(s[i]) : "r"(a)); asm volatile("xor.b32 %0, %0, %1;" : "+r"(s[(i+3)&7]) : "r"(b)); }
            if (OP == 7) { uint32_t lo = (uint32_t)r[i], hi = (uint32_t)(r[i] >> 32); asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(lo), "r"(b)); r[i] = ((uint64_t)hi << 32) | lo; } // IMAD lo into hi half
            if (OP == 8) asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r[i]) : "r"(s[i]), "r"(b), "l"(r[(i+1)&7]));   // wide, different acc src
            if (OP == 9) { double d = __longlong_as_double(r[i]); d = fma(d, 1.0000001, 0.5); r[i] = __double_as_longlong(d); }
            if (OP == 10) { double d = __longlong_as_double(r[i]); d = fma(d, 1.0000001, 0.5); r[i] = __double_as_longlong(d); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(s[i]) : "r"(b), "r"(a)); }
            if (OP == 11) { asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(s[i]), "+r"(s[(i+1)&7]) : "r"(a), "r"(b)); }
        }
        }
    }
    long long t1 = clock64();
    uint64_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= r[i] ^ s[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP>
void run(const char* name, int ops_per_unit, int threads, int bps) {
    const int blocks = 148 * bps;
    uint64_t* out; long long* cyc;
    cudaMalloc(&out, sizeof(uint64_t) * blocks * threads);
    cudaMalloc(&cyc, sizeof(long long) * blocks);
    probe<OP><<<blocks, threads>>>(out, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<OP><<<blocks, threads>>>(out, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    static long long h[148 * 8]; cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; i++) avg += h[i]; avg /= blocks;
    double warps_smsp = threads / 32.0 * bps / 4.0;
    double units = (double)ITER * 32;  // units per warp
    // cycles per warp-instruction per SMSP
    printf("%-36s warps/SMSP=%4.1f  %6.3f cyc per warp-unit per SMSP (%d instr/unit)  [blk cyc %.0f, %.3f ms, clk %.0f MHz]\n", name, warps_smsp,
           avg / (units * warps_smsp), ops_per_unit, avg, ms, avg / (ms * 1e3));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int pass = 0; pass < 2; pass++) {
        int th = pass == 0 ? 512 : 512, bps = pass == 0 ? 2 : 1;
        run<0>("IMAD.WIDE acc", 1, th, bps);
        run<8>("IMAD.WIDE other acc", 1, th, bps);
        run<1>("IMAD lo dep", 1, th, bps);
        run<7>("IMAD lo into hi", 1, th, bps);
        run<2>("IADD", 1, th, bps);
        run<11>("IADD.CC+ADDC", 2, th, bps);
        run<3>("WIDE + IADD", 2, th, bps);
        run<4>("IMAD lo + IADD", 2, th, bps);
        run<5>("WIDE + IMAD lo", 2, th, bps);
        run<6>("WIDE + IADD + LOP", 3, th, bps);
        run<9>("DFMA", 1, th, bps);
        run<10>("DFMA + IMAD lo", 2, th, bps);
    }
    return 0;
}
