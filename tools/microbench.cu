// Integer pipe throughput probe for sm_100a (decides the modmul instruction mix).
// Each warp runs ITER iterations of 8 independent chains of one op; reports
// thread-ops per clock per SM at full occupancy.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 2048
template <int OP>
__global__ void probe(uint64_t* out, long long* cyc) {
    uint32_t a = threadIdx.x * 2654435761u + 12345u, b = blockIdx.x * 40503u + 7u;
    uint64_t r0 = a, r1 = b, r2 = a ^ b, r3 = a + b, r4 = a * 3, r5 = b * 5, r6 = a - b, r7 = ~a;
    uint32_t s0 = a, s1 = b, s2 = a ^ b, s3 = a + b, s4 = a * 3, s5 = b * 5, s6 = a - b, s7 = ~a;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITER; i++) {
        if (OP == 0) {  // IMAD.WIDE.U32 with 64-bit accumulate
#define W(r) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(r) : "r"(a), "r"(b));
            W(r0) W(r1) W(r2) W(r3) W(r4) W(r5) W(r6) W(r7)
        } else if (OP == 1) {  // IMAD (32-bit lo)
#define M(s) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(s) : "r"(a), "r"(b));
            M(s0) M(s1) M(s2) M(s3) M(s4) M(s5) M(s6) M(s7)
        } else if (OP == 2) {  // IADD3
#define A(s) asm volatile("add.u32 %0, %0, %1;" : "+r"(s) : "r"(a));
            A(s0) A(s1) A(s2) A(s3) A(s4) A(s5) A(s6) A(s7)
        } else if (OP == 3) {  // LOP3
#define X(s) asm volatile("xor.b32 %0, %0, %1;" : "+r"(s) : "r"(a));
            X(s0) X(s1) X(s2) X(s3) X(s4) X(s5) X(s6) X(s7)
        } else if (OP == 4) {  // mix: 4 IMAD.WIDE + 4 IADD3
            W(r0) A(s0) W(r1) A(s1) W(r2) A(s2) W(r3) A(s3)
        } else if (OP == 5) {  // mul.hi.u64 (what __umul64hi expands to)
#define H(r) asm volatile("mul.hi.u64 %0, %0, %1;" : "+l"(r) : "l"(r7));
            H(r0) H(r1) H(r2) H(r3) H(r4) H(r5) H(r6) r7 += i;
        } else if (OP == 6) {  // 64-bit add (2 instr)
#define D(r) asm volatile("add.u64 %0, %0, %1;" : "+l"(r) : "l"(r7));
            D(r0) D(r1) D(r2) D(r3) D(r4) D(r5) D(r6) r7 += i;
        } else if (OP == 7) {  // IMAD.HI.U32
#define HI(s) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(s) : "r"(a), "r"(b));
            HI(s0) HI(s1) HI(s2) HI(s3) HI(s4) HI(s5) HI(s6) HI(s7)
        } else if (OP == 8) {  // mix: 4 IMAD lo + 4 IADD3
            M(s0) A(s1) M(s2) A(s3) M(s4) A(s5) M(s6) A(s7)
        } else if (OP == 9) {  // DFMA
            double d0 = __longlong_as_double(r0), d1 = __longlong_as_double(r1), d2 = __longlong_as_double(r2), d3 = __longlong_as_double(r3);
            double k = __longlong_as_double(r7);
            d0 = fma(d0, k, d0); d1 = fma(d1, k, d1); d2 = fma(d2, k, d2); d3 = fma(d3, k, d3);
            d0 = fma(d0, k, d0); d1 = fma(d1, k, d1); d2 = fma(d2, k, d2); d3 = fma(d3, k, d3);
            r0 = __double_as_longlong(d0); r1 = __double_as_longlong(d1); r2 = __double_as_longlong(d2); r3 = __double_as_longlong(d3);
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = r0 ^ r1 ^ r2 ^ r3 ^ r4 ^ r5 ^ r6 ^ r7 ^ s0 ^ s1 ^ s2 ^ s3 ^ s4 ^ s5 ^ s6 ^ s7;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP>
void run(const char* name, int ops_per_iter) {
    const int blocks = 148 * 2, threads = 512;  // 32 warps/SM
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
    long long h[148 * 2]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; i++) avg += h[i]; avg /= blocks;
    // per SM: 2 blocks x 512 threads resident together
    double ops_sm = 2.0 * threads * (double)ITER * ops_per_iter;
    printf("%-28s %8.1f thread-ops/clk/SM  (block cycles %.0f, %.3f ms, eff clk %.0f MHz)\n", name, ops_sm / avg, avg, ms,
           avg / (ms * 1e3));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>("IMAD.WIDE.U32 (+64 acc)", 8);
    run<1>("IMAD lo", 8);
    run<7>("IMAD.HI.U32", 8);
    run<2>("IADD (32-bit)", 8);
    run<3>("LOP3", 8);
    run<4>("4 IMAD.WIDE + 4 IADD", 8);
    run<8>("4 IMAD lo + 4 IADD", 8);
    run<5>("mul.hi.u64 (x7)", 7);
    run<6>("add.u64 (x7)", 7);
    run<9>("DFMA (x8)", 8);
    return 0;
}
