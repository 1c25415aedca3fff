// pipe probe v5: is IMAD.WIDE / IMAD.HI throughput operand-dependent?  Same SASS, operand magnitudes set at run time.
// 512 threads x 1 CTA per SM (4 warps per SMSP), 8 independent chains, asm volatile forms (ptxas keeps them as written).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint64_t u64; typedef uint32_t u32;
#define ITER 1024
template <int OP>
__global__ void __launch_bounds__(512, 1) probe(u64* out, long long* cyc, u32 smask, u32 bmask) {
    u32 a = threadIdx.x * 2654435761u + 12345u;
    u32 b = ((blockIdx.x * 40503u + 7u) * 2654435761u | 1u) & bmask;
    u64 r[8]; u32 s[8], t[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { r[i] = (u64)(a * (i + 1)) * 0x9E3779B97F4A7C15ull; s[i] = ((a + i * 0x9E3779B9u) | 1u) & smask; t[i] = a ^ i; }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int rep = 0; rep < 4; rep++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                u32 lo = (((u32)r[i] ^ (u32)(r[i] >> 32)) | 1u) & smask, tl = (t[i] | 1u) & smask;
                if (OP == 0) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(r[i]) : "r"(lo), "r"(b));
                if (OP == 1) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(r[i]) : "r"(lo), "r"(b));
                if (OP == 2) asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(t[i]) : "r"(tl), "r"(b));
                if (OP == 3) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(t[i]) : "r"(tl), "r"(b));
                if (OP == 5) asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r[i]) : "r"(lo), "r"(b), "l"(r[(i + 1) & 7]));
                if (OP == 6) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(t[i]) : "r"(tl), "r"(b));
                if (OP == 7) { asm volatile("xor.b32 %0, %1, %2;" : "=r"(t[i]) : "r"(tl), "r"(b)); }
            }
        }
    }
    long long t1 = clock64();
    u64 acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= r[i] ^ t[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP>
void run(const char* name) {
    const int blocks = 148, threads = 512;
    static u64* out = nullptr; static long long* cyc = nullptr;
    if (!out) { cudaMalloc(&out, sizeof(u64) * blocks * threads); cudaMalloc(&cyc, sizeof(long long) * blocks); }
    const u32 masks[5] = {0xffu, 0xffffu, 0xffffffu, 0xfffffffu, 0xffffffffu};
    printf("%-34s", name);
    for (int sm = 0; sm < 5; sm += 2) for (int bm = 0; bm < 5; bm++) {
        for (int k = 0; k < 2; k++) { probe<OP><<<blocks, threads>>>(out, cyc, masks[sm], masks[bm]); cudaDeviceSynchronize(); }
        static long long h[148]; cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < blocks; i++) avg += h[i]; avg /= blocks;
        printf(" %5.2f", avg / ((double)ITER * 32 * 4.0));
        if (bm == 4) printf(" |");
    }
    printf("\n");
}
int main() {
    printf("cycles per warp-instr per SMSP; columns: s bits {8,24,32} x b bits {8,16,24,28,32}\n");
    run<0>("mad.wide acc in place");
    run<1>("mul.wide (RZ addend)");
    run<5>("mad.wide other addend");
    run<2>("mul.hi.u32");
    run<6>("mad.hi.u32 in place");
    run<3>("mad.lo.u32 in place");
    run<7>("LOP3 only (baseline)");
    return 0;
}
