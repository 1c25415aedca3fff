// Device-side sampling for keygen / encrypt: RingSampler over the ring
// (poly.jl:7-23) with the two distributions the schemes use --
//   uniform ring elements: every residue drawn independently and uniformly in [0, q_i)   (crt.jl:146-148, 277-279;
//                          key masks, rlwe_she.jl:156)
//   rounded Gaussian:      x = round(sigma z), z ~ N(0,1), the same integer embedded in every prime
//                          (DiscreteNormal(0, sigma): bfv.jl:31-32, ckks.jl:24-25; secrets and errors).
// The reference draws from Julia's UNSEEDED global RNG (rlwe_she.jl:169-170,197), so no bit-level parity with it
// exists; this sampler is a counter-based Philox4x32-10 keyed by (seed, stream) -- every value is a pure function
// of (seed, stream, position), so batches are reproducible and independent of the launch geometry -- and
// oracle/sampler_oracle.py restates it on the CPU bit for bit (integer path) for the parity tests.
#include "engine.h"

namespace {
struct u32x4 { u32 x, y, z, w; };
__host__ __device__ __forceinline__ u32x4 philox4x32_10(u32x4 c, u32 k0, u32 k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const u64 p0 = (u64)0xD2511F53u * c.x, p1 = (u64)0xCD9E8D57u * c.z;
        const u32x4 n = {(u32)(p1 >> 32) ^ c.y ^ k0, (u32)p1, (u32)(p0 >> 32) ^ c.w ^ k1, (u32)p0};
        c = n;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}
// counter = (index low, index high, stream, attempt); key = seed
__device__ __forceinline__ u32x4 draw(u64 seed, u32 stream, u64 index, u32 attempt) {
    const u32x4 c = {(u32)index, (u32)(index >> 32), stream, attempt};
    return philox4x32_10(c, (u32)seed, (u32)(seed >> 32));
}

// out[p][i][n] uniform in [0, q_i): 64 random bits, rejected above the largest multiple of q_i (no modulo bias)
__global__ void sample_uniform_kernel(u64* __restrict__ out, const PrimeParams* __restrict__ pp, const u32 L, const u32 logN,
                                      const u64 seed, const u32 stream, const u64 total) {
    for (u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (u64)gridDim.x * blockDim.x) {
        const PrimeConst pc = pp[(idx >> logN) % L].pc;
        const u64 limit = 0ull - (0ull - pc.q) % pc.q;      // 2^64 - (2^64 mod q): multiples of q below it are unbiased (0 = 2^64)
        u64 r;
        for (u32 attempt = 0;; attempt++) {
            const u32x4 v = draw(seed, stream, idx, attempt);
            r = ((u64)v.y << 32) | v.x;
            if (limit == 0 || r < limit) break;
            r = ((u64)v.w << 32) | v.z;
            if (r < limit) break;
        }
        out[idx] = r % pc.q;
    }
}
// out[p][i][n] = round(sigma z_(p,n)) mod q_i, z by Box-Muller from 2 x 53 random bits (round half to even)
__global__ void sample_gaussian_kernel(u64* __restrict__ out, const PrimeParams* __restrict__ pp, const u32 L, const u32 logN,
                                       const double sigma, const u64 seed, const u32 stream, const u64 total) {
    const u32 N = 1u << logN;
    for (u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (u64)gridDim.x * blockDim.x) {
        const u32 n = (u32)(idx & (N - 1));
        const u64 p = idx >> logN;
        const u32x4 v = draw(seed, stream, idx, 0);
        const u64 r0 = ((u64)v.y << 32) | v.x, r1 = ((u64)v.w << 32) | v.z;
        const double u1 = (double)((r0 >> 11) + 1) * 0x1.0p-53;     // (0, 1]
        const double u2 = (double)(r1 >> 11) * 0x1.0p-53;           // [0, 1)
        const double z = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
        const long long x = llrint(sigma * z);
        for (u32 i = 0; i < L; i++) {
            const u64 q = pp[i].pc.q;
            const u64 m = (u64)(x < 0 ? -x : x) % q;
            out[((p * L + i) << logN) + n] = (x < 0 && m) ? q - m : m;
        }
    }
}
}  // namespace

int launch_sample_uniform(tfb_ctx* c, u64 seed, u32 stream, u64* out, u64 polys, cudaStream_t st) {
    if (!polys) return TFB_OK;
    const u64 total = polys * c->L * c->N;
    const unsigned tb = 256;
    const u64 nb = (total + tb - 1) / tb;
    { ProfScope ps(PC_ELEMENTWISE, st); sample_uniform_kernel<<<(unsigned)(nb < 148ull * 32 ? nb : 148ull * 32), tb, 0, st>>>(out, c->d_pp, c->L, c->logN, seed, stream, total); }
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}
int launch_sample_gaussian(tfb_ctx* c, double sigma, u64 seed, u32 stream, u64* out, u64 polys, cudaStream_t st) {
    if (!polys) return TFB_OK;
    if (!(sigma >= 0.0) || sigma > 1e15) { tfb_set_error("sample_gaussian: sigma out of range"); return TFB_EINVAL; }
    const u64 total = polys * c->N;
    const unsigned tb = 256;
    const u64 nb = (total + tb - 1) / tb;
    { ProfScope ps(PC_ELEMENTWISE, st); sample_gaussian_kernel<<<(unsigned)(nb < 148ull * 32 ? nb : 148ull * 32), tb, 0, st>>>(out, c->d_pp, c->L, c->logN, sigma, seed, stream, total); }
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}
