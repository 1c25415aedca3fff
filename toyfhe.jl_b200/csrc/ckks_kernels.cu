// CKKS encoding on the device: the complex-FFT maps between N/2 complex slots and a
// real-coefficient plaintext polynomial of src/ckksencoding.jl -- the per-weight-diagonal work of the reference's
// plaintext-vector multiplies (ckksencoding.jl:106-111; 320 per pipeline in examples/encrypted_mnist).
//   encode (ckksencoding.jl:76-101): slot i goes to position (3^(i+1) mod 2N) >> 1 of a length-N vector and its
//           conjugate to ((2N - 3^(i+1)) mod 2N) >> 1 (ZmstarPermutation, :47-58); inverse DFT; twist by exp(i pi k / N)
//           ("make it negacyclic"); the real part times the scale is rounded to the nearest integer (FixedRational,
//           ckks.jl:38-44) and embedded in every prime.
//   decode (ckksencoding.jl:60-70): centred lift of every coefficient (ckks.jl:52-58), divide by the scale, twist by
//           exp(-i pi k / N), forward DFT, read the slots at (3^(i+1) mod 2N) >> 1.
// The DFT is a batched Stockham radix-2 autosort FFT in float64 (log2 N passes over global memory, twiddles from
// sincospi): the reference uses FFTW; both are floating point, so parity is the reference tests' tolerance
// (test/ckks_triv.jl, ckks_modswitch.jl: atol 1e-5 .. 1e-8), not bit-exactness -- encoded integers may differ by one
// unit in the last place where scale * x falls within rounding error of a half-integer.
#include "engine.h"

namespace {
struct cplx { double re, im; };

// one Stockham pass: p = current sub-transform length, sign = -1 forward / +1 inverse
__global__ void fft_pass_kernel(const cplx* __restrict__ x, cplx* __restrict__ y, const u32 logN, const u32 p, const double sign,
                                const u64 total) {
    const u32 half = 1u << (logN - 1);
    for (u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (u64)gridDim.x * blockDim.x) {
        const u32 i = (u32)(idx & (half - 1));
        const u64 base = (idx >> (logN - 1)) << logN;
        const u32 k = i & (p - 1);
        const cplx u0 = x[base + i], u1 = x[base + i + half];
        double s, c;
        sincospi(sign * (double)k / (double)p, &s, &c);
        const cplx t = {u1.re * c - u1.im * s, u1.re * s + u1.im * c};
        const u32 j = (i << 1) - k;
        y[base + j] = {u0.re + t.re, u0.im + t.im};
        y[base + j + p] = {u0.re - t.re, u0.im - t.im};
    }
}
// slots [polys][N/2] -> conjugate-symmetric vector [polys][N]
__global__ void ckks_scatter_kernel(const cplx* __restrict__ slots, cplx* __restrict__ v, const u32* __restrict__ pos, const u32 logN,
                                    const u64 total) {
    const u32 nh = 1u << (logN - 1), N = 1u << logN;
    for (u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (u64)gridDim.x * blockDim.x) {
        const u32 i = (u32)(idx & (nh - 1));
        const u64 p = idx >> (logN - 1);
        const cplx d = slots[idx];
        const u32 a = pos[i];                 // (3^(i+1) mod 2N) >> 1
        v[(p << logN) + a] = d;
        v[(p << logN) + (N - 1 - a)] = {d.re, -d.im};   // ((2N - 3^(i+1)) mod 2N) >> 1 = N - 1 - a
    }
}
// ifft output -> round(scale * Re(x_k e^{i pi k / N} / N)) embedded in every prime; *flag is raised when a value leaves int64
__global__ void ckks_round_kernel(const cplx* __restrict__ v, u64* __restrict__ out, const PrimeParams* __restrict__ pp, const u32 L,
                                  const u32 logN, const double scale, int* __restrict__ flag, const u64 total) {
    const u32 N = 1u << logN;
    for (u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (u64)gridDim.x * blockDim.x) {
        const u32 k = (u32)(idx & (N - 1));
        const u64 p = idx >> logN;
        const cplx z = v[idx];
        double s, c;
        sincospi((double)k / (double)N, &s, &c);
        const double re = (z.re * c - z.im * s) / (double)N;
        const double scaled = re * scale;
        if (!(fabs(scaled) < 0x1p126)) { *flag = 1; continue; }
        if (fabs(scaled) < 0x1p62) {
            const long long x = llrint(scaled);             // round half to even, like round(BigInt, .) (ckks.jl:39)
            for (u32 i = 0; i < L; i++) {
                const u64 q = pp[i].pc.q;
                const u64 m = (u64)(x < 0 ? -x : x) % q;
                out[((p * L + i) << logN) + k] = (x < 0 && m) ? q - m : m;
            }
        } else {
            // a double of this size is an integer already: |scaled| = m 2^e with a 53-bit m and e >= 9, below 2^126
            int e;
            const double fr = frexp(fabs(scaled), &e);
            const u64 m = (u64)ldexp(fr, 53);
            e -= 53;
            const u64 lo = e < 64 ? m << e : 0, hi = e < 64 ? m >> (64 - e) : m << (e - 64);
            for (u32 i = 0; i < L; i++) {
                const PrimeConst pc = pp[i].pc;
                const u64 r = red128_any(hi, lo, pc);
                out[((p * L + i) << logN) + k] = (scaled < 0 && r) ? pc.q - r : r;
            }
        }
    }
}
struct HornerQ { double q[TFB_MAX_L]; };
}  // namespace

// Garner helpers live in rns_kernels.cu; the decode front end is there too (needs garner_digits / mr_above_half)
int launch_ckks_lift(tfb_ctx* c, double scale, const u64* in, double* v, u64 polys, cudaStream_t st);

static int fft_passes(tfb_ctx* c, cplx* a, cplx* b, u64 polys, double sign, cudaStream_t st, cplx** result) {
    const u64 total = polys << (c->logN - 1);
    const unsigned tb = 256;
    const u64 nbl = (total + tb - 1) / tb;
    const unsigned nb = (unsigned)(nbl < 148ull * 32 ? nbl : 148ull * 32);
    cplx *x = a, *y = b;
    for (u32 p = 1; p < c->N; p <<= 1) {
        { ProfScope ps(PC_ELEMENTWISE, st); fft_pass_kernel<<<nb, tb, 0, st>>>(x, y, c->logN, p, sign, total); }
        cplx* t = x; x = y; y = t;
    }
    TFB_CUDA(cudaGetLastError());
    *result = x;
    return TFB_OK;
}
static int ckks_pos_table(tfb_ctx* c) {
    if (c->d_ckks_pos) return TFB_OK;
    const u32 nh = c->N / 2;
    std::vector<u32> pos(nh);
    u64 g = 1;
    for (u32 i = 0; i < nh; i++) {
        g = g * 3 % (2ull * c->N);
        pos[i] = (u32)(g >> 1);
    }
    TFB_CUDA(cudaMalloc(&c->d_ckks_pos, nh * sizeof(u32)));
    TFB_CUDA(cudaMemcpy(c->d_ckks_pos, pos.data(), nh * sizeof(u32), cudaMemcpyHostToDevice));
    return TFB_OK;
}

int launch_ckks_encode(tfb_ctx* c, double scale, const double* slots, u64* out, u64 polys, cudaStream_t st) {
    if (!polys) return TFB_OK;
    if (c->logN < 2) { tfb_set_error("ckks encode: ring degree too small"); return TFB_EUNSUPPORTED; }
    if (!(scale > 0.0) || !(scale < 1e300)) { tfb_set_error("ckks encode: scale must be positive and finite"); return TFB_EINVAL; }
    int rc = ckks_pos_table(c);
    if (rc) return rc;
    const size_t vec = (size_t)polys * c->N * sizeof(cplx);
    if ((rc = ws_reserve(c, 2 * vec + 256))) return rc;
    cplx* a = (cplx*)c->ws;
    cplx* b = a + (size_t)polys * c->N;
    int* flag = (int*)(b + (size_t)polys * c->N);
    TFB_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
    const unsigned tb = 256;
    const u64 th = polys << (c->logN - 1), tf = polys << c->logN;
    { ProfScope ps(PC_ELEMENTWISE, st); ckks_scatter_kernel<<<(unsigned)((th + tb - 1) / tb < 4736 ? (th + tb - 1) / tb : 4736), tb, 0, st>>>((const cplx*)slots, a, c->d_ckks_pos, c->logN, th); }
    cplx* res;
    if ((rc = fft_passes(c, a, b, polys, +1.0, st, &res))) return rc;
    { ProfScope ps(PC_ELEMENTWISE, st); ckks_round_kernel<<<(unsigned)((tf + tb - 1) / tb < 4736 ? (tf + tb - 1) / tb : 4736), tb, 0, st>>>(res, out, c->d_pp, c->L, c->logN, scale, flag, tf); }
    TFB_CUDA(cudaGetLastError());
    int h = 0;
    TFB_CUDA(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    TFB_CUDA(cudaStreamSynchronize(st));
    if (h) { tfb_set_error("ckks encode: |scale * coefficient| must stay below 2^126"); return TFB_EUNSUPPORTED; }
    return TFB_OK;
}

namespace {
__global__ void ckks_gather_kernel(const cplx* __restrict__ v, cplx* __restrict__ slots, const u32* __restrict__ pos, const u32 logN,
                                   const u64 total) {
    const u32 nh = 1u << (logN - 1);
    for (u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (u64)gridDim.x * blockDim.x) {
        const u32 i = (u32)(idx & (nh - 1));
        const u64 p = idx >> (logN - 1);
        slots[idx] = v[(p << logN) + pos[i]];
    }
}
}  // namespace

int launch_ckks_decode(tfb_ctx* c, double scale, const u64* in, double* slots, u64 polys, cudaStream_t st) {
    if (!polys) return TFB_OK;
    if (c->logN < 2) { tfb_set_error("ckks decode: ring degree too small"); return TFB_EUNSUPPORTED; }
    if (!(scale > 0.0) || !(scale < 1e300)) { tfb_set_error("ckks decode: scale must be positive and finite"); return TFB_EINVAL; }
    int rc = ckks_pos_table(c);
    if (rc) return rc;
    const size_t vec = (size_t)polys * c->N * sizeof(cplx);
    if ((rc = ws_reserve(c, 2 * vec))) return rc;
    cplx* a = (cplx*)c->ws;
    cplx* b = a + (size_t)polys * c->N;
    if ((rc = launch_ckks_lift(c, scale, in, (double*)a, polys, st))) return rc;
    cplx* res;
    if ((rc = fft_passes(c, a, b, polys, -1.0, st, &res))) return rc;
    const unsigned tb = 256;
    const u64 th = polys << (c->logN - 1);
    { ProfScope ps(PC_ELEMENTWISE, st); ckks_gather_kernel<<<(unsigned)((th + tb - 1) / tb < 4736 ? (th + tb - 1) / tb : 4736), tb, 0, st>>>(res, (cplx*)slots, c->d_ckks_pos, c->logN, th); }
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}
