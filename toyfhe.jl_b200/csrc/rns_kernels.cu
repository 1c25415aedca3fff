// RNS / CRT kernels: the engine's replacement for the per-coefficient Julia code
// of crt.jl (CRTEncoded + - *, CRTExpand, modswitch), bfv.jl (switch/switchel,
// multround), rlwe_she.jl (digit decomposition, key accumulation) and
// pow2_cyc_rings.jl:321-329 (apply_galois_element).
//
// Where the reference reconstructs a BigInt per coefficient (crt.jl:105-112) the
// engine stays in word arithmetic: an exact Garner mixed-radix conversion
//     X = d_0 + d_1 q_0 + d_2 q_0 q_1 + ...,   0 <= d_i < q_i
// gives comparisons with Q/2 (centred lift, signedmod.jl:12-19), residues modulo
// any other prime, exact division by Q (round-half-away, div_hacks.jl:120-135) and
// base-2^w digits, all bit-identical to the BigInt path.
//
// Layout everywhere: u64 [..][L][N], thread index runs along N (coalesced).
#include <map>
#include <mutex>

#include "engine.h"

#define MAXD 32  // max primes in a basis that takes part in base conversion

// testing hook: force the generic runtime-L kernels even where a specialisation exists
bool g_force_generic = false;

// 128-bit accumulator helpers ------------------------------------------------
struct acc128 {
    u64 lo, hi;
};
// a += x * y for residues x, y (mac_wide62, modarith.cuh).
// Measured (profiles/r02_keyswitch_latency.txt): -14 % on the two-coefficient accumulate at large batches (C3: 0.76 -> 0.66
// ms), -10 % on the MNIST linear combinations; the one-coefficient kernels that keep 24 loads in flight per thread are
// faster with the compiler's form (mac128_plain), which ptxas interleaves with the loads more freely.
__device__ __forceinline__ void mac128_plain(acc128& a, u64 x, u64 y) {
    u64 lo = x * y, hi = __umul64hi(x, y);
    a.lo += lo;
    a.hi += hi + (a.lo < lo);
}
__device__ __forceinline__ void mac128(acc128& a, u64 x, u64 y) { mac_wide62(a.lo, a.hi, x, y); }   // modarith.cuh
// full reduction of a 128-bit value modulo q (any size of z)
__device__ __forceinline__ u64 red128_full(acc128 a, const PrimeConst& pc) {
    return red128_any(a.hi, a.lo, pc);
}

// (r / d, r % d) for a small divisor d that is loop-invariant in the caller (a number of primes, of digits): the
// reciprocal M = ceil(2^32 / d) is computed once per thread, after which a row index below 2^23 (always, in practice) costs
// one 32-bit high multiply instead of the 64-bit division routine -- the elementwise kernels below spent most of their
// instructions there (ks_finish_raised: 180 per word, add_plain: 175 per pair of words; profiles/r02_workloads.txt).
// Exact for d <= 2^9 and r < 2^23: with M d = 2^32 + e, e < d, the error r e / (d 2^32) stays below 1 / d.
struct SmallDiv {
    u32 d, M;
    __device__ __forceinline__ explicit SmallDiv(const u32 d_) : d(d_), M(d_ > 1 ? (u32)((0x100000000ull + d_ - 1) / d_) : 0) {}
    __device__ __forceinline__ u64 divmod(const u64 r, u32& rem) const {
        if (r < (1ull << 23) && d <= 512) {
            const u32 r32 = (u32)r, q = M ? __umulhi(r32, M) : r32;
            rem = r32 - q * d;
            return q;
        }
        const u64 q = r / d;
        rem = (u32)(r - q * d);
        return q;
    }
    __device__ __forceinline__ u32 mod(const u64 r) const {
        u32 rem;
        divmod(r, rem);
        return rem;
    }
};

// ------------------------------------------------------------- elementwise
template <int OP>
__global__ void binop_kernel(const ulonglong2* __restrict__ a, const ulonglong2* __restrict__ b,
                             ulonglong2* __restrict__ out, const PrimeParams* __restrict__ pp, const u32 L,
                             const u32 logN, const u64 total2) {
    const SmallDiv dl(L);
    for (u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x; idx < total2; idx += (u64)gridDim.x * blockDim.x) {
        const u64 row = idx >> (logN - 1);
        const PrimeConst pc = pp[dl.mod(row)].pc;
        const ulonglong2 x = a[idx], y = b[idx];
        ulonglong2 r;
        if (OP == 0) {
            r.x = add_mod(x.x, y.x, pc.q);
            r.y = add_mod(x.y, y.y, pc.q);
        } else if (OP == 1) {
            r.x = sub_mod(x.x, y.x, pc.q);
            r.y = sub_mod(x.y, y.y, pc.q);
        } else if (OP == 3) {   // product on 2^60 + e primes: full product + Solinas folds (modarith.cuh red126_sp60)
            const u32 e = (u32)(pc.q - (1ull << 60));
            r.x = red126_sp60(__umul64hi(x.x, y.x), x.x * y.x, pc.q, e);
            r.y = red126_sp60(__umul64hi(x.y, y.y), x.y * y.y, pc.q, e);
        } else {
            r.x = barrett_mul(x.x, y.x, pc);
            r.y = barrett_mul(x.y, y.y, pc);
        }
        out[idx] = r;
    }
}

__global__ void neg_kernel(const ulonglong2* __restrict__ a, ulonglong2* __restrict__ out,
                           const PrimeParams* __restrict__ pp, const u32 L, const u32 logN, const u64 total2) {
    const SmallDiv dl(L);
    for (u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x; idx < total2; idx += (u64)gridDim.x * blockDim.x) {
        const u64 q = pp[dl.mod(idx >> (logN - 1))].pc.q;
        const ulonglong2 x = a[idx];
        ulonglong2 r;
        r.x = neg_mod(x.x, q);
        r.y = neg_mod(x.y, q);
        out[idx] = r;
    }
}

struct ScalarArgs {
    tw_t s[TFB_MAX_L];
};

__global__ void scalar_mul_kernel(const ulonglong2* __restrict__ a, ulonglong2* __restrict__ out,
                                  const PrimeParams* __restrict__ pp, const ScalarArgs sa, const u32 L,
                                  const u32 logN, const u64 total2) {
    const SmallDiv dl(L);
    for (u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x; idx < total2; idx += (u64)gridDim.x * blockDim.x) {
        const u32 prime = dl.mod(idx >> (logN - 1));
        const u64 q = pp[prime].pc.q;
        const tw_t s = sa.s[prime];
        const ulonglong2 x = a[idx];
        ulonglong2 r;
        r.x = shoup_full(x.x, s.w, s.wp, q);
        r.y = shoup_full(x.y, s.w, s.wp, q);
        out[idx] = r;
    }
}

// Plaintext multiply, broadcast over the batch (ckksencoding.jl:106-111 `a .* c`: every component of every ciphertext
// times ONE ring element, here in the dual domain; the diagonal-method matmuls of test/ckks_matmul.jl and
// examples/encrypted_mnist accumulate these products): out[p] (+)= a[p] (.) plain, p < polys, plain [L][N] read once
// per polynomial from L2.  A plain whose rows are constant is a scalar multiply (ckksencoding.jl:100-103).
template <bool ACC, bool SP>
__global__ void mul_plain_kernel(const ulonglong2* __restrict__ a, const ulonglong2* __restrict__ plain, ulonglong2* __restrict__ out,
                                 const PrimeParams* __restrict__ pp, const u32 L, const u32 logN, const u64 total2) {
    const SmallDiv dl(L);
    const u64 low = (1ull << (logN - 1)) - 1;
    for (u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x; idx < total2; idx += (u64)gridDim.x * blockDim.x) {
        const u32 prime = dl.mod(idx >> (logN - 1));
        const u64 within = ((u64)prime << (logN - 1)) | (idx & low);
        const PrimeConst pc = pp[prime].pc;
        const ulonglong2 x = a[idx], y = plain[within];
        ulonglong2 r;
        if (SP) {
            const u32 e = (u32)(pc.q - (1ull << 60));
            r.x = red126_sp60(__umul64hi(x.x, y.x), x.x * y.x, pc.q, e);
            r.y = red126_sp60(__umul64hi(x.y, y.y), x.y * y.y, pc.q, e);
        } else {
            r.x = barrett_mul(x.x, y.x, pc);
            r.y = barrett_mul(x.y, y.y, pc);
        }
        if (ACC) {
            const ulonglong2 o = out[idx];
            r.x = add_mod(r.x, o.x, pc.q);
            r.y = add_mod(r.y, o.y, pc.q);
        }
        out[idx] = r;
    }
}

// Plaintext add, broadcast over a batch of polynomials that sit `stride2` (in 16-byte units) apart -- `c .+ b` of
// ckksencoding.jl:113-125 adds the encoded plaintext to component 1 of every ciphertext: a/out point at that component
// of the first ciphertext and the stride is the ciphertext size.  In place allowed.
__global__ void add_plain_kernel(const ulonglong2* __restrict__ a, const ulonglong2* __restrict__ plain, ulonglong2* __restrict__ out,
                                 const PrimeParams* __restrict__ pp, const u32 L, const u32 logN, const u64 stride2, const u64 total2) {
    const SmallDiv dl(L);
    const u64 low = (1ull << (logN - 1)) - 1;
    for (u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x; idx < total2; idx += (u64)gridDim.x * blockDim.x) {
        u32 prime;
        const u64 p = dl.divmod(idx >> (logN - 1), prime);
        const u64 within = ((u64)prime << (logN - 1)) | (idx & low);
        const u64 q = pp[prime].pc.q;
        const ulonglong2 x = a[p * stride2 + within], y = plain[within];
        ulonglong2 r;
        r.x = add_mod(x.x, y.x, q);
        r.y = add_mod(x.y, y.y, q);
        out[p * stride2 + within] = r;
    }
}

// Linear combinations of ciphertexts with scalar weights: out[c] = sum_j w[c][j] * in[j], c < C <= 4, j < J <= 63 -- the
// convolution of examples/encrypted_mnist/infer.jl:117-121 (sum of C_Iij * weight over the 49 window offsets, 4 channels;
// `c * b::AbstractFloat` of ckksencoding.jl:100-103 followed by `+`).  Every input element is read ONCE for all C outputs
// and the J products of an output are summed as 128-bit integers (J * q^2 < 2^128) and reduced once; the per-term
// scalar-multiply + add sequence it replaces moved 3 polynomial passes per (c, j).
// in: J buffers `in_stride` words apart, each [polys][L][N]; w: device [C][J][L] residues; out: [C][polys][L][N].
template <int C>
__global__ void lincomb_kernel(const u64* __restrict__ in, const u64 in_stride, const u32 J, const u64* __restrict__ w, u64* __restrict__ out,
                               const PrimeParams* __restrict__ pp, const u32 L, const u32 logN, const u64 total) {
    const SmallDiv dl(L);
    for (u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (u64)gridDim.x * blockDim.x) {
        const u32 prime = dl.mod(idx >> logN);
        acc128 a[C];
#pragma unroll
        for (int c = 0; c < C; c++) a[c] = {0, 0};
        for (u32 j = 0; j < J; j++) {
            const u64 x = in[(u64)j * in_stride + idx];
#pragma unroll
            for (int c = 0; c < C; c++) mac128(a[c], x, w[((u64)c * J + j) * L + prime]);
        }
        const PrimeConst pc = pp[prime].pc;
#pragma unroll
        for (int c = 0; c < C; c++) out[(u64)c * total + idx] = red128_full(a[c], pc);
    }
}

static inline unsigned grid_for(u64 work, unsigned tb) {
    u64 nb = (work + tb - 1) / tb;
    const u64 cap = 148ull * 16;
    return (unsigned)(nb < cap ? (nb ? nb : 1) : cap);
}

int launch_binop(tfb_ctx* c, int op, const u64* a, const u64* b, u64* out, u64 rows, cudaStream_t st) {
    if (!rows) return TFB_OK;
    const u64 total2 = rows * c->N / 2;
    const unsigned tb = 256, nb = grid_for(total2, tb);
    if (op == 0) { ProfScope ps(PC_ELEMENTWISE, st); binop_kernel<0><<<nb, tb, 0, st>>>((const ulonglong2*)a, (const ulonglong2*)b, (ulonglong2*)out, c->d_pp, c->L, c->logN, total2); }
    else if (op == 1) { ProfScope ps(PC_ELEMENTWISE, st); binop_kernel<1><<<nb, tb, 0, st>>>((const ulonglong2*)a, (const ulonglong2*)b, (ulonglong2*)out, c->d_pp, c->L, c->logN, total2); }
    else if (c->ntt_mode == 2 && !g_force_generic) { ProfScope ps(PC_ELEMENTWISE, st); binop_kernel<3><<<nb, tb, 0, st>>>((const ulonglong2*)a, (const ulonglong2*)b, (ulonglong2*)out, c->d_pp, c->L, c->logN, total2); }
    else { ProfScope ps(PC_ELEMENTWISE, st); binop_kernel<2><<<nb, tb, 0, st>>>((const ulonglong2*)a, (const ulonglong2*)b, (ulonglong2*)out, c->d_pp, c->L, c->logN, total2); }
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}

int launch_mul_plain(tfb_ctx* c, const u64* a, const u64* plain, u64* out, u64 polys, bool accumulate, cudaStream_t st) {
    if (!polys) return TFB_OK;
    const u64 total2 = polys * c->L * c->N / 2;
    const unsigned tb = 256, nb = grid_for(total2, tb);
    const bool sp = c->ntt_mode == 2 && !g_force_generic;
    ProfScope ps(PC_ELEMENTWISE, st);
#define MP(A, S) mul_plain_kernel<A, S><<<nb, tb, 0, st>>>((const ulonglong2*)a, (const ulonglong2*)plain, (ulonglong2*)out, c->d_pp, c->L, c->logN, total2)
    if (accumulate) { if (sp) MP(true, true); else MP(true, false); }
    else { if (sp) MP(false, true); else MP(false, false); }
#undef MP
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}

int launch_add_plain(tfb_ctx* c, const u64* a, const u64* plain, u64* out, u64 polys, u64 stride_words, cudaStream_t st) {
    if (!polys) return TFB_OK;
    const u64 total2 = polys * c->L * c->N / 2;
    const unsigned tb = 256, nb = grid_for(total2, tb);
    ProfScope ps(PC_ELEMENTWISE, st);
    add_plain_kernel<<<nb, tb, 0, st>>>((const ulonglong2*)a, (const ulonglong2*)plain, (ulonglong2*)out, c->d_pp, c->L, c->logN, stride_words / 2, total2);
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}

int launch_lincomb(tfb_ctx* c, const u64* in, u64 in_stride, u32 J, const u64* w, u32 C, u64* out, u64 polys, cudaStream_t st) {
    if (!polys) return TFB_OK;
    const u64 total = polys * c->L * c->N;
    const unsigned tb = 256, nb = grid_for(total, tb);
    ProfScope ps(PC_ELEMENTWISE, st);
    switch (C) {
        case 1: lincomb_kernel<1><<<nb, tb, 0, st>>>(in, in_stride, J, w, out, c->d_pp, c->L, c->logN, total); break;
        case 2: lincomb_kernel<2><<<nb, tb, 0, st>>>(in, in_stride, J, w, out, c->d_pp, c->L, c->logN, total); break;
        case 3: lincomb_kernel<3><<<nb, tb, 0, st>>>(in, in_stride, J, w, out, c->d_pp, c->L, c->logN, total); break;
        default: lincomb_kernel<4><<<nb, tb, 0, st>>>(in, in_stride, J, w, out, c->d_pp, c->L, c->logN, total); break;
    }
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}

int launch_neg(tfb_ctx* c, const u64* a, u64* out, u64 rows, cudaStream_t st) {
    if (!rows) return TFB_OK;
    const u64 total2 = rows * c->N / 2;
    const unsigned tb = 256, nb = grid_for(total2, tb);
    { ProfScope ps(PC_ELEMENTWISE, st); neg_kernel<<<nb, tb, 0, st>>>((const ulonglong2*)a, (ulonglong2*)out, c->d_pp, c->L, c->logN, total2); }
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}

int launch_scalar_mul(tfb_ctx* c, const u64* a, const u64* s_host, u64* out, u64 rows, cudaStream_t st) {
    if (!rows) return TFB_OK;
    ScalarArgs sa;
    for (u32 i = 0; i < c->L; i++) sa.s[i] = h_tw(s_host[i] % c->q[i], c->q[i]);
    const u64 total2 = rows * c->N / 2;
    const unsigned tb = 256, nb = grid_for(total2, tb);
    { ProfScope ps(PC_ELEMENTWISE, st); scalar_mul_kernel<<<nb, tb, 0, st>>>((const ulonglong2*)a, (ulonglong2*)out, c->d_pp, sa, c->L, c->logN, total2); }
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}

// ---------------------------------------------------- ciphertext tensor (dual)
// a,b: [B][2][L][N] (NTT domain) -> out [B][3][L][N]: d0=a0 b0, d1=a0 b1+a1 b0, d2=a1 b1
// (the coefficient-wise body of rlwe_she.jl:255-258 once everything is in dual form)
// SP: every prime is 2^60 + e (e < 2^28): products and the two-term sum are reduced by Solinas folds (red126_sp60)
template <bool SP>
__device__ __forceinline__ u64 tensor_red(const acc128 s, const PrimeConst& pc) {
    return SP ? red126_sp60(s.hi, s.lo, pc.q, (u32)(pc.q - (1ull << 60))) : red128_full(s, pc);
}
template <bool SP>
__global__ void tensor_dual_kernel(const ulonglong2* __restrict__ a, const ulonglong2* __restrict__ b,
                                   ulonglong2* __restrict__ out, const PrimeParams* __restrict__ pp, const u32 L,
                                   const u32 logN, const u64 total2) {
    const u64 rowlen = 1ull << (logN - 1);
    for (u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x; idx < total2; idx += (u64)gridDim.x * blockDim.x) {
        const u64 n2 = idx & (rowlen - 1);
        const u64 r = idx >> (logN - 1);  // (batch, prime)
        const u64 bi = r / L, pi = r % L;
        const PrimeConst pc = pp[pi].pc;
        const u64 ia0 = ((bi * 2 + 0) * L + pi) * rowlen + n2, ia1 = ((bi * 2 + 1) * L + pi) * rowlen + n2;
        const ulonglong2 a0 = a[ia0], a1 = a[ia1], b0 = b[ia0], b1 = b[ia1];
        ulonglong2 d0, d1, d2;
        acc128 s = {0, 0};
        if (SP) {
            mac128(s, a0.x, b0.x); d0.x = tensor_red<SP>(s, pc); s.lo = s.hi = 0;
            mac128(s, a0.y, b0.y); d0.y = tensor_red<SP>(s, pc); s.lo = s.hi = 0;
            mac128(s, a1.x, b1.x); d2.x = tensor_red<SP>(s, pc); s.lo = s.hi = 0;
            mac128(s, a1.y, b1.y); d2.y = tensor_red<SP>(s, pc); s.lo = s.hi = 0;
        } else {
            d0.x = barrett_mul(a0.x, b0.x, pc);
            d0.y = barrett_mul(a0.y, b0.y, pc);
            d2.x = barrett_mul(a1.x, b1.x, pc);
            d2.y = barrett_mul(a1.y, b1.y, pc);
        }
        mac128(s, a0.x, b1.x);
        mac128(s, a1.x, b0.x);
        d1.x = tensor_red<SP>(s, pc);
        s.lo = s.hi = 0;
        mac128(s, a0.y, b1.y);
        mac128(s, a1.y, b0.y);
        d1.y = tensor_red<SP>(s, pc);
        out[((bi * 3 + 0) * L + pi) * rowlen + n2] = d0;
        out[((bi * 3 + 1) * L + pi) * rowlen + n2] = d1;
        out[((bi * 3 + 2) * L + pi) * rowlen + n2] = d2;
    }
}

int launch_tensor_dual(tfb_ctx* c, const u64* a, const u64* b, u64* out, u64 batch, cudaStream_t st) {
    if (!batch) return TFB_OK;
    const u64 total2 = batch * c->L * c->N / 2;
    const unsigned tb = 256, nb = grid_for(total2, tb);
    {
        extern bool g_force_generic_red;
        ProfScope ps(PC_TENSOR, st);
        if (c->ntt_mode == 2 && !g_force_generic_red)   // every prime 2^60 + e, e < 2^28
            tensor_dual_kernel<true><<<nb, tb, 0, st>>>((const ulonglong2*)a, (const ulonglong2*)b, (ulonglong2*)out, c->d_pp, c->L, c->logN, total2);
        else
            tensor_dual_kernel<false><<<nb, tb, 0, st>>>((const ulonglong2*)a, (const ulonglong2*)b, (ulonglong2*)out, c->d_pp, c->L, c->logN, total2);
    }
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}

// ------------------------------------------------------------------- Galois
// out[(g i) mod N] = floor(g i / N) odd ? -in[i] : in[i]  (pow2_cyc_rings.jl:321-329),
// evaluated as a gather: i = g^-1 r mod 2N; i < N ? in[i] : -in[i-N].
__global__ void galois_kernel(const u64* __restrict__ in, u64* __restrict__ out, const PrimeParams* __restrict__ pp,
                              const u32 L, const u32 logN, const u32 ginv, const u64 total) {
    const u32 N = 1u << logN;
    const SmallDiv dl(L);
    for (u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (u64)gridDim.x * blockDim.x) {
        const u64 row = idx >> logN;
        const u32 r = (u32)(idx & (N - 1));
        const u64 q = pp[dl.mod(row)].pc.q;
        const u32 i = (u32)(((u64)ginv * r) & (2 * N - 1));
        const u64 v = in[(row << logN) + (i & (N - 1))];
        out[idx] = i < N ? v : neg_mod(v, q);
    }
}

// Rows of 2^10 .. 2^14 positions: the gather above reads one word out of every 32-byte sector it touches (stride g^-1), so
// the L2 -> SM traffic is four times the row; here a CTA brings its row into shared memory with full-width loads and
// gathers there (an odd word stride is bank-conflict-free for 64-bit accesses), then writes coalesced.
__global__ void __launch_bounds__(512) galois_row_kernel(const u64* __restrict__ in, u64* __restrict__ out, const PrimeParams* __restrict__ pp,
                                                         const u32 L, const u32 logN, const u32 ginv, const u64 rows) {
    extern __shared__ __align__(16) u64 srow[];
    const u32 N = 1u << logN;
    for (u64 row = blockIdx.x; row < rows; row += gridDim.x) {
        const ulonglong2* src = reinterpret_cast<const ulonglong2*>(in + (row << logN));
        for (u32 j = threadIdx.x; j < N / 2; j += blockDim.x) reinterpret_cast<ulonglong2*>(srow)[j] = src[j];
        __syncthreads();
        const u64 q = pp[row % L].pc.q;
        u64* dst = out + (row << logN);
        for (u32 r = threadIdx.x; r < N; r += blockDim.x) {
            const u32 i = (ginv * r) & (2 * N - 1);
            const u64 v = srow[i & (N - 1)];
            dst[r] = i < N ? v : neg_mod(v, q);
        }
        __syncthreads();
    }
}

int launch_galois(tfb_ctx* c, u64 g, const u64* in, u64* out, u64 rows, cudaStream_t st) {
    if (!rows) return TFB_OK;
    if (in == out) { tfb_set_error("tfb_galois cannot run in place"); return TFB_EINVAL; }
    const u64 twoN = 2ull * c->N;
    g %= twoN;
    if ((g & 1) == 0) { tfb_set_error("galois element must be odd"); return TFB_EINVAL; }
    // inverse of g modulo 2N (power of two): Newton iteration
    u64 x = g;
    for (int i = 0; i < 6; i++) x = x * (2 - g * x);
    const u32 ginv = (u32)(x & (twoN - 1));
    const u64 total = rows * c->N;
    if (!g_force_generic && c->logN >= 10 && c->logN <= 14) {
        const size_t smem = (size_t)c->N * sizeof(u64);
        static bool attr_done[64] = {};                      // per device (set once; harmless if two threads race to set it)
        if (c->device >= 0 && c->device < 64 && !attr_done[c->device]) {
            TFB_CUDA(cudaFuncSetAttribute(galois_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
            attr_done[c->device] = true;
        }
        const u64 per_sm = smem <= 32 * 1024 ? 4 : (smem <= 64 * 1024 ? 3 : 1);
        const u64 slots = (u64)(c->num_sms > 0 ? c->num_sms : 148) * per_sm;
        const unsigned nb = (unsigned)(rows < slots ? rows : slots);
        { ProfScope ps(PC_LEVEL, st); galois_row_kernel<<<nb, 512, smem, st>>>(in, out, c->d_pp, c->L, c->logN, ginv, rows); }
        TFB_CUDA(cudaGetLastError());
        return TFB_OK;
    }
    const unsigned tb = 256, nb = grid_for(total, tb);
    { ProfScope ps(PC_LEVEL, st); galois_kernel<<<nb, tb, 0, st>>>(in, out, c->d_pp, c->L, c->logN, ginv, total); }
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}

// ----------------------------------------------- rescale (modswitch, crt.jl:215-220)
// in [P][L][N] -> out [P][L-1][N]: c'_i = (q_L mod q_i)^-1 (c_i - (c_L mod q_i)), c_L un-centred
__global__ void rescale_kernel(const u64* __restrict__ in, u64* __restrict__ out, const PrimeParams* __restrict__ pp,
                               const ScalarArgs inv, const u32 L, const u32 logN, const u64 total) {
    const u32 N = 1u << logN;
    const SmallDiv dl(L - 1);
    for (u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (u64)gridDim.x * blockDim.x) {
        const u32 n = (u32)(idx & (N - 1));
        const u64 r = idx >> logN;  // (poly, i) with i < L-1
        u32 i;
        const u64 p = dl.divmod(r, i);
        const PrimeConst pc = pp[i].pc;
        const u64 ci = in[((p * L + i) << logN) + n];
        const u64 cl = barrett_red64(in[((p * L + (L - 1)) << logN) + n], pc);
        out[idx] = shoup_full(sub_mod(ci, cl, pc.q), inv.s[i].w, inv.s[i].wp, pc.q);
    }
}

int launch_rescale(tfb_ctx* c, const u64* in, u64* out, u64 polys, cudaStream_t st) {
    if (!polys) return TFB_OK;
    if (c->L < 2) { tfb_set_error("rescale needs at least two primes"); return TFB_EINVAL; }
    ScalarArgs sa;
    const u64 qL = c->q[c->L - 1];
    for (u32 i = 0; i + 1 < c->L; i++) {
        if (qL % c->q[i] == 0) { tfb_set_error("rescale: primes not coprime"); return TFB_EINVAL; }
        sa.s[i] = h_tw(h_invmod(qL % c->q[i], c->q[i]), c->q[i]);
    }
    const u64 total = polys * (c->L - 1) * c->N;
    const unsigned tb = 256, nb = grid_for(total, tb);
    { ProfScope ps(PC_LEVEL, st); rescale_kernel<<<nb, tb, 0, st>>>(in, out, c->d_pp, sa, c->L, c->logN, total); }
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}

// -------------------------------------------- CRTExpand (crt.jl:35-40): x P, append 0
__global__ void crt_expand_kernel(const u64* __restrict__ in, u64* __restrict__ out, const PrimeParams* __restrict__ pp,
                                  const ScalarArgs pm, const u32 L, const u32 logN, const u64 total) {
    const u32 N = 1u << logN;
    const SmallDiv dl(L + 1);
    for (u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (u64)gridDim.x * blockDim.x) {
        const u32 n = (u32)(idx & (N - 1));
        const u64 r = idx >> logN;  // (poly, i) with i <= L
        u32 i;
        const u64 p = dl.divmod(r, i);
        u64 v = 0;
        if (i < L) v = shoup_full(in[((p * L + i) << logN) + n], pm.s[i].w, pm.s[i].wp, pp[i].pc.q);
        out[idx] = v;
    }
}

int launch_crt_expand(tfb_ctx* c, u64 P, const u64* in, u64* out, u64 polys, cudaStream_t st) {
    if (!polys) return TFB_OK;
    ScalarArgs sa;
    for (u32 i = 0; i < c->L; i++) sa.s[i] = h_tw(P % c->q[i], c->q[i]);
    const u64 total = polys * (c->L + 1) * c->N;
    const unsigned tb = 256, nb = grid_for(total, tb);
    { ProfScope ps(PC_LEVEL, st); crt_expand_kernel<<<nb, tb, 0, st>>>(in, out, c->d_pp, sa, c->L, c->logN, total); }
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}

// ------------------------------------------------ Garner tables (per basis)
// d_gm   [L][L] : gm[i][j]  = (prod_{k<j} q_k) mod q_i      (j < i)
// d_ginv [L]    : Shoup pair of (prod_{k<i} q_k)^-1 mod q_i
// d_halfmr [L]  : mixed-radix digits of floor(Q/2)
struct GarnerTab {
    const u64* gm;
    const tw_t* ginv;
    const u64* halfmr;
};

// mixed-radix digits of X given residues r[0..L)
__device__ __forceinline__ void garner_digits(const u64* r, u64* d, const u32 L, const GarnerTab g,
                                              const PrimeParams* __restrict__ pp) {
    d[0] = r[0];
    for (u32 i = 1; i < L; i++) {
        const PrimeConst pc = pp[i].pc;
        acc128 a = {0, 0};
        for (u32 j = 0; j < i; j++) mac128(a, d[j], g.gm[i * L + j]);
        const u64 s = red128_full(a, pc);
        const tw_t iv = g.ginv[i];
        d[i] = shoup_full(sub_mod(r[i], s, pc.q), iv.w, iv.wp, pc.q);
    }
}
// X > floor(Q/2) ?
__device__ __forceinline__ bool mr_above_half(const u64* d, const u32 L, const u64* __restrict__ half) {
    for (int i = (int)L - 1; i >= 0; i--) {
        const u64 h = half[i];
        if (d[i] != h) return d[i] > h;
    }
    return false;
}

int build_garner(tfb_ctx* c) {
    const u32 L = c->L;
    std::vector<u64> gm((size_t)L * L, 0);
    std::vector<tw_t> ginv(L);
    for (u32 i = 0; i < L; i++) {
        u64 M = 1 % c->q[i];
        for (u32 j = 0; j < i; j++) {
            gm[(size_t)i * L + j] = M;
            if (c->q[j] % c->q[i] == 0) { tfb_set_error("RNS primes must be pairwise coprime"); return TFB_EINVAL; }
            M = h_mulmod(M, c->q[j] % c->q[i], c->q[i]);
        }
        ginv[i] = h_tw(i ? h_invmod(M, c->q[i]) : 1 % c->q[i], c->q[i]);
    }
    // lazy 128-bit accumulation of L products d*m needs L * qmax^2 < 2^128
    {
        long double qm = 0;
        for (u32 i = 0; i < L; i++) qm = c->q[i] > qm ? (long double)c->q[i] : qm;
        c->conv_ok = (long double)L * qm * qm < 3.4e38L;
    }
    // digits of floor(Q/2): Q-1 has digits (q_i - 1); halve from the top
    c->halfmr.assign(L, 0);
    u64 carry = 0;
    for (int i = (int)L - 1; i >= 0; i--) {
        u128 cur = (u128)carry * c->q[i] + (c->q[i] - 1);
        c->halfmr[i] = (u64)(cur / 2);
        carry = (u64)(cur % 2);
    }
    // (Q-1)/2 == floor(Q/2) only for odd Q; for even Q (q_0 = 2) floor(Q/2) = Q/2 = (Q-1)/2 + carry-handling
    bool even = false;
    for (u32 i = 0; i < L; i++) even |= (c->q[i] % 2 == 0);
    if (even) { tfb_set_error("even modulus not supported in RNS conversion"); return TFB_EUNSUPPORTED; }
    u64* d_gm = nullptr;
    TFB_CUDA(cudaMalloc(&d_gm, (size_t)L * L * sizeof(u64) + L * sizeof(u64)));
    TFB_CUDA(cudaMemcpy(d_gm, gm.data(), (size_t)L * L * sizeof(u64), cudaMemcpyHostToDevice));
    TFB_CUDA(cudaMemcpy(d_gm + (size_t)L * L, c->halfmr.data(), L * sizeof(u64), cudaMemcpyHostToDevice));
    c->d_halfmr = d_gm;  // owns the allocation: [L*L] gm then [L] halfmr
    TFB_CUDA(cudaMalloc(&c->d_ginv, L * sizeof(tw_t)));
    TFB_CUDA(cudaMemcpy(c->d_ginv, ginv.data(), L * sizeof(tw_t), cudaMemcpyHostToDevice));
    return TFB_OK;
}
static inline GarnerTab garner_of(const tfb_ctx* c) {
    GarnerTab g;
    g.gm = c->d_halfmr;
    g.halfmr = c->d_halfmr + (size_t)c->L * c->L;
    g.ginv = c->d_ginv;
    return g;
}

// ---------------------------------------------- pair tables (from basis -> to basis)
// ev   [Lt][Lf] : (prod_{k<i} qf_k) mod qt_j
// qmod [Lt]     : Qf mod qt_j
// hmod [Lt]     : floor(Qf/2) mod qt_j
// qinv [Lt]     : Shoup pair of Qf^-1 mod qt_j (zero pair if not invertible)
struct PairTab {
    u64* ev;
    u64* qmod;
    u64* hmod;
    tw_t* qinv;
};
struct PairKey {
    const tfb_ctx* a;
    const tfb_ctx* b;
    bool operator<(const PairKey& o) const { return a < o.a || (a == o.a && b < o.b); }
};
static std::map<PairKey, PairTab> g_pairs;
static std::mutex g_pairs_mu;   // contexts may be used from different host threads

void tfb_forget_ctx_pairs(const tfb_ctx* c) {
    std::lock_guard<std::mutex> lk(g_pairs_mu);
    for (auto it = g_pairs.begin(); it != g_pairs.end();) {
        if (it->first.a == c || it->first.b == c) {
            cudaFree(it->second.ev);
            cudaFree(it->second.qinv);
            it = g_pairs.erase(it);
        } else
            ++it;
    }
}

static int get_pair(const tfb_ctx* from, const tfb_ctx* to, PairTab* out) {
    std::lock_guard<std::mutex> lk(g_pairs_mu);
    PairKey k{from, to};
    auto it = g_pairs.find(k);
    if (it != g_pairs.end()) { *out = it->second; return TFB_OK; }
    const u32 Lf = from->L, Lt = to->L;
    std::vector<u64> buf((size_t)Lt * Lf + 2 * Lt);
    std::vector<tw_t> qinv(Lt);
    for (u32 j = 0; j < Lt; j++) {
        const u64 m = to->q[j];
        u64 M = 1 % m;
        for (u32 i = 0; i < Lf; i++) {
            buf[(size_t)j * Lf + i] = M;
            M = h_mulmod(M, from->q[i] % m, m);
        }
        buf[(size_t)Lt * Lf + j] = M;  // Qf mod m
        // floor(Qf/2) mod m by Horner over its mixed-radix digits
        u64 h = 0;
        for (int i = (int)Lf - 1; i >= 0; i--) h = (u64)(((u128)h * (from->q[i] % m) + from->halfmr[i] % m) % m);
        buf[(size_t)Lt * Lf + Lt + j] = h;
        tw_t z;
        z.w = z.wp = 0;
        qinv[j] = M ? h_tw(h_invmod(M, m), m) : z;
    }
    PairTab t;
    TFB_CUDA(cudaMalloc(&t.ev, buf.size() * sizeof(u64)));
    TFB_CUDA(cudaMemcpy(t.ev, buf.data(), buf.size() * sizeof(u64), cudaMemcpyHostToDevice));
    t.qmod = t.ev + (size_t)Lt * Lf;
    t.hmod = t.qmod + Lt;
    TFB_CUDA(cudaMalloc(&t.qinv, Lt * sizeof(tw_t)));
    TFB_CUDA(cudaMemcpy(t.qinv, qinv.data(), Lt * sizeof(tw_t), cudaMemcpyHostToDevice));
    g_pairs[k] = t;
    *out = t;
    return TFB_OK;
}

// residue of the mixed-radix number d (basis "from") modulo target prime j
__device__ __forceinline__ u64 mr_eval(const u64* d, const u32 Lf, const u64* __restrict__ ev_row,
                                       const PrimeConst& pc) {
    acc128 a = {0, 0};
    for (u32 i = 0; i < Lf; i++) mac128(a, d[i], ev_row[i]);
    return red128_full(a, pc);
}

// ------------------------------------ switch / switchel (bfv.jl:202-226)
// in [P][Lf][N] -> out [P][Lt][N]: centred lift from Qf (strict '>' Qf>>1), reduce into target basis
__global__ void base_switch_kernel(const u64* __restrict__ in, u64* __restrict__ out, const u32 Lf, const u32 Lt,
                                   const u32 logN, const GarnerTab g, const PrimeParams* __restrict__ ppf,
                                   const PrimeParams* __restrict__ ppt, const PairTab pt, const u64 total) {
    const u32 N = 1u << logN;
    const u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const u64 p = idx >> logN;
    const u32 n = (u32)(idx & (N - 1));
    u64 r[MAXD], d[MAXD];
    for (u32 i = 0; i < Lf; i++) r[i] = in[((p * Lf + i) << logN) + n];
    garner_digits(r, d, Lf, g, ppf);
    const bool neg = mr_above_half(d, Lf, g.halfmr);
    for (u32 j = 0; j < Lt; j++) {
        const PrimeConst pc = ppt[j].pc;
        u64 v = mr_eval(d, Lf, pt.ev + (size_t)j * Lf, pc);
        if (neg) v = sub_mod(v, pt.qmod[j], pc.q);
        out[((p * Lt + j) << logN) + n] = v;
    }
}

int launch_base_switch(tfb_ctx* from, tfb_ctx* to, const u64* in, u64* out, u64 polys, cudaStream_t st) {
    if (!polys) return TFB_OK;
    if (from->N != to->N) { tfb_set_error("base switch: ring degrees differ"); return TFB_EINVAL; }
    if (from->L > MAXD || !from->conv_ok) { tfb_set_error("base switch: basis too large for exact conversion"); return TFB_EUNSUPPORTED; }
    int rc = TFB_OK;
    if (!g_force_generic && fast_base_switch(from, to, in, out, polys, st, &rc)) return rc;
    PairTab pt;
    rc = get_pair(from, to, &pt);
    if (rc) return rc;
    const u64 total = polys * from->N;
    const unsigned tb = 128;
    const u64 nb = (total + tb - 1) / tb;
    { ProfScope ps(PC_BASE_SWITCH, st); base_switch_kernel<<<(unsigned)nb, tb, 0, st>>>(in, out, from->L, to->L, from->logN, garner_of(from), from->d_pp, to->d_pp, pt, total); }
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}

// ------------------------- mul_contract (bfv.jl:35-40, 172-190)
// in [P][Lb][N] over the big basis, out [P][L][N].  The reference's multround(SignedMod(x), t, Q) multiplies by t IN THE
// CRT FIELD (SignedMod{T}(x * T(t)), signedmod.jl:24-28), i.e. residue-wise modulo Q_big, and only then takes the
// centred lift and divides (signedmod.jl:30-32, bfv.jl:172-174):  x' = centre((t X) mod Q_big),  y = rha(x' / Q).
// Exact, all in word arithmetic:  r_j <- t r_j mod p_j;  |x'| = X' or Qb - X';  a = |x'| + (Q-1)/2;  R = a mod Q
// (known through its residues a mod q_i);  y_abs = (a - R)/Q computed modulo every
// p_j, converted back to the q_i basis;  sign restored at the end.  For odd Q,
// floor((|x'| + (Q-1)/2)/Q) is exactly round-half-away (div_hacks.jl:120-135).  y_abs <= Qb/(2Q) + 1 < Qb for any t.
struct ContractArgs {
    tw_t t_b[MAXD];  // t mod p_j
};
__global__ void bfv_contract_kernel(const u64* __restrict__ in, u64* __restrict__ out, const u32 L, const u32 Lb,
                                    const u32 logN, const GarnerTab gq, const GarnerTab gb,
                                    const PrimeParams* __restrict__ ppq, const PrimeParams* __restrict__ ppb,
                                    const PairTab b2q, const PairTab q2b, const u64* __restrict__ hq,
                                    const ContractArgs ca, const u64 total) {
    const u32 N = 1u << logN;
    const u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const u64 p = idx >> logN;
    const u32 n = (u32)(idx & (N - 1));
    u64 rb[MAXD], db[MAXD], aq[MAXD], dq[MAXD];
    for (u32 j = 0; j < Lb; j++)   // e.x * T(t): the product in the field, before any lift
        rb[j] = shoup_full(in[((p * Lb + j) << logN) + n], ca.t_b[j].w, ca.t_b[j].wp, ppb[j].pc.q);
    garner_digits(rb, db, Lb, gb, ppb);
    const bool neg = mr_above_half(db, Lb, gb.halfmr);
    // a mod q_i = |x'| + h
    for (u32 i = 0; i < L; i++) {
        const PrimeConst pc = ppq[i].pc;
        u64 v = mr_eval(db, Lb, b2q.ev + (size_t)i * Lb, pc);  // X' mod q_i
        if (neg) v = sub_mod(b2q.qmod[i], v, pc.q);             // (Qb - X') mod q_i
        aq[i] = add_mod(v, hq[i], pc.q);
    }
    // R = a mod Q, mixed radix over the q basis
    garner_digits(aq, dq, L, gq, ppq);
    // y_abs mod p_j = (|x'| + h - R) Q^-1
    for (u32 j = 0; j < Lb; j++) {
        const PrimeConst pc = ppb[j].pc;
        const u64 xa = neg ? neg_mod(rb[j], pc.q) : rb[j];
        const u64 a = add_mod(xa, q2b.hmod[j], pc.q);
        const u64 Rj = mr_eval(dq, L, q2b.ev + (size_t)j * L, pc);
        rb[j] = shoup_full(sub_mod(a, Rj, pc.q), q2b.qinv[j].w, q2b.qinv[j].wp, pc.q);
    }
    // y_abs back to the q basis
    garner_digits(rb, db, Lb, gb, ppb);
    for (u32 i = 0; i < L; i++) {
        const PrimeConst pc = ppq[i].pc;
        const u64 v = mr_eval(db, Lb, b2q.ev + (size_t)i * Lb, pc);
        out[((p * L + i) << logN) + n] = neg ? neg_mod(v, pc.q) : v;
    }
}

int launch_bfv_contract(tfb_ctx* cq, tfb_ctx* cb, u64 t, const u64* in, u64* out, u64 polys, cudaStream_t st) {
    if (!polys) return TFB_OK;
    if (cq->N != cb->N) { tfb_set_error("bfv contract: ring degrees differ"); return TFB_EINVAL; }
    if (cq->L > MAXD || cb->L > MAXD || !cq->conv_ok || !cb->conv_ok) { tfb_set_error("bfv contract: basis too large for exact conversion"); return TFB_EUNSUPPORTED; }
    for (u32 i = 0; i < cq->L; i++)
        for (u32 j = 0; j < cb->L; j++)
            if (cq->q[i] == cb->q[j]) { tfb_set_error("bfv contract: the two bases must be disjoint"); return TFB_EINVAL; }
    int rc = TFB_OK;
    if (!g_force_generic && fast_bfv_contract(cq, cb, t, in, out, polys, st, &rc)) return rc;
    PairTab b2q, q2b, q2q;
    rc = get_pair(cb, cq, &b2q);
    if (rc) return rc;
    if ((rc = get_pair(cq, cb, &q2b))) return rc;
    if ((rc = get_pair(cq, cq, &q2q))) return rc;
    ContractArgs ca;
    for (u32 j = 0; j < cb->L; j++) ca.t_b[j] = h_tw(t % cb->q[j], cb->q[j]);
    const u64 total = polys * cq->N;
    const unsigned tb = 128;
    const u64 nb = (total + tb - 1) / tb;
    { ProfScope ps(PC_BFV_CONTRACT, st); bfv_contract_kernel<<<(unsigned)nb, tb, 0, st>>>(in, out, cq->L, cb->L, cq->logN, garner_of(cq), garner_of(cb),
                                                     cq->d_pp, cb->d_pp, b2q, q2b, q2q.hmod, ca, total); }
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}

// ---------------------------------------- keyswitch digit polys (rlwe_she.jl:326-338)
// cend: last component of each ciphertext, rows [B][stride_polys][L][N] starting at the
// component (the caller passes the pointer to component comps-1 and the per-ciphertext
// stride in words).  out [B][Dn][Lt][N] holds digits k0 .. k0+Dn-1.
// w == 0: CRT digits -- digit i = centred residue i re-embedded in every target prime (:329)
__global__ void ks_digits_crt_kernel(const u64* __restrict__ cend, const u64 ct_stride, u64* __restrict__ out,
                                     const u32 L, const u32 Lt, const u32 logN, const u32 k0, const u32 Dn,
                                     const PrimeParams* __restrict__ ppq, const PrimeParams* __restrict__ ppt,
                                     const u64 total) {
    const u32 N = 1u << logN;
    for (u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (u64)gridDim.x * blockDim.x) {
        const u32 n = (u32)(idx & (N - 1));
        const u64 r = idx >> logN;  // (b, kk)
        const u64 b = r / Dn;
        const u32 i = k0 + (u32)(r % Dn);
        const u64 qi = ppq[i].pc.q;
        const u64 c = cend[b * ct_stride + ((u64)i << logN) + n];
        const bool neg = c > (qi >> 1);
        const u64 mag = neg ? qi - c : c;
        for (u32 j = 0; j < Lt; j++) {
            const PrimeConst pc = ppt[j].pc;
            const u64 v = (qi >> 1) < pc.q ? mag : barrett_red64(mag, pc);   // |digit| <= q_i / 2 (uniform branch)
            out[((r * Lt + j) << logN) + n] = neg ? neg_mod(v, pc.q) : v;
        }
    }
}
// w > 0: base-2^w digits of the un-centred integer X in [0,Q) (:331-337)
__global__ void ks_digits_pow2_kernel(const u64* __restrict__ cend, const u64 ct_stride, u64* __restrict__ out,
                                      const u32 L, const u32 Lt, const u32 logN, const u32 w, const u32 k0,
                                      const u32 Dn, const GarnerTab g, const PrimeParams* __restrict__ ppq,
                                      const PrimeParams* __restrict__ ppt, const u64 total) {
    const u32 N = 1u << logN;
    const u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const u64 b = idx >> logN;
    const u32 n = (u32)(idx & (N - 1));
    u64 r[MAXD], d[MAXD], X[MAXD + 1];
    for (u32 i = 0; i < L; i++) r[i] = cend[b * ct_stride + ((u64)i << logN) + n];
    garner_digits(r, d, L, g, ppq);
    // binary limbs by Horner: X = (..(d_{L-1} q_{L-2} + d_{L-2}) q_{L-3} + ..) q_0 + d_0
    u32 nl = 1;
    X[0] = d[L - 1];
    for (int i = (int)L - 2; i >= 0; i--) {
        const u64 m = ppq[i].pc.q;
        u64 carry = d[i];
        for (u32 k = 0; k < nl; k++) {
            const u64 lo = X[k] * m, hi = __umul64hi(X[k], m);
            const u64 s = lo + carry;
            X[k] = s;
            carry = hi + (s < lo);
        }
        X[nl++] = carry;
    }
    const u64 mask = w >= 64 ? ~0ull : ((1ull << w) - 1);
    // blockIdx.y splits the digit range: small batches have too few coefficients to fill the GPU, so several CTAs
    // repeat the (cheap) reconstruction of a coefficient and each writes its own slice of the digits
    const u32 per = (Dn + gridDim.y - 1) / gridDim.y;
    const u32 kbeg = blockIdx.y * per, kend = kbeg + per < Dn ? kbeg + per : Dn;
    for (u32 kk = kbeg; kk < kend; kk++) {
        const u32 bit = (k0 + kk) * w, limb = bit >> 6, off = bit & 63;
        u64 v = limb < nl ? X[limb] >> off : 0;
        if (off + w > 64 && limb + 1 < nl) v |= X[limb + 1] << (64 - off);
        v &= mask;
        if (Lt == 0) {   // compact: one row per digit (the digit is below every prime; launch_ntt_bcast transforms it under all of them)
            out[((b * Dn + kk) << logN) + n] = v;
            continue;
        }
        for (u32 j = 0; j < Lt; j++) {
            const PrimeConst pc = ppt[j].pc;
            out[(((b * Dn + kk) * Lt + j) << logN) + n] = v < pc.q ? v : barrett_red64(v, pc);
        }
    }
}

// The integers X in [0,Q) themselves, as L binary limbs each, limb-major: limbs [B][L][N].  One Garner conversion per
// coefficient, done once; the forward transform then cuts each base-2^w digit out of a limb row while it loads it
// (ntt_core3.cuh pass1_pow2), so neither the digit rows nor the per-slice repeats of this reconstruction exist.
__global__ void ks_limbs_kernel(const u64* __restrict__ cend, const u64 ct_stride, u64* __restrict__ limbs, const u32 L,
                                const u32 logN, const GarnerTab g, const PrimeParams* __restrict__ ppq, const u64 total) {
    const u32 N = 1u << logN;
    const u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const u64 b = idx >> logN;
    const u32 n = (u32)(idx & (N - 1));
    u64 r[MAXD], d[MAXD], X[MAXD + 1];
    for (u32 i = 0; i < L; i++) r[i] = cend[b * ct_stride + ((u64)i << logN) + n];
    garner_digits(r, d, L, g, ppq);
    u32 nl = 1;
    X[0] = d[L - 1];
    for (int i = (int)L - 2; i >= 0; i--) {          // Horner over the mixed-radix digits, as in ks_digits_pow2_kernel
        const u64 m = ppq[i].pc.q;
        u64 carry = d[i];
        for (u32 k = 0; k < nl; k++) {
            const u64 lo = X[k] * m, hi = __umul64hi(X[k], m);
            const u64 s = lo + carry;
            X[k] = s;
            carry = hi + (s < lo);
        }
        X[nl++] = carry;
    }
    for (u32 k = 0; k < L; k++) limbs[((b * L + k) << logN) + n] = X[k];
}
int launch_ks_limbs(tfb_ctx* c, const u64* cend, u64 ct_stride, u64* limbs, u64 batch, cudaStream_t st) {
    if (!batch) return TFB_OK;
    if (c->L > MAXD || !c->conv_ok) { tfb_set_error("keyswitch digits: basis too large for exact conversion"); return TFB_EUNSUPPORTED; }
    const u64 total = batch * c->N;
    const unsigned tb = 64;
    { ProfScope ps(PC_KS_DIGITS, st); ks_limbs_kernel<<<(unsigned)((total + tb - 1) / tb), tb, 0, st>>>(cend, ct_stride, limbs, c->L, c->logN, garner_of(c), c->d_pp, total); }
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}

int launch_ks_digits(tfb_ctx* c, tfb_ctx* target, int w, const u64* cend, u64 ct_stride, u64* out, u32 k0, u32 Dn,
                     u64 batch, cudaStream_t st, bool compact) {
    if (!batch || !Dn) return TFB_OK;
    if (c->N != target->N) { tfb_set_error("keyswitch digits: ring degrees differ"); return TFB_EINVAL; }
    if (w < 0 || w > 63) { tfb_set_error("keyswitch digits: relin_window must be in 0..63"); return TFB_EINVAL; }
    if (compact && w == 0) { tfb_set_error("internal: CRT digits have no compact form"); return TFB_EINVAL; }
    if (w == 0) {
        if (k0 + Dn > c->L) { tfb_set_error("keyswitch digits: digit range out of bounds"); return TFB_EINVAL; }
        const u64 total = batch * Dn * c->N;
        const unsigned tb = 256, nb = grid_for(total, tb);
        { ProfScope ps(PC_KS_DIGITS, st); ks_digits_crt_kernel<<<nb, tb, 0, st>>>(cend, ct_stride, out, c->L, target->L, c->logN, k0, Dn, c->d_pp, target->d_pp, total); }
    } else {
        if (c->L > MAXD || !c->conv_ok) { tfb_set_error("keyswitch digits: basis too large for exact conversion"); return TFB_EUNSUPPORTED; }
        const u64 total = batch * c->N;
        const unsigned tb = 128;
        const u64 nb = (total + tb - 1) / tb;
        u64 chunks = nb < 2368 ? 2368 / nb : 1;             // aim at 16 CTAs per SM
        if (chunks > (Dn + 7) / 8) chunks = (Dn + 7) / 8;   // at least 8 digits per slice
        if (chunks < 1) chunks = 1;
        { ProfScope ps(PC_KS_DIGITS, st); ks_digits_pow2_kernel<<<dim3((unsigned)nb, (unsigned)chunks), tb, 0, st>>>(cend, ct_stride, out, c->L, compact ? 0 : target->L, c->logN, (u32)w, k0, Dn,
                                                          garner_of(c), c->d_pp, target->d_pp, total); }
    }
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}

// --------------------- key accumulation in the NTT domain (rlwe_she.jl:340-344)
// acc [B][2][L][N] (dual): acc[b][0] (+)= sum_k masked_k . p_k ; acc[b][1] (+)= sum_k mask_k . p_k
// dig [B][Dn][L][N] dual; key [D][2][L][N] dual, component 0 = mask, 1 = masked; uses
// key digits k0 .. k0+Dn-1.
// The sums are kept as 128-bit integers and reduced only when `cap` more products could overflow (cap =
// floor(2^128 / qmax^2) - 1: 254 terms for 60-bit primes); the digit loop runs in groups of 4 with all 12 loads of a
// group issued before its multiply-accumulates (the kernel streams the key and the digit rows once: memory-level
// parallelism, not arithmetic, bounds it); each thread owns two adjacent coefficients so every access is 128 bits wide.
__global__ void ks_accum_kernel(const ulonglong2* __restrict__ dig, const ulonglong2* __restrict__ key, ulonglong2* __restrict__ acc,
                                const u32 L, const u32 logN, const u32 k0, const u32 Dn, const int accumulate,
                                const PrimeParams* __restrict__ pp, const u64 total2, const u32 cap) {
    constexpr u32 U = 4;
    const u32 lg2 = logN - 1;                    // rows in units of two coefficients (128-bit accesses)
    const u64 kstride = (u64)L << lg2;           // between consecutive digit rows of one (b, i) and between key components
    // CTAs that run together work on the SAME key rows for different ciphertexts (b varies fastest over the tiles), so
    // the key -- larger than L2 at base 4 -- comes from HBM once per batch instead of once per ciphertext
    const u64 tiles = total2 / blockDim.x;                 // launch_ks_accum guarantees blockDim.x | N/2
    const u64 B = total2 / ((u64)L << lg2), cpr = (1ull << lg2) / blockDim.x;
    for (u64 tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const u64 b = tile % B, rest = tile / B;
        const u64 n2 = (rest % cpr) * blockDim.x + threadIdx.x;
        const u32 i = (u32)(rest / cpr);
        const PrimeConst pc = pp[i].pc;
        acc128 a1x = {0, 0}, a1y = {0, 0}, a2x = {0, 0}, a2y = {0, 0};
        ulonglong2* o1 = acc + (((b * 2 + 0) * L + i) << lg2) + n2;
        ulonglong2* o2 = acc + (((b * 2 + 1) * L + i) << lg2) + n2;
        if (accumulate) {
            const ulonglong2 v1 = *o1, v2 = *o2;
            a1x.lo = v1.x; a1y.lo = v1.y; a2x.lo = v2.x; a2y.lo = v2.y;
        }
        const ulonglong2* dp = dig + (((b * Dn) * L + i) << lg2) + n2;
        const ulonglong2* kp = key + ((((u64)k0 * 2) * L + i) << lg2) + n2;
        u32 pending = 0, kk = 0;
#define KS_REDUCE()                                                                     \
        do {                                                                            \
            a1x.lo = red128_full(a1x, pc); a1x.hi = 0; a1y.lo = red128_full(a1y, pc); a1y.hi = 0; \
            a2x.lo = red128_full(a2x, pc); a2x.hi = 0; a2y.lo = red128_full(a2y, pc); a2y.hi = 0; \
            pending = 0;                                                                \
        } while (0)
        for (; kk + U <= Dn; kk += U) {
            if (pending + U > cap) KS_REDUCE();
            ulonglong2 p[U], km[U], kd[U];
#pragma unroll
            for (u32 u = 0; u < U; u++) {
                p[u] = dp[(u64)(kk + u) * kstride];
                km[u] = kp[(u64)(kk + u) * 2 * kstride];
                kd[u] = kp[(u64)(kk + u) * 2 * kstride + kstride];
            }
#pragma unroll
            for (u32 u = 0; u < U; u++) {
                mac128(a1x, kd[u].x, p[u].x); mac128(a1y, kd[u].y, p[u].y);
                mac128(a2x, km[u].x, p[u].x); mac128(a2y, km[u].y, p[u].y);
            }
            pending += U;
        }
        for (; kk < Dn; kk++) {
            if (pending + 1 > cap) KS_REDUCE();
            const ulonglong2 p = dp[(u64)kk * kstride];
            const ulonglong2 kd = kp[(u64)kk * 2 * kstride + kstride], km = kp[(u64)kk * 2 * kstride];
            mac128(a1x, kd.x, p.x); mac128(a1y, kd.y, p.y);
            mac128(a2x, km.x, p.x); mac128(a2y, km.y, p.y);
            pending++;
        }
#undef KS_REDUCE
        *o1 = make_ulonglong2(red128_full(a1x, pc), red128_full(a1y, pc));
        *o2 = make_ulonglong2(red128_full(a2x, pc), red128_full(a2y, pc));
    }
}

// scalar variant (one coefficient per thread, groups of 8): more threads, used when the batch is small
__global__ void ks_accum1_kernel(const u64* __restrict__ dig, const u64* __restrict__ key, u64* __restrict__ acc,
                                 const u32 L, const u32 logN, const u32 k0, const u32 Dn, const int accumulate,
                                 const PrimeParams* __restrict__ pp, const u64 total, const u32 cap) {
    constexpr u32 U = 8;
    const u32 N = 1u << logN;
    const u64 kstride = (u64)L << logN;
    for (u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (u64)gridDim.x * blockDim.x) {
        const u32 n = (u32)(idx & (N - 1));
        const u64 r = idx >> logN;  // (b, i)
        const u64 b = r / L;
        const u32 i = (u32)(r % L);
        const PrimeConst pc = pp[i].pc;
        acc128 a1 = {0, 0}, a2 = {0, 0};
        u64* o1 = acc + (((b * 2 + 0) * L + i) << logN) + n;
        u64* o2 = acc + (((b * 2 + 1) * L + i) << logN) + n;
        if (accumulate) {
            a1.lo = *o1;
            a2.lo = *o2;
        }
        const u64* dp = dig + (((b * Dn) * L + i) << logN) + n;
        const u64* kp = key + ((((u64)k0 * 2) * L + i) << logN) + n;
        u32 pending = 0, kk = 0;
        for (; kk + U <= Dn; kk += U) {
            if (pending + U > cap) {
                a1.lo = red128_full(a1, pc); a1.hi = 0;
                a2.lo = red128_full(a2, pc); a2.hi = 0;
                pending = 0;
            }
            u64 p[U], km[U], kd[U];
#pragma unroll
            for (u32 u = 0; u < U; u++) {
                p[u] = dp[(u64)(kk + u) * kstride];
                km[u] = kp[(u64)(kk + u) * 2 * kstride];
                kd[u] = kp[(u64)(kk + u) * 2 * kstride + kstride];
            }
#pragma unroll
            for (u32 u = 0; u < U; u++) {
                mac128_plain(a1, kd[u], p[u]);
                mac128_plain(a2, km[u], p[u]);
            }
            pending += U;
        }
        for (; kk < Dn; kk++) {
            if (pending + 1 > cap) {
                a1.lo = red128_full(a1, pc); a1.hi = 0;
                a2.lo = red128_full(a2, pc); a2.hi = 0;
                pending = 0;
            }
            const u64 p = dp[(u64)kk * kstride];
            mac128_plain(a1, kp[(u64)kk * 2 * kstride + kstride], p);
            mac128_plain(a2, kp[(u64)kk * 2 * kstride], p);
            pending++;
        }
        *o1 = red128_full(a1, pc);
        *o2 = red128_full(a2, pc);
    }
}

// Few coefficients (one ciphertext, or one residue shard of it): one thread per coefficient leaves the GPU nearly empty and
// every thread walks all Dn digit rows one dependent group of loads after the other (101 us for 95 MB at N = 2^14, one
// prime, D = 241).  Here blockDim = (32, S): warp y sums the digits kk = y, y + S, .. of 32 adjacent coefficients and the S
// partial sums of a coefficient meet in shared memory, so S times as many loads are in flight.
template <int S>
__global__ void __launch_bounds__(32 * S)
ks_accum_split_kernel(const u64* __restrict__ dig, const u64* __restrict__ key, u64* __restrict__ acc, const u32 L,
                      const u32 logN, const u32 k0, const u32 Dn, const int accumulate, const PrimeParams* __restrict__ pp,
                      const u32 cap) {
    constexpr u32 U = 4;
    __shared__ u64 part[S][2][32];
    const u32 N = 1u << logN, x = threadIdx.x, y = threadIdx.y;
    const u64 kstride = (u64)L << logN;
    const u64 idx = (u64)blockIdx.x * 32 + x;      // the launch covers B L N / 32 blocks exactly (32 | N)
    const u32 n = (u32)(idx & (N - 1));
    const u64 r = idx >> logN;  // (b, i)
    const u64 b = r / L;
    const u32 i = (u32)(r % L);
    const PrimeConst pc = pp[i].pc;
    acc128 a1 = {0, 0}, a2 = {0, 0};
    const u64* dp = dig + (((b * Dn) * L + i) << logN) + n;
    const u64* kp = key + ((((u64)k0 * 2) * L + i) << logN) + n;
    u32 pending = 0, kk = y;
    for (; kk + (U - 1) * S < Dn; kk += U * S) {
        if (pending + U > cap) {
            a1.lo = red128_full(a1, pc); a1.hi = 0;
            a2.lo = red128_full(a2, pc); a2.hi = 0;
            pending = 0;
        }
        u64 p[U], km[U], kd[U];
#pragma unroll
        for (u32 u = 0; u < U; u++) {
            p[u] = dp[(u64)(kk + u * S) * kstride];
            km[u] = kp[(u64)(kk + u * S) * 2 * kstride];
            kd[u] = kp[(u64)(kk + u * S) * 2 * kstride + kstride];
        }
#pragma unroll
        for (u32 u = 0; u < U; u++) {
            mac128_plain(a1, kd[u], p[u]);
            mac128_plain(a2, km[u], p[u]);
        }
        pending += U;
    }
    for (; kk < Dn; kk += S) {
        if (pending + 1 > cap) {
            a1.lo = red128_full(a1, pc); a1.hi = 0;
            a2.lo = red128_full(a2, pc); a2.hi = 0;
            pending = 0;
        }
        const u64 p = dp[(u64)kk * kstride];
        mac128_plain(a1, kp[(u64)kk * 2 * kstride + kstride], p);
        mac128_plain(a2, kp[(u64)kk * 2 * kstride], p);
        pending++;
    }
    part[y][0][x] = red128_full(a1, pc);
    part[y][1][x] = red128_full(a2, pc);
    __syncthreads();
    if (y < 2) {                                   // warp y finishes component y of its 32 coefficients
        u64* o = acc + (((b * 2 + y) * L + i) << logN) + n;
        acc128 t = {accumulate ? *o : 0, 0};
#pragma unroll
        for (u32 s = 0; s < (u32)S; s++) {
            const u64 v = part[s][y][x];
            t.lo += v;
            t.hi += t.lo < v;
        }
        *o = red128_full(t, pc);
    }
}

int launch_ks_accum(tfb_ctx* c, u32 k0, u32 Dn, const u64* dig, const u64* key, u64* acc, int accumulate, u64 batch,
                    cudaStream_t st) {
    if (!batch) return TFB_OK;
    long double qm = 0;
    for (u32 i = 0; i < c->L; i++) qm = (long double)c->q[i] > qm ? (long double)c->q[i] : qm;
    // products of canonical residues are below qmax^2; one slot is left for the carried (reduced) value
    long double terms = 3.402823669209384634e38L / (qm * qm);
    const u32 cap = terms >= 1e6L ? 1000000u : (terms >= 3.0L ? (u32)terms - 1 : 1u);
    if (c->logN < 1) { tfb_set_error("keyswitch: ring degree too small"); return TFB_EUNSUPPORTED; }
    const u64 total2 = batch * c->L * c->N / 2;
    const unsigned tb = 256;
    ProfScope ps(PC_KS_ACCUM, st);
    const u64 total = 2 * total2;
    // split geometry measured on B200 at N = 2^14, D = 241 (profiles/r02_keyswitch_latency.txt): it pays
    // for one ciphertext over a few primes (95 -> 23 us at one prime, 117 -> 80 us at four) and not beyond
    if (!g_force_generic && c->N % 32 == 0 && Dn >= 64 && total <= 100000) {
        const unsigned nb = (unsigned)(total / 32);
        if (total <= 40000) ks_accum_split_kernel<8><<<nb, dim3(32, 8), 0, st>>>(dig, key, acc, c->L, c->logN, k0, Dn, accumulate, c->d_pp, cap);
        else ks_accum_split_kernel<4><<<nb, dim3(32, 4), 0, st>>>(dig, key, acc, c->L, c->logN, k0, Dn, accumulate, c->d_pp, cap);
    } else if (total2 < 100000 || (c->N / 2) % tb != 0) {   // small batches: one coefficient per thread keeps more loads in flight (with the PTX mac128 the two-coefficient kernel wins from 128 K pairs on: 217 -> 175 us, profiles/r02_keyswitch_latency.txt)
        ks_accum1_kernel<<<grid_for(2 * total2, tb), tb, 0, st>>>(dig, key, acc, c->L, c->logN, k0, Dn, accumulate, c->d_pp, 2 * total2, cap);
    } else {
        ks_accum_kernel<<<grid_for(total2, tb), tb, 0, st>>>((const ulonglong2*)dig, (const ulonglong2*)key, (ulonglong2*)acc, c->L, c->logN, k0, Dn, accumulate, c->d_pp, total2, cap);
    }
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}

// ---------------- keyswitch epilogue: add the switched part onto (c1, c2)
// plain params:      out[b][k] = (k < comps-1 ? ct[b][k] : 0) + acc[b][k]            (rlwe_she.jl:323-324,343-346)
// ModulusRaised:     v = P * ct[b][k] (+0 on the special row) + acc_ext[b][k];  out = modswitch(v)
//                    (modulusraising.jl:35-42 with crt.jl:215-220)
// (shards: acc/out cover the primes first..first+L-1 of a ciphertext with Lct primes, pp = the shard's primes)
__global__ void ks_finish_kernel(const u64* __restrict__ ct, const u32 comps, const u64* __restrict__ acc,
                                 u64* __restrict__ out, const u32 L, const u32 logN,
                                 const PrimeParams* __restrict__ pp, const u64 total, const u32 Lct, const u32 first) {
    const u32 N = 1u << logN;
    const SmallDiv dl(L);
    for (u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (u64)gridDim.x * blockDim.x) {
        const u32 n = (u32)(idx & (N - 1));
        const u64 r = idx >> logN;  // (b, k, i)
        u32 i;
        const u64 bk = dl.divmod(r, i);
        const u32 k = (u32)(bk & 1);
        const u64 b = bk >> 1;
        const u64 q = pp[i].pc.q;
        u64 v = acc[idx];
        if (k + 1 < comps) v = add_mod(v, ct[(((b * comps + k) * Lct + first + i) << logN) + n], q);
        out[idx] = v;
    }
}
// Sharded keyswitch epilogue fused with the exchange of the result rows (BASELINE config 4, residues sharded over GPUs):
// the value ks_finish_kernel would write is stored straight into EVERY rank's result buffer over NVLink peer memory (this
// rank's rows of [B][2][Lct][N]), then the last CTA to finish publishes `epoch` in every peer's flag word and waits until
// all peers have published theirs -- when the kernel ends, this rank's buffer holds all rows.  One launch replaces
// ks_finish + all-gather + reassembly copy.  Ordering: data stores, fence.sys, CTA counter (device atomics), fence.sys, flag
// stores (release.sys) on the writer; acquire.sys flag loads on the reader.
struct PeerTab {
    u64* out[TFB_MAX_PEERS];     // rank p's result buffer of this epoch
    u64* flag[TFB_MAX_PEERS];    // rank p's flag words [world]; word r is written by rank r only
    u32* count;                  // local CTA counter (zero between launches)
    u32* err;                    // local: set when the wait timed out
    u32 rank, world;
};
__device__ __forceinline__ void st_release_sys(u64* p, u64 v) { asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ u64 ld_acquire_sys(const u64* p) {
    u64 v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ u64 globaltimer_ns() {
    u64 t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__global__ void ks_finish_push_kernel(const u64* __restrict__ ct, const u32 comps, const u64* __restrict__ acc, const u32 L,
                                      const u32 logN, const PrimeParams* __restrict__ pp, const u64 total, const u32 Lct,
                                      const u32 first, const PeerTab pt, const u64 epoch) {
    const u32 N = 1u << logN;
    const SmallDiv dl(L);
    for (u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (u64)gridDim.x * blockDim.x) {
        const u32 n = (u32)(idx & (N - 1));
        const u64 r = idx >> logN;  // (b, k, i)
        u32 i;
        const u64 bk = dl.divmod(r, i);
        const u32 k = (u32)(bk & 1);
        const u64 b = bk >> 1;
        const u64 q = pp[i].pc.q;
        u64 v = acc[idx];
        if (k + 1 < comps) v = add_mod(v, ct[(((b * comps + k) * Lct + first + i) << logN) + n], q);
        const u64 o = (((bk * Lct) + first + i) << logN) + n;
        for (u32 p = 0; p < pt.world; p++) pt.out[p][o] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const u32 prev = atomicAdd(pt.count, 1u);
        if (prev == gridDim.x - 1) {                 // every CTA's stores are fenced before its increment
            *pt.count = 0;
            __threadfence_system();
            for (u32 p = 0; p < pt.world; p++) st_release_sys(pt.flag[p] + pt.rank, epoch);
            const u64 t0 = globaltimer_ns();
            for (u32 p = 0; p < pt.world; p++)
                while (ld_acquire_sys(pt.flag[pt.rank] + p) < epoch)
                    if (globaltimer_ns() - t0 > 2000000000ull) { *pt.err = 1; return; }   // a peer never arrived: report, do not hang
        }
    }
}

struct RaiseArgs {
    tw_t inv[TFB_MAX_L];  // (P mod q_i)^-1
};
__global__ void ks_finish_raised_kernel(const u64* __restrict__ ct, const u32 comps, const u64* __restrict__ acc,
                                        u64* __restrict__ out, const u32 l, const u32 logN,
                                        const PrimeParams* __restrict__ pp, const RaiseArgs ra, const u64 total) {
    const u32 N = 1u << logN;
    const SmallDiv dl(l);
    for (u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (u64)gridDim.x * blockDim.x) {
        const u32 n = (u32)(idx & (N - 1));
        const u64 r = idx >> logN;  // (b, k, i) with i < l
        u32 i;
        const u64 bk = dl.divmod(r, i);
        const u32 k = (u32)(bk & 1);
        const u64 b = bk >> 1;
        const PrimeConst pc = pp[i].pc;
        const u64 v = acc[((bk * (l + 1) + i) << logN) + n];
        const u64 sp = barrett_red64(acc[((bk * (l + 1) + l) << logN) + n], pc);  // special-prime row, un-centred
        // modswitch(P ct + acc) = (P ct_i + acc_i - sp) P^-1 = ct_i + (acc_i - sp) P^-1 (mod q_i): one product, not two
        u64 res = shoup_full(sub_mod(v, sp, pc.q), ra.inv[i].w, ra.inv[i].wp, pc.q);
        if (k + 1 < comps) res = add_mod(res, ct[(((b * comps + k) * l + i) << logN) + n], pc.q);
        out[idx] = res;
    }
}

int launch_ks_finish(tfb_ctx* c, const u64* ct, u32 comps, const u64* acc, u64* out, u64 batch, cudaStream_t st, u32 Lct, u32 first) {
    if (!batch) return TFB_OK;
    const u64 total = batch * 2 * c->L * c->N;
    const unsigned tb = 256, nb = grid_for(total, tb);
    { ProfScope ps(PC_KS_FINISH, st); ks_finish_kernel<<<nb, tb, 0, st>>>(ct, comps, acc, out, c->L, c->logN, c->d_pp, total, Lct ? Lct : c->L, first); }
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}
int launch_ks_finish_push(tfb_ctx* c, const u64* ct, u32 comps, const u64* acc, u64 batch, cudaStream_t st, u32 Lct, u32 first,
                          u64* const* outs, u64* const* flags, u32* count, u32* err, u32 rank, u32 world, u64 epoch) {
    if (!batch) return TFB_OK;
    PeerTab pt;
    for (u32 p = 0; p < world; p++) { pt.out[p] = outs[p]; pt.flag[p] = flags[p]; }
    pt.count = count; pt.err = err; pt.rank = rank; pt.world = world;
    const u64 total = batch * 2 * c->L * c->N;
    const unsigned tb = 256;
    unsigned nb = grid_for(total, tb);
    { ProfScope ps(PC_KS_FINISH, st); ks_finish_push_kernel<<<nb, tb, 0, st>>>(ct, comps, acc, c->L, c->logN, c->d_pp, total, Lct, first, pt, epoch); }
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}
int launch_ks_finish_raised(tfb_ctx* c, tfb_ctx* ext, const u64* ct, u32 comps, const u64* acc, u64* out, u64 batch,
                            cudaStream_t st) {
    if (!batch) return TFB_OK;
    const u64 P = ext->q[ext->L - 1];
    RaiseArgs ra;
    for (u32 i = 0; i < c->L; i++) {
        if (P % c->q[i] == 0) { tfb_set_error("special prime must differ from the ciphertext primes"); return TFB_EINVAL; }
        ra.inv[i] = h_tw(h_invmod(P % c->q[i], c->q[i]), c->q[i]);
    }
    const u64 total = batch * 2 * c->L * c->N;
    const unsigned tb = 256, nb = grid_for(total, tb);
    { ProfScope ps(PC_KS_FINISH, st); ks_finish_raised_kernel<<<nb, tb, 0, st>>>(ct, comps, acc, out, c->L, c->logN, c->d_pp, ra, total); }
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}

// ---------------------------------------- BFV plaintext maps (bfv.jl:21-29)
// pi^-1:  m in Z_t^N  ->  Delta * m  embedded in every prime          (bfv.jl:21-24)
// pi:     b in R_Q    ->  mod(divround(SignedMod(b_n), Delta), t)     (bfv.jl:26-29; rounding div_hacks.jl:120-135,
//                                                                       centred lift signedmod.jl:12-19)
// Exact decode in word arithmetic.  |x| = X or Q - X (X > floor(Q/2));  a = |x| + floor(Delta/2) < Q;
// y = floor(a / Delta) = round-half-away(|x| / Delta) for either parity of Delta.  y is small (about t/2), so it is
// estimated from the two leading mixed-radix digits of a and then CORRECTED with exact sign tests of
// rho(y) = a - y Delta (one Garner conversion each; |rho| << Q/2 near the true y, so "canonical value above Q/2"
// means negative): the loops below stop at the largest y with rho(y) >= 0 whatever the error of the estimate.
struct PlainArgs {
    u64 dm[MAXD];    // Delta mod q_i
    u64 hd[MAXD];    // floor(Delta/2) mod q_i
    double ratio1;   // (prod_{k<L-1} q_k) / Delta
    double ratio2;   // (prod_{k<L-2} q_k) / Delta   (0 if L == 1)
};
__global__ void bfv_encode_kernel(const u64* __restrict__ m, u64* __restrict__ out, const u32 L, const u32 logN, const u64 t,
                                  const PrimeParams* __restrict__ pp, const PlainArgs pa, const u64 total) {
    const u32 N = 1u << logN;
    for (u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (u64)gridDim.x * blockDim.x) {
        const u32 n = (u32)(idx & (N - 1));
        const u64 r = idx >> logN;  // (p, i)
        const u32 i = (u32)(r % L);
        const u64 p = r / L;
        const PrimeConst pc = pp[i].pc;
        out[idx] = barrett_mul((m[(p << logN) + n] % t) % pc.q, pa.dm[i], pc);
    }
}
__device__ __forceinline__ bool plain_rho_nonneg(const u64* a, const u64 y, u64* tmp, u64* dg, const u32 L, const GarnerTab g,
                                                 const PrimeParams* __restrict__ pp, const PlainArgs& pa) {
    for (u32 i = 0; i < L; i++) {
        const PrimeConst pc = pp[i].pc;
        tmp[i] = sub_mod(a[i], barrett_mul(y % pc.q, pa.dm[i], pc), pc.q);
    }
    garner_digits(tmp, dg, L, g, pp);
    return !mr_above_half(dg, L, g.halfmr);
}
__global__ void bfv_decode_kernel(const u64* __restrict__ in, u64* __restrict__ out, const u32 L, const u32 logN, const u64 t,
                                  const GarnerTab g, const PrimeParams* __restrict__ pp, const PlainArgs pa, const u64 total) {
    const u32 N = 1u << logN;
    const u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const u64 p = idx >> logN;
    const u32 n = (u32)(idx & (N - 1));
    u64 a[MAXD], tmp[MAXD], dg[MAXD];
    for (u32 i = 0; i < L; i++) a[i] = in[((p * L + i) << logN) + n];
    garner_digits(a, dg, L, g, pp);
    const bool neg = mr_above_half(dg, L, g.halfmr);
    for (u32 i = 0; i < L; i++) {
        const u64 q = pp[i].pc.q;
        a[i] = add_mod(neg ? neg_mod(a[i], q) : a[i], pa.hd[i], q);
    }
    garner_digits(a, dg, L, g, pp);
    double est = (double)dg[L - 1] * pa.ratio1;
    if (L > 1) est += (double)dg[L - 2] * pa.ratio2;
    u64 y = est > 0.0 ? (u64)est : 0;
    while (y > 0 && !plain_rho_nonneg(a, y, tmp, dg, L, g, pp, pa)) y--;
    while (plain_rho_nonneg(a, y + 1, tmp, dg, L, g, pp, pa)) y++;
    const u64 ym = y % t;
    out[idx] = (neg && ym) ? t - ym : ym;
}

// host: little-endian multi-word helpers for Delta (a free BFVParams field in the reference, bfv.jl:5-15)
static u64 limbs_mod(const u64* x, u32 n, u64 m) {
    u128 r = 0;
    for (int i = (int)n - 1; i >= 0; i--) r = ((r << 64) | x[i]) % m;
    return (u64)r;
}
static long double limbs_ld(const std::vector<u64>& x) {
    long double v = 0;
    for (int i = (int)x.size() - 1; i >= 0; i--) v = v * 18446744073709551616.0L + (long double)x[i];
    return v;
}
static void limbs_mul(std::vector<u64>& x, u64 m) {
    u64 carry = 0;
    for (auto& w : x) {
        const u128 cur = (u128)w * m + carry;
        w = (u64)cur;
        carry = (u64)(cur >> 64);
    }
    if (carry) x.push_back(carry);
}
static int plain_args(tfb_ctx* c, const u64* delta, u32 nl, PlainArgs* pa) {
    if (c->L > MAXD || !c->conv_ok) { tfb_set_error("bfv plaintext maps: basis too large for exact conversion"); return TFB_EUNSUPPORTED; }
    while (nl > 0 && delta[nl - 1] == 0) nl--;
    if (nl == 0) { tfb_set_error("bfv plaintext maps: Delta must be positive"); return TFB_EINVAL; }
    std::vector<u64> d(delta, delta + nl), half(nl);
    u64 carry = 0;
    for (int i = (int)nl - 1; i >= 0; i--) {
        half[i] = (d[i] >> 1) | (carry << 63);
        carry = d[i] & 1;
    }
    std::vector<u64> W1{1}, W2{1};
    for (u32 k = 0; k + 1 < c->L; k++) limbs_mul(W1, c->q[k]);
    for (u32 k = 0; k + 2 < c->L; k++) limbs_mul(W2, c->q[k]);
    std::vector<u64> Q = W1;
    limbs_mul(Q, c->q[c->L - 1]);
    const long double dl = limbs_ld(d);
    if (dl * 1099511627776.0L < limbs_ld(Q)) { tfb_set_error("bfv plaintext maps: Delta too small (Q / Delta must stay below 2^40)"); return TFB_EUNSUPPORTED; }
    for (u32 i = 0; i < c->L; i++) {
        pa->dm[i] = limbs_mod(d.data(), nl, c->q[i]);
        pa->hd[i] = limbs_mod(half.data(), nl, c->q[i]);
    }
    pa->ratio1 = (double)(limbs_ld(W1) / dl);
    pa->ratio2 = c->L > 1 ? (double)(limbs_ld(W2) / dl) : 0.0;
    return TFB_OK;
}
int launch_bfv_encode(tfb_ctx* c, u64 t, const u64* delta, u32 nl, const u64* m, u64* out, u64 polys, cudaStream_t st) {
    if (!polys) return TFB_OK;
    if (t == 0) { tfb_set_error("bfv encode: plaintext modulus must be positive"); return TFB_EINVAL; }
    PlainArgs pa;
    int rc = plain_args(c, delta, nl, &pa);
    if (rc) return rc;
    const u64 total = polys * c->L * c->N;
    const unsigned tb = 256, nb = grid_for(total, tb);
    { ProfScope ps(PC_ELEMENTWISE, st); bfv_encode_kernel<<<nb, tb, 0, st>>>(m, out, c->L, c->logN, t, c->d_pp, pa, total); }
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}
int launch_bfv_decode(tfb_ctx* c, u64 t, const u64* delta, u32 nl, const u64* in, u64* out, u64 polys, cudaStream_t st) {
    if (!polys) return TFB_OK;
    if (t == 0) { tfb_set_error("bfv decode: plaintext modulus must be positive"); return TFB_EINVAL; }
    PlainArgs pa;
    int rc = plain_args(c, delta, nl, &pa);
    if (rc) return rc;
    const u64 total = polys * c->N;
    const unsigned tb = 128;
    const u64 nb = (total + tb - 1) / tb;
    { ProfScope ps(PC_BFV_CONTRACT, st); bfv_decode_kernel<<<(unsigned)nb, tb, 0, st>>>(in, out, c->L, c->logN, t, garner_of(c), c->d_pp, pa, total); }
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}

// ---------------------------------------- CKKS decode front end (ckksencoding.jl:60-66, ckks.jl:52-58)
// v[p][k] = (centred lift of coefficient k) / scale * exp(-i pi k / N) as a complex double.  The lift is evaluated
// from the mixed-radix digits by Horner in float64 (|x| = sum d_i prod_{j<i} q_j): relative error ~ L 2^-53, the
// precision of the reference's own Float64(n / denom).
struct LiftQ { double q[MAXD]; };
__global__ void ckks_lift_kernel(const u64* __restrict__ in, double* __restrict__ v, const u32 L, const u32 logN, const double scale,
                                 const GarnerTab g, const PrimeParams* __restrict__ pp, const LiftQ lq, const u64 total) {
    const u32 N = 1u << logN;
    const u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const u64 p = idx >> logN;
    const u32 k = (u32)(idx & (N - 1));
    u64 r[MAXD], d[MAXD];
    for (u32 i = 0; i < L; i++) r[i] = in[((p * L + i) << logN) + k];
    garner_digits(r, d, L, g, pp);
    const bool neg = mr_above_half(d, L, g.halfmr);
    if (neg) {   // |x| = Q - X: digits of the negated residues
        for (u32 i = 0; i < L; i++) r[i] = neg_mod(r[i], pp[i].pc.q);
        garner_digits(r, d, L, g, pp);
    }
    double x = 0.0;
    for (int i = (int)L - 1; i >= 0; i--) x = x * lq.q[i] + (double)d[i];
    x = (neg ? -x : x) / scale;
    double s, c;
    sincospi(-(double)k / (double)N, &s, &c);
    v[2 * idx] = x * c;
    v[2 * idx + 1] = x * s;
}
int launch_ckks_lift(tfb_ctx* c, double scale, const u64* in, double* v, u64 polys, cudaStream_t st) {
    if (c->L > MAXD || !c->conv_ok) { tfb_set_error("ckks decode: basis too large for exact conversion"); return TFB_EUNSUPPORTED; }
    LiftQ lq;
    for (u32 i = 0; i < c->L; i++) lq.q[i] = (double)c->q[i];
    const u64 total = polys * c->N;
    const unsigned tb = 128;
    const u64 nb = (total + tb - 1) / tb;
    { ProfScope ps(PC_LEVEL, st); ckks_lift_kernel<<<(unsigned)nb, tb, 0, st>>>(in, v, c->L, c->logN, scale, garner_of(c), c->d_pp, lq, total); }
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}

// ---------------------------------------- BGV plaintext map pi (bgv.jl:22-25): mod(SignedMod(x), t) per coefficient
// X in [0,Q) from its mixed-radix digits by Horner modulo t (exact integer arithmetic), minus Q mod t when the centred
// lift is negative (X > floor(Q/2), signedmod.jl:12-19); result in [0, t).
struct ModTArgs { u64 qt[MAXD]; u64 Qt; };   // q_i mod t, Q mod t
__global__ void centered_mod_kernel(const u64* __restrict__ in, u64* __restrict__ out, const u32 L, const u32 logN, const u64 t,
                                    const GarnerTab g, const PrimeParams* __restrict__ pp, const ModTArgs ma, const u64 total) {
    const u32 N = 1u << logN;
    const u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const u64 p = idx >> logN;
    const u32 n = (u32)(idx & (N - 1));
    u64 r[MAXD], d[MAXD];
    for (u32 i = 0; i < L; i++) r[i] = in[((p * L + i) << logN) + n];
    garner_digits(r, d, L, g, pp);
    const bool neg = mr_above_half(d, L, g.halfmr);
    u64 v = 0;
    for (int i = (int)L - 1; i >= 0; i--) v = (u64)(((u128)v * ma.qt[i] + d[i] % t) % t);
    if (neg) v = v >= ma.Qt ? v - ma.Qt : v + t - ma.Qt;
    out[idx] = v;
}
int launch_centered_mod(tfb_ctx* c, u64 t, const u64* in, u64* out, u64 polys, cudaStream_t st) {
    if (!polys) return TFB_OK;
    if (t == 0) { tfb_set_error("centered_mod: modulus must be positive"); return TFB_EINVAL; }
    if (c->L > MAXD || !c->conv_ok) { tfb_set_error("centered_mod: basis too large for exact conversion"); return TFB_EUNSUPPORTED; }
    ModTArgs ma;
    u64 Q = 1 % t;
    for (u32 i = 0; i < c->L; i++) {
        ma.qt[i] = c->q[i] % t;
        Q = (u64)((u128)Q * ma.qt[i] % t);
    }
    ma.Qt = Q;
    const u64 total = polys * c->N;
    const unsigned tb = 128;
    const u64 nb = (total + tb - 1) / tb;
    { ProfScope ps(PC_LEVEL, st); centered_mod_kernel<<<(unsigned)nb, tb, 0, st>>>(in, out, c->L, c->logN, t, garner_of(c), c->d_pp, ma, total); }
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}
