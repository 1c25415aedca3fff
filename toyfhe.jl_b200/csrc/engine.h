// Internal declarations shared by the .cu translation units of libtoyfhe_b200.so
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/toyfhe_b200.h"
#include "modarith.cuh"

struct PrimeParams {
    PrimeConst pc;
    tw_t ninv;     // N^-1
    tw_t ninv_w1;  // N^-1 * psi^-brev(1)
    u32 sh;        // floor(log2 q) (shift of the lazy X reduction, ntt_core.cuh MODE 1)
    u32 pad_;
};

struct tfb_ctx {
    u64 uid;  // process-unique id (cache key for derived tables)
    int device;
    u32 N, logN, L;
    int num_sms;
    bool v3_ok;    // every prime is 2^b + e, 32 <= b <= 60, e < 2^28: third-generation kernels apply (ntt_core3.cuh)
    int ntt_mode;  // 0 = Harvey ladder, 1 = lazy ladder (all primes 2^b + small, 15q < 2^64), 2 = lazy + approximate quotient (all primes 2^60 + e, e < 2^28)
    std::vector<u64> q, psi;
    tw_t* d_fwd;      // [L][N]
    tw_t* d_inv;      // [L][N]
    PrimeParams* d_pp;  // [L]
    // Garner constants for conversions *from* this basis (rns_kernels.cu: build_garner)
    tw_t* d_ginv;     // [L]: Shoup pair of (prod_{k<i} q_k)^-1 mod q_i
    u64* d_halfmr;    // allocation: [L*L] Garner matrix, then [L] mixed-radix digits of floor(Q/2)
    std::vector<u64> halfmr;
    bool conv_ok;     // basis small enough for 128-bit lazy accumulation in conversions
    // scratch (device) and staging (device, for the *_host entry points)
    void* ws;
    size_t ws_bytes;
    void* stage;
    size_t stage_bytes;
    void* io;
    size_t io_bytes;
    // the scratch above is shared by every asynchronous call on this context: `scratch_done` is recorded on the stream of
    // the last call that used it and the next call's stream waits on it when it is a different stream (api.cu ScratchGuard)
    cudaEvent_t scratch_done;
    cudaStream_t scratch_stream;
    bool scratch_used;
    unsigned* d_ckks_pos;   // [N/2] slot positions (3^(i+1) mod 2N) >> 1 of the CKKS encoding (ckks_kernels.cu), built on first use
};

void tfb_set_error(const std::string& msg);
int tfb_cuda_fail(cudaError_t e, const char* what);
#define TFB_CUDA(x)                                                \
    do {                                                           \
        cudaError_t _e = (x);                                      \
        if (_e != cudaSuccess) return tfb_cuda_fail(_e, #x);       \
    } while (0)


int ws_reserve(tfb_ctx* c, size_t bytes);
int stage_reserve(tfb_ctx* c, size_t bytes);

// per-kernel-class timing with CUDA events on the launching stream (bench.py roofline)
enum ProfClass {
    PC_NTT_FWD = 0, PC_NTT_INV, PC_NTT_OTHER, PC_ELEMENTWISE, PC_TENSOR, PC_BASE_SWITCH, PC_BFV_CONTRACT,
    PC_KS_DIGITS, PC_KS_ACCUM, PC_KS_FINISH, PC_LEVEL, PC_COUNT
};
struct ProfScope {
    int cls;
    cudaStream_t st;
    cudaEvent_t stop;
    ProfScope(int cls, cudaStream_t st);
    ~ProfScope();
};

// ntt_kernels.cu
int launch_ntt(tfb_ctx* c, const u64* in, u64* out, u64 rows, bool inverse, cudaStream_t st);
int launch_ks_crt_ntt(tfb_ctx* c, tfb_ctx* r, const u64* cend, u64 ct_stride, u64* dig, u32 k0, u32 Dn, u64 batch, cudaStream_t st);
unsigned long long tfb_launch_count();
void tfb_count_launch(int n = 1);

// rns_kernels.cu
int launch_binop(tfb_ctx* c, int op, const u64* a, const u64* b, u64* out, u64 rows, cudaStream_t st);
int launch_neg(tfb_ctx* c, const u64* a, u64* out, u64 rows, cudaStream_t st);
int launch_add_plain(tfb_ctx* c, const u64* a, const u64* plain, u64* out, u64 polys, u64 stride_words, cudaStream_t st);
int launch_lincomb(tfb_ctx* c, const u64* in, u64 in_stride, u32 J, const u64* w, u32 C, u64* out, u64 polys, cudaStream_t st);
int launch_mul_plain(tfb_ctx* c, const u64* a, const u64* plain, u64* out, u64 polys, bool accumulate, cudaStream_t st);
int launch_scalar_mul(tfb_ctx* c, const u64* a, const u64* s_host, u64* out, u64 rows, cudaStream_t st);
int launch_tensor_dual(tfb_ctx* c, const u64* a, const u64* b, u64* out, u64 batch, cudaStream_t st);
int launch_galois(tfb_ctx* c, u64 g, const u64* in, u64* out, u64 rows, cudaStream_t st);
int launch_rescale(tfb_ctx* c, const u64* in, u64* out, u64 polys, cudaStream_t st);
int launch_crt_expand(tfb_ctx* c, u64 P, const u64* in, u64* out, u64 polys, cudaStream_t st);
int launch_base_switch(tfb_ctx* from, tfb_ctx* to, const u64* in, u64* out, u64 polys, cudaStream_t st);
int launch_bfv_contract(tfb_ctx* cq, tfb_ctx* cb, u64 t, const u64* in, u64* out, u64 polys, cudaStream_t st);
int launch_ks_digits(tfb_ctx* c, tfb_ctx* target, int w, const u64* cend, u64 ct_stride, u64* out, u32 k0, u32 Dn,
                     u64 batch, cudaStream_t st, bool compact = false);
int launch_ks_accum(tfb_ctx* c, u32 k0, u32 Dn, const u64* dig, const u64* key, u64* acc, int accumulate, u64 batch,
                    cudaStream_t st);
int launch_ks_finish(tfb_ctx* c, const u64* ct, u32 comps, const u64* acc, u64* out, u64 batch, cudaStream_t st, u32 Lct = 0, u32 first = 0);
int launch_ks_finish_raised(tfb_ctx* c, tfb_ctx* ext, const u64* ct, u32 comps, const u64* acc, u64* out, u64 batch,
                            cudaStream_t st);
int launch_bfv_encode(tfb_ctx* c, u64 t, const u64* delta, u32 nl, const u64* m, u64* out, u64 polys, cudaStream_t st);
int launch_bfv_decode(tfb_ctx* c, u64 t, const u64* delta, u32 nl, const u64* in, u64* out, u64 polys, cudaStream_t st);
int launch_centered_mod(tfb_ctx* c, u64 t, const u64* in, u64* out, u64 polys, cudaStream_t st);
int build_garner(tfb_ctx* c);
// ckks_kernels.cu
int launch_ckks_encode(tfb_ctx* c, double scale, const double* slots, u64* out, u64 polys, cudaStream_t st);
int launch_ckks_decode(tfb_ctx* c, double scale, const u64* in, double* slots, u64 polys, cudaStream_t st);
// sample_kernels.cu
int launch_sample_uniform(tfb_ctx* c, u64 seed, u32 stream, u64* out, u64 polys, cudaStream_t st);
int launch_sample_gaussian(tfb_ctx* c, double sigma, u64 seed, u32 stream, u64* out, u64 polys, cudaStream_t st);
// rns_fast.cu: specialised register-resident conversions; return false if no specialisation fits
bool fast_base_switch(tfb_ctx* from, tfb_ctx* to, const u64* in, u64* out, u64 polys, cudaStream_t st, int* rc);
bool fast_bfv_contract(tfb_ctx* cq, tfb_ctx* cb, u64 t, const u64* in, u64* out, u64 polys, cudaStream_t st, int* rc);
// joint-basis BFV multiply (Q u first K primes of the big ring); K = 0: not applicable
int fast_bfv_joint_k(const tfb_ctx* cq, const tfb_ctx* cb, u64 t);
int fast_expand_joint(tfb_ctx* cq, tfb_ctx* cb, int K, const u64* in, u64* out, u64 polys, cudaStream_t st, bool copyq = true);   // copyq = false: out [polys][K][N], new residues only
int fast_contract_joint(tfb_ctx* cq, tfb_ctx* cb, int K, u64 t, const u64* in, u64* out, u64 polys, cudaStream_t st);
void tfb_forget_ctx_pairs(const tfb_ctx* c);
int ntt_setup_device();
// ntt_kernels3.cu
int ntt3_setup_device();
int launch_ntt14p(tfb_ctx* c, const u64* in, u64* out, u64 rows, bool inverse, u32 s0, cudaStream_t st);
// ntt_kernels4.cu
int ntt4_setup_device();
int launch_ntt_s(tfb_ctx* c, const u64* in, u64* out, u64 rows, bool inverse, cudaStream_t st);
int launch_ntt_s_bcast(tfb_ctx* c, const u64* in, u64* out, u64 polys, cudaStream_t st);
int launch_ntt_s_gather(tfb_ctx* c, const void* src, u64* out, u64 rows, cudaStream_t st);   // src: v3k::NttSrc
int launch_ntt_s_crt(tfb_ctx* c, tfb_ctx* r, const u64* cend, u64 ct_stride, u64* dig, u32 k0, u32 dn, u64 batch, cudaStream_t st);
int launch_ntt_crt(tfb_ctx* c, tfb_ctx* r, const u64* cend, u64 ct_stride, u64* dig, u32 k0, u32 dn, u64 batch, cudaStream_t st);   // ntt_kernels3.cu
int launch_ntt_s_pow2(tfb_ctx* r, const u64* limbs, u32 nl, u32 w, u64* dig, u32 k0, u32 dn, u64 batch, cudaStream_t st);
int launch_ntt_pow2(tfb_ctx* r, const u64* limbs, u32 nl, u32 w, u64* dig, u32 k0, u32 dn, u64 batch, cudaStream_t st);              // ntt_kernels3.cu
#define TFB_MAX_PEERS 8
int launch_ks_finish_push(tfb_ctx* c, const u64* ct, u32 comps, const u64* acc, u64 batch, cudaStream_t st, u32 Lct, u32 first,
                          u64* const* outs, u64* const* flags, u32* count, u32* err, u32 rank, u32 world, u64 epoch);                 // rns_kernels.cu
int launch_ks_limbs(tfb_ctx* c, const u64* cend, u64 ct_stride, u64* limbs, u64 batch, cudaStream_t st);                             // rns_kernels.cu
// forward transform of `polys` polynomials of c->L rows each: rows [0, lq) of polynomial p from base0 (p < polys0) or base1,
// rows [lq, c->L) from ext [polys][c->L - lq][N]; out contiguous.  -1: kernel family not applicable.
int launch_ntt_gather(tfb_ctx* c, const u64* base0, const u64* base1, u32 polys0, u32 lq, const u64* ext, u64* out, u64 polys, cudaStream_t st);   // ntt_kernels3.cu
int launch_ntt_bcast(tfb_ctx* c, const u64* in, u64* out, u64 polys, cudaStream_t st);   // ntt_kernels3.cu
int launch_ntt_inv_sub(tfb_ctx* c, const u64* in, u64* out, u64 rows, u32 s0, cudaStream_t st);   // ntt_kernels3.cu
int launch_ntt_fwd_cross(tfb_ctx* c, const u64* in, u64* out, u64 rows, u32 s0, cudaStream_t st);   // ntt_kernels3.cu
extern bool g_ntt_force_harvey;
extern int g_ntt_max_mode;  // debug cap on the ladder mode (2 = no cap)
extern int g_ntt_version;  // 1 = one CTA per row (ntt_core.cuh), 3 = persistent third-generation kernels (default)
