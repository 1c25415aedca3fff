/* toyfhe_b200.h -- C-ABI of libtoyfhe_b200.so, the B200 (sm_100a) engine behind
 * ToyFHE.jl's power-of-two cyclotomic ring path.
 *
 * The reference has no FFI today: the seam is Julia multiple dispatch on the
 * RingElement storage type.  Each entry point below names the reference method
 * it replaces (paths relative to the ToyFHE.jl repository); INTEGRATION.md shows
 * the `ccall` stubs a maintainer would add at exactly those methods.
 *
 * Conventions
 *   - element  = one residue in one 64-bit word, canonical in [0,q)  (PrimeField.n)
 *   - RNS poly = u64 [L][N], residue-major (StructArray fieldarrays, crt.jl:150-156)
 *   - batches  = u64 [batch][components][L][N]; "rows" counts [N]-rows, row r
 *                belongs to prime r % L (rows must be a multiple of L)
 *   - all buffer arguments are DEVICE pointers, except in the *_host entry points,
 *     which take HOST pointers and do the host<->device copies themselves
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); calls are
 *     asynchronous on that stream.  HOST side: calls on one ctx must not overlap in time (one host thread
 *     at a time per ctx).  DEVICE side: a ctx owns scratch buffers that its composite operations share;
 *     consecutive calls on the same stream are ordered by the stream, and calls on DIFFERENT streams are
 *     ordered by the library (the later call's stream waits on an event recorded on the earlier call's
 *     stream), so issuing work on one ctx from several streams is safe but does not overlap.
 *     Distinct contexts (and devices) are independent.  Every entry point runs with ctx's device current
 *     and restores the caller's device before returning
 *   - in-place (out == in) is allowed unless stated otherwise
 *   - every function returns 0 on success or a TFB_E* code; the message is
 *     available from tfb_last_error() (thread-local).  Nothing throws or aborts
 *     across the ABI (the reference raises Julia exceptions / @assert:
 *     pow2_cyc_rings.jl:31,61,116; rlwe_she.jl:223-225,248; crt.jl:269-275).
 *   - moduli must be odd primes < 2^62 with q = 1 (mod 2N); psi a primitive 2N-th
 *     root (pow2_cyc_rings.jl:27-47).
 */
#ifndef TOYFHE_B200_H
#define TOYFHE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tfb_ctx tfb_ctx;

enum {
    TFB_OK = 0,
    TFB_EINVAL = 1,       /* invalid argument (UsageError / @assert in the reference) */
    TFB_ECUDA = 2,        /* CUDA runtime failure */
    TFB_EUNSUPPORTED = 3, /* e.g. N > 2^16, modulus >= 2^62 */
    TFB_ENOMEM = 4
};

/* ---- diagnostics -------------------------------------------------------- */
const char* tfb_last_error(void);
int tfb_version(void);
/* number of engine kernels launched by this process so far (bench.py "gpu_launches") */
unsigned long long tfb_kernel_launches(void);

/* per-kernel-class device timing (CUDA events on the launching stream), used by
 * bench.py for the roofline figures: enable, run, then read totals per class. */
int tfb_profile_enable(int on);
int tfb_profile_classes(void);
const char* tfb_profile_class_name(int cls);
int tfb_profile_read(unsigned long long* counts, double* total_ms, int reset);
/* testing hook: route base conversions through the generic runtime-L kernels even
 * where a register-resident specialisation exists (on = 1), or keep the specialised kernels but reduce 128-bit sums with
 * the generic Shoup/Barrett step instead of the Solinas folds used on 2^60 + e primes (on = 2); all must agree bit for bit */
int tfb_debug_force_generic(int on);
/* kernel selection hook (testing): 1 = one CTA per row (512 threads x 32 residues, ntt_core.cuh),
 * 3 (default) = persistent TMA-prefetched third-generation kernels; both must agree bit for bit */
int tfb_debug_ntt_version(int v);
/* testing hook for N = 2^15 / 2^16, forward, out of place: 1 (default) = the row's last global level is applied while the
 * sub-block kernel loads its operands (one HBM pass less), 0 = every global level as its own pass; must agree bit for bit */
int tfb_debug_ntt_cross(int on);
/* testing hook: force the Harvey (conditional subtract per level) forward ladder even when every prime
 * qualifies for the lazy ladder */
int tfb_debug_ntt_force_harvey(int on);
/* testing hook: cap the forward-ladder range policy (0 = Harvey, 1 = lazy, 2 = lazy + approximate quotient,
 * the default when every prime is 2^60 + e with e < 2^28); all must agree bit for bit */
int tfb_debug_ntt_max_mode(int m);

/* ---- ring construction helpers (host only, no GPU needed) ---------------- */
/* NegacyclicRing(N, logqs) prime chain, crt.jl:282-295: ascending-logq order,
 * p = nextprime(max(2^logq+1, last+2N); interval=2N); psi = minimal primitive
 * 2N-th root of each prime (GaloisFields.minimal_primitive_root, crt.jl:142-144). */
int tfb_prime_chain(uint32_t N, const int32_t* logqs, uint32_t n, uint64_t* q_out, uint64_t* psi_out);
/* GaloisFields.minimal_primitive_root(F_q, n), n a power of two (pow2_cyc_rings.jl:40) */
int tfb_minimal_primitive_root(uint64_t q, uint64_t n, uint64_t* out);
/* ndigits(Q, base=2^w) for Q = prod q_i (rlwe_she.jl:280,333) */
int tfb_ndigits(const uint64_t* q, uint32_t L, uint32_t w, uint32_t* out);

/* ---- context = one NegacyclicRing{CRTEncoded{L}, N}(psi) ------------------ */
/* pow2_cyc_rings.jl:27-47 + crt.jl:282-295.  Precomputes twiddle / Garner tables
 * on `device` (the reference rebuilds twiddles on every transform, :298-301). */
int tfb_ctx_create(int device, uint32_t N, uint32_t L, const uint64_t* q, const uint64_t* psi, tfb_ctx** out);
int tfb_ctx_destroy(tfb_ctx* ctx);
int tfb_ctx_info(const tfb_ctx* ctx, uint32_t* N, uint32_t* L, uint64_t* q /*[L] or NULL*/, uint64_t* psi /*[L] or NULL*/);

/* ---- device memory helpers for hosts without their own CUDA binding ------- */
int tfb_malloc(tfb_ctx* ctx, size_t bytes, void** dptr);
int tfb_free(tfb_ctx* ctx, void* dptr);
int tfb_memcpy_h2d(tfb_ctx* ctx, void* dst, const void* src, size_t bytes, void* stream);
int tfb_memcpy_d2h(tfb_ctx* ctx, void* dst, const void* src, size_t bytes, void* stream);
int tfb_sync(tfb_ctx* ctx, void* stream);

/* ---- transforms ----------------------------------------------------------- */
/* NTT.nntt: c^[k] = sum_j c[j] psi^(j(2k+1)), natural order in and out
 * (pow2_cyc_rings.jl:295-303; RNS dispatch crt.jl:247-256). */
int tfb_ntt_fwd(tfb_ctx* ctx, const uint64_t* in, uint64_t* out, uint64_t rows, void* stream);
/* NTT.inntt (pow2_cyc_rings.jl:308-318; crt.jl:258-267). */
int tfb_ntt_inv(tfb_ctx* ctx, const uint64_t* in, uint64_t* out, uint64_t rows, void* stream);

/* ---- coefficient-wise ring ops (either domain) ----------------------------- */
/* CRTEncoded + - * (crt.jl:120-134) broadcast over RingElement storage
 * (pow2_cyc_rings.jl:167, 192-219) */
int tfb_add(tfb_ctx* ctx, const uint64_t* a, const uint64_t* b, uint64_t* out, uint64_t rows, void* stream);
int tfb_sub(tfb_ctx* ctx, const uint64_t* a, const uint64_t* b, uint64_t* out, uint64_t rows, void* stream);
int tfb_mul(tfb_ctx* ctx, const uint64_t* a, const uint64_t* b, uint64_t* out, uint64_t rows, void* stream);
int tfb_neg(tfb_ctx* ctx, const uint64_t* a, uint64_t* out, uint64_t rows, void* stream);
/* scalar_mul (pow2_cyc_rings.jl:177-185); s_residues = HOST array [L], s mod q_i */
int tfb_scalar_mul(tfb_ctx* ctx, const uint64_t* a, const uint64_t* s_residues, uint64_t* out, uint64_t rows, void* stream);
/* plaintext multiply broadcast over a batch: out[p] (+)= a[p] (.) plain for p < polys, a/out [polys][L][N], plain [L][N],
 * all in the same domain (dual for a ring product).  Replaces `map(c -> c * re, c.cs)` of the CKKS plaintext-vector
 * multiply (ckksencoding.jl:106-111) over every ciphertext of a batch, and with accumulate != 0 the `result += ...` of
 * the diagonal-method matmuls (test/ckks_matmul.jl:34-42, examples/encrypted_mnist/infer.jl:142-151). */
int tfb_mul_plain(tfb_ctx* ctx, const uint64_t* a, const uint64_t* plain, uint64_t* out, uint64_t polys, int accumulate, void* stream);
/* linear combinations with scalar weights: out[c] = sum_{j<J} weights[c][j] * in[j] for c < C (C <= 4, J <= 63), where in[j] is
 * the [polys][L][N] buffer in + j*in_stride_words, weights is a DEVICE array [C][J][L] of residues (weight mod q_i) and out is
 * [C][polys][L][N].  Replaces the sums of `c * b::AbstractFloat` (ckksencoding.jl:100-103) that make up a convolution over
 * ciphertexts (examples/encrypted_mnist/infer.jl:117-121): every input is read once for all C outputs. */
int tfb_lincomb(tfb_ctx* ctx, const uint64_t* in, uint64_t in_stride_words, uint32_t J, const uint64_t* weights, uint32_t C, uint64_t* out,
                uint64_t polys, void* stream);
/* plaintext add broadcast over a batch: out[p] = a[p] + plain for p < polys, where polynomial p starts stride_words words
 * after polynomial p-1 (>= L*N, even) and plain is one [L][N] element.  `c .+ b` of ckksencoding.jl:113-125 adds the encoded
 * plaintext to the FIRST component of every ciphertext of a batch: a = out = component 0 of ciphertext 0, stride = comps*L*N. */
int tfb_add_plain(tfb_ctx* ctx, const uint64_t* a, const uint64_t* plain, uint64_t* out, uint64_t polys, uint64_t stride_words, void* stream);

/* ring_multiply / * (pow2_cyc_rings.jl:147-173): primal in, primal out */
int tfb_ring_mul(tfb_ctx* ctx, const uint64_t* a, const uint64_t* b, uint64_t* out, uint64_t rows, void* stream);

/* apply_galois_element on primal rows (pow2_cyc_rings.jl:321-329); g odd; not in place */
int tfb_galois(tfb_ctx* ctx, uint64_t g, const uint64_t* in, uint64_t* out, uint64_t rows, void* stream);

/* ---- RNS level changes ------------------------------------------------------ */
/* modswitch(::RingElement) = exact division by the last prime, CKKS rescale
 * (crt.jl:215-220, 226-228; ckksencoding.jl:127-130): in [polys][L][N] -> out [polys][L-1][N] */
int tfb_rescale(tfb_ctx* ctx, const uint64_t* in, uint64_t* out, uint64_t polys, void* stream);
/* c .* CRTExpand{P}: multiply by P, append a zero residue (crt.jl:35-40;
 * modulusraising.jl:35-41): in [polys][L][N] -> out [polys][L+1][N] */
int tfb_crt_expand(tfb_ctx* ctx, uint64_t P, const uint64_t* in, uint64_t* out, uint64_t polys, void* stream);

/* ---- ciphertext multiply ----------------------------------------------------- */
/* enc_mul without basis change (CKKS / BGV form, rlwe_she.jl:247-262 with the
 * default mul_expand/mul_contract): c1,c2 [batch][2][L][N] primal -> out [batch][3][L][N] primal */
int tfb_ct_tensor(tfb_ctx* ctx, const uint64_t* c1, const uint64_t* c2, uint64_t* out, uint64_t batch, void* stream);

/* BFV mul_expand = switch(R_big, c) (bfv.jl:34, 202-226): centred lift from
 * ctx_from's modulus, reduced into ctx_to's basis: in [polys][Lf][N] -> out [polys][Lt][N] */
int tfb_bfv_switch(tfb_ctx* ctx_from, tfb_ctx* ctx_to, const uint64_t* in, uint64_t* out, uint64_t polys, void* stream);
/* BFV mul_contract = switch(R, multround(e, t, Q)) (bfv.jl:35-40, 172-190, rounding
 * div_hacks.jl:120-135): in [polys][Lb][N] over ctx_big -> out [polys][L][N] over ctx_q */
int tfb_bfv_contract(tfb_ctx* ctx_q, tfb_ctx* ctx_big, uint64_t t, const uint64_t* in, uint64_t* out, uint64_t polys, void* stream);
/* BFV ciphertext multiply = expand, tensor in R_big, contract
 * (rlwe_she.jl:247-262 with bfv.jl:34-40): c1,c2 [batch][2][L][N] -> out [batch][3][L][N] */
int tfb_bfv_mul(tfb_ctx* ctx_q, tfb_ctx* ctx_big, uint64_t t, const uint64_t* c1, const uint64_t* c2, uint64_t* out, uint64_t batch, void* stream);

/* BFV plaintext maps.  Delta is the BFVParams field of the reference (bfv.jl:5-15;
 * floor(Q/t) in test/bfv_crt.jl:25-32 and bfv.jl:118), passed as n_limbs little-endian 64-bit words; Q/Delta < 2^40.
 * pi^-1 (bfv.jl:21-24): m [polys][N] (any words; reduced mod t) -> Delta * m over ctx's primes, out [polys][L][N].
 * pi (bfv.jl:26-29): b [polys][L][N] primal -> mod(divround(SignedMod(b_n), Delta), t), out [polys][N], with the
 * centred lift of signedmod.jl:12-19 and round-half-away-from-zero of div_hacks.jl:120-135; exact. */
int tfb_bfv_encode(tfb_ctx* ctx, uint64_t t, const uint64_t* delta_limbs, uint32_t n_limbs, const uint64_t* m, uint64_t* out, uint64_t polys, void* stream);
int tfb_bfv_decode(tfb_ctx* ctx, uint64_t t, const uint64_t* delta_limbs, uint32_t n_limbs, const uint64_t* b, uint64_t* out, uint64_t polys, void* stream);

/* BGV plaintext map pi (bgv.jl:22-25): mod(SignedMod(b_n), t) per coefficient -- centred lift of signedmod.jl:12-19,
 * exact.  b [polys][L][N] primal -> out [polys][N] in [0, t).  (pi^-1 is the plain embedding of the plaintext.) */
int tfb_centered_mod(tfb_ctx* ctx, uint64_t t, const uint64_t* b, uint64_t* out, uint64_t polys, void* stream);

/* ---- CKKS encoding on the device -------------------------
 * The complex-FFT maps of src/ckksencoding.jl between N/2 complex slots (interleaved re, im float64, device memory)
 * and a real-coefficient plaintext polynomial; `scale` is the FixedRational denominator (ckks.jl:30-47).
 * encode (ckksencoding.jl:76-101): slots [polys][N/2] -> out [polys][L][N] primal, round(scale * coefficient) embedded
 *        in every prime; |scale * coefficient| must stay below 2^126 (the 2^70 scales of docs/src/man/ckks.md fit);
 *        synchronises `stream` before returning (the range check is read back).
 * decode (ckksencoding.jl:60-70):  in [polys][L][N] primal -> slots [polys][N/2] (centred lift / scale, DFT).
 * Floating point like the reference's FFTW calls: parity is the reference tests' tolerance, not bit-exactness. */
int tfb_ckks_encode(tfb_ctx* ctx, double scale, const double* slots, uint64_t* out, uint64_t polys, void* stream);
int tfb_ckks_decode(tfb_ctx* ctx, double scale, const uint64_t* in, double* slots, uint64_t polys, void* stream);

/* ---- sampling on the device ------------------------------
 * RingSampler (poly.jl:7-23) for keygen / encrypt (rlwe_she.jl:155-195).  The reference's RNG is Julia's unseeded
 * global one, so parity with it is statistical only; these are counter-based (Philox4x32-10, key = seed, counter =
 * (position, stream, attempt)): reproducible, independent of launch geometry, restated bit for bit on the CPU by
 * oracle/sampler_oracle.py.
 * uniform:  out [polys][L][N], every residue independent and uniform in [0, q_i) (crt.jl:146-148, 277-279).
 * gaussian: out [polys][L][N], x = round(sigma z), z ~ N(0,1) (Box-Muller), the same integer under every prime
 *           (DiscreteNormal(0, sigma): bfv.jl:31-32, ckks.jl:24-25). */
int tfb_sample_uniform(tfb_ctx* ctx, uint64_t seed, uint32_t stream_id, uint64_t* out, uint64_t polys, void* stream);
int tfb_sample_gaussian(tfb_ctx* ctx, double sigma, uint64_t seed, uint32_t stream_id, uint64_t* out, uint64_t polys, void* stream);

/* ---- key switching ------------------------------------------------------------ */
/* Digit polynomials of keyswitch (rlwe_she.jl:326-338).  cend = last ciphertext
 * component [batch][L][N] (primal, contiguous); relin_window w == 0 -> CRT digits
 * (D = L, centred residue re-embedded), w > 0 -> base-2^w digits of the un-centred
 * integer (D = ndigits(Q, 2^w)).  Digits are embedded in ctx_target's basis
 * (ctx itself, or the raised ring): out [batch][D][Lt][N] primal. */
int tfb_keyswitch_digits(tfb_ctx* ctx, tfb_ctx* ctx_target, uint32_t w, const uint64_t* cend, uint64_t* out, uint64_t batch, void* stream);
/* keyswitch(ek, c) (rlwe_she.jl:315-347).  key_dual = evaluation key already in
 * the NTT domain (the reference caches key.mask/key.masked duals in the
 * RingElement), [D][2][L'][N] with component 0 = mask, 1 = masked, over ctx's
 * primes (L' = L) or, when ctx_ext != NULL (ModulusRaised, modulusraising.jl:35-49),
 * over ctx_ext's primes = ctx's primes followed by the special prime (L' = L+1);
 * the host picks those key rows (downswitch_keyelement, crt.jl:238-244,
 * modulusraising.jl:43-49).  ct [batch][comps][L][N] primal, comps in {2,3};
 * out [batch][2][L][N] primal. */
int tfb_keyswitch(tfb_ctx* ctx, tfb_ctx* ctx_ext, uint32_t w, const uint64_t* key_dual, uint32_t D,
                  const uint64_t* ct, uint32_t comps, uint64_t* out, uint64_t batch, void* stream);

/* Residue-sharded keyswitch (BASELINE config 4, "residues sharded 2/4/8 GPU"): this rank owns
 * the primes [first, first + Ls) of ctx, ctx_shard = the ring over exactly those primes.  ct is the whole ciphertext
 * [batch][comps][L][N] (replicated on every rank): the digit decomposition needs every residue of a coefficient
 * (rlwe_she.jl:328-337), but each digit polynomial is then transformed and multiplied only under this rank's primes
 * with this rank's rows of the key, key_dual_shard [D][2][Ls][N] -- per-prime work is independent (crt.jl:250-254).
 * out_shard [batch][2][Ls][N] = rows first..first+Ls-1 of what tfb_keyswitch returns; the caller assembles the rows
 * of all ranks with ONE all-gather (toyfhe.jl_b200/sharding.py).  Plain parameters only (no special prime). */
int tfb_keyswitch_shard(tfb_ctx* ctx, tfb_ctx* ctx_shard, uint32_t first, uint32_t w, const uint64_t* key_dual_shard, uint32_t D,
                        const uint64_t* ct, uint32_t comps, uint64_t* out_shard, uint64_t batch, void* stream);

/* The same, with the exchange of the result rows fused into the last kernel (one process per GPU, NVLink peer memory
 * instead of a collective): every rank creates an exchange (two result slots of slot_bytes >= batch*2*L*N*8 plus flag
 * words, one cudaMalloc), the ranks swap the 64-byte IPC handles through whatever channel they have (the Python layer uses
 * torch.distributed.all_gather_object) and attach each other's buffers.  tfb_keyswitch_shard_push then runs the sharded
 * keyswitch and its epilogue stores this rank's rows into EVERY rank's slot, publishes a per-call epoch in the peers' flag
 * words and waits for theirs: when the launch completes (stream order), *result -- this rank's slot of the call, valid until
 * the call after next -- holds [batch][2][L][N], what tfb_keyswitch returns.  Every rank must make the same sequence of calls.
 * A peer that never arrives makes the wait give up after 2 s (tfb_xchg_check reports it) instead of hanging the GPU.
 * tfb_xchg_attach_ptr is the same-process form (threads / tests): the peer's tfb_xchg_local pointer itself.
 * tfb_xchg_destroy: the caller first makes sure (barrier) that no peer still writes into this rank's buffer. */
typedef struct tfb_xchg tfb_xchg;
int tfb_xchg_create(tfb_ctx* ctx, uint32_t rank, uint32_t world, uint64_t slot_bytes, tfb_xchg** out);
int tfb_xchg_export(tfb_xchg* x, uint8_t handle[64]);
int tfb_xchg_attach_ipc(tfb_xchg* x, uint32_t peer, const uint8_t handle[64]);
int tfb_xchg_attach_ptr(tfb_xchg* x, uint32_t peer, void* base);
int tfb_xchg_local(tfb_xchg* x, void** base);
int tfb_xchg_check(tfb_xchg* x, int* timed_out);
int tfb_xchg_destroy(tfb_xchg* x);
int tfb_keyswitch_shard_push(tfb_ctx* ctx, tfb_ctx* ctx_shard, uint32_t first, uint32_t w, const uint64_t* key_dual_shard, uint32_t D,
                             const uint64_t* ct, uint32_t comps, tfb_xchg* x, uint64_t** result, uint64_t batch, void* stream);

/* ---- host-buffer entry points (pinned or pageable host memory) ------------------ */
/* Same semantics as the device versions; copies in, runs, copies out and
 * synchronises `stream` before returning. */
int tfb_ntt_fwd_host(tfb_ctx* ctx, const uint64_t* in, uint64_t* out, uint64_t rows, void* stream);
int tfb_ntt_inv_host(tfb_ctx* ctx, const uint64_t* in, uint64_t* out, uint64_t rows, void* stream);
int tfb_ring_mul_host(tfb_ctx* ctx, const uint64_t* a, const uint64_t* b, uint64_t* out, uint64_t rows, void* stream);
int tfb_ct_tensor_host(tfb_ctx* ctx, const uint64_t* c1, const uint64_t* c2, uint64_t* out, uint64_t batch, void* stream);
int tfb_bfv_mul_host(tfb_ctx* ctx_q, tfb_ctx* ctx_big, uint64_t t, const uint64_t* c1, const uint64_t* c2, uint64_t* out, uint64_t batch, void* stream);
int tfb_rescale_host(tfb_ctx* ctx, const uint64_t* in, uint64_t* out, uint64_t polys, void* stream);
int tfb_bfv_encode_host(tfb_ctx* ctx, uint64_t t, const uint64_t* delta_limbs, uint32_t n_limbs, const uint64_t* m, uint64_t* out, uint64_t polys, void* stream);
int tfb_bfv_decode_host(tfb_ctx* ctx, uint64_t t, const uint64_t* delta_limbs, uint32_t n_limbs, const uint64_t* b, uint64_t* out, uint64_t polys, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TOYFHE_B200_H */
