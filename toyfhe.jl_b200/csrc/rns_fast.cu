// Register-resident, fully unrolled versions of the exact RNS base conversions
// for fixed basis sizes (the generic runtime-L kernels in rns_kernels.cu keep
// their per-thread digit arrays in local memory and are ~4x slower).
//
//   switch_fast<LF,LT>      : bfv.jl:202-226  switch/switchel   (mul_expand)
//   contract_fast<L,LB,KB>  : bfv.jl:35-40,172-190 multround + switch (mul_contract)
//
// All conversion constants travel in the kernel-parameter constant bank
// (__grid_constant__), so after unrolling every Garner / evaluation constant is
// an immediate c[0x0][..] operand of the IMADs -- no table loads at all.
//
// Sub-basis trick in contract: y_abs = floor((|x'| + (Q-1)/2) / Q), x' = centre(t x mod Qb), is bounded by
// Qb/(2Q) + 1, so it is already determined by its residues modulo the first KB
// primes of the big basis (host picks the smallest KB with prod_{j<KB} p_j above
// that bound; the Garner tables of a prefix basis are prefixes of the full ones).
#include <map>
#include <mutex>
#include <tuple>

#include "engine.h"

// the per-(context pair) constant tables below are process-wide caches: distinct contexts may be used from different
// host threads (include/toyfhe_b200.h "Threading"), so every lookup/insert holds this lock
static std::mutex g_fast_mu;
bool g_force_generic_red = false;  // testing hook: Shoup/Barrett reductions even on 2^60 + e primes

__device__ __forceinline__ u64 red128(u128 a, const PrimeConst& pc) {
    return red128_any((u64)(a >> 64), (u64)a, pc);
}

__host__ __device__ constexpr int tri(int i) { return i * (i - 1) / 2; }

template <int LN>
struct GarnerC {
    u64 gm[LN > 1 ? tri(LN) : 1];  // gm[tri(i)+j] = (prod_{k<j} q_k) mod q_i, j < i
    tw_t ginv[LN];                 // (prod_{k<i} q_k)^-1 mod q_i
    u64 half[LN];                  // mixed-radix digits of floor(Q/2)
    PrimeConst pc[LN];
};

template <int LN, int USE>
__device__ __forceinline__ void garner_reg(const u64 (&r)[LN], u64 (&d)[LN], const GarnerC<LN>& g) {
    d[0] = r[0];
#pragma unroll
    for (int i = 1; i < USE; i++) {
        u128 a = 0;
#pragma unroll
        for (int j = 0; j < i; j++) a += (u128)d[j] * g.gm[tri(i) + j];
        const u64 s = red128(a, g.pc[i]);
        d[i] = shoup_full(sub_mod(r[i], s, g.pc[i].q), g.ginv[i].w, g.ginv[i].wp, g.pc[i].q);
    }
}
template <int LN>
__device__ __forceinline__ bool above_half_reg(const u64 (&d)[LN], const GarnerC<LN>& g) {
    bool gt = false, decided = false;
#pragma unroll
    for (int i = LN - 1; i >= 0; i--) {
        const bool ne = d[i] != g.half[i];
        if (!decided && ne) gt = d[i] > g.half[i];
        decided = decided || ne;
    }
    return gt;
}
template <int LN, int USE>
__device__ __forceinline__ u64 eval_reg(const u64 (&d)[LN], const u64* ev, const PrimeConst& pc) {
    u128 a = 0;
#pragma unroll
    for (int i = 0; i < USE; i++) a += (u128)d[i] * ev[i];
    return red128(a, pc);
}

// ------------------------------------------------------------------ switch
template <int LF, int LT>
struct SwitchTab {
    GarnerC<LF> g;
    u64 ev[LT * LF];   // (prod_{k<i} qf_k) mod qt_j
    u64 qmod[LT];      // Qf mod qt_j
    PrimeConst pct[LT];
};

template <int LF, int LT>
__global__ void __launch_bounds__(128) switch_fast_kernel(const u64* __restrict__ in, u64* __restrict__ out, const u32 logN,
                                                          const u64 total, const __grid_constant__ SwitchTab<LF, LT> T) {
    const u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const u64 p = idx >> logN;
    const u32 n = (u32)(idx & ((1u << logN) - 1));
    u64 r[LF], d[LF];
#pragma unroll
    for (int i = 0; i < LF; i++) r[i] = in[((p * LF + i) << logN) + n];
    garner_reg<LF, LF>(r, d, T.g);
    const bool neg = above_half_reg<LF>(d, T.g);
#pragma unroll
    for (int j = 0; j < LT; j++) {
        u64 v = eval_reg<LF, LF>(d, T.ev + j * LF, T.pct[j]);
        if (neg) v = sub_mod(v, T.qmod[j], T.pct[j].q);
        out[((p * LT + j) << logN) + n] = v;
    }
}

// ------------------------------------------------------------------ contract
// Reference semantics (bfv.jl:172-174 with signedmod.jl:24-32): the multiplication by t happens in the CRT field
// (residue-wise, modulo Q_big) BEFORE the centred lift:  x' = centre((t X) mod Q_big),  y = rha(x' / Q).
template <int L, int LB, int KB>
struct ContractTab {
    GarnerC<LB> gb;
    GarnerC<L> gq;
    u64 ev_bq[L * LB];   // (prod_{m<k} p_m) mod q_i
    u64 pb_mod_q[L];     // Qb mod q_i
    u64 h_q[L];          // (Q-1)/2 mod q_i
    u64 ev_qb[KB * L];   // (prod_{m<i} q_m) mod p_j, j < KB
    tw_t t_b[LB];        // t mod p_j
    u64 h_b[KB];         // (Q-1)/2 mod p_j
    tw_t qinv_b[KB];     // Q^-1 mod p_j
};

template <int L, int LB, int KB>
__global__ void __launch_bounds__(128) contract_fast_kernel(const u64* __restrict__ in, u64* __restrict__ out, const u32 logN,
                                                            const u64 total, const __grid_constant__ ContractTab<L, LB, KB> T) {
    const u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const u64 p = idx >> logN;
    const u32 n = (u32)(idx & ((1u << logN) - 1));
    u64 rb[LB], db[LB];
#pragma unroll
    for (int j = 0; j < LB; j++)   // e.x * T(t): the product in the field, before any lift
        rb[j] = shoup_full(in[((p * LB + j) << logN) + n], T.t_b[j].w, T.t_b[j].wp, T.gb.pc[j].q);
    garner_reg<LB, LB>(rb, db, T.gb);
    const bool neg = above_half_reg<LB>(db, T.gb);
    // a = |x'| + h modulo every q_i, then R = a mod Q in mixed radix over the q basis
    u64 aq[L], dq[L];
#pragma unroll
    for (int i = 0; i < L; i++) {
        const PrimeConst& pc = T.gq.pc[i];
        u64 v = eval_reg<LB, LB>(db, T.ev_bq + i * LB, pc);
        if (neg) v = sub_mod(T.pb_mod_q[i], v, pc.q);
        aq[i] = add_mod(v, T.h_q[i], pc.q);
    }
    garner_reg<L, L>(aq, dq, T.gq);
    // y_abs = (a - R)/Q modulo the first KB big primes
#pragma unroll
    for (int j = 0; j < KB; j++) {
        const PrimeConst& pc = T.gb.pc[j];
        const u64 xa = neg ? neg_mod(rb[j], pc.q) : rb[j];
        const u64 a = add_mod(xa, T.h_b[j], pc.q);
        const u64 Rj = eval_reg<L, L>(dq, T.ev_qb + j * L, pc);
        rb[j] = shoup_full(sub_mod(a, Rj, pc.q), T.qinv_b[j].w, T.qinv_b[j].wp, pc.q);
    }
    garner_reg<LB, KB>(rb, db, T.gb);   // prefix basis: same tables
#pragma unroll
    for (int i = 0; i < L; i++) {
        const PrimeConst& pc = T.gq.pc[i];
        const u64 v = eval_reg<LB, KB>(db, T.ev_bq + i * LB, pc);
        out[((p * L + i) << logN) + n] = neg ? neg_mod(v, pc.q) : v;
    }
}

// ------------------------------------------------------------------ host side
template <int LN>
static void fill_garner(GarnerC<LN>& g, const tfb_ctx* c) {
    for (int i = 0; i < LN; i++) {
        const u64 qi = c->q[i];
        u64 M = 1 % qi;
        for (int j = 0; j < i; j++) {
            g.gm[tri(i) + j] = M;
            M = h_mulmod(M, c->q[j] % qi, qi);
        }
        g.ginv[i] = h_tw(i ? h_invmod(M, qi) : 1 % qi, qi);
        g.half[i] = c->halfmr[i];
        g.pc[i] = h_prime_const(qi);
    }
}
// (prod_{k<i} from_k) mod m for i < n, and the full product
static void fill_eval(u64* ev, int n, const tfb_ctx* from, u64 m, u64* full) {
    u64 M = 1 % m;
    for (int i = 0; i < n; i++) {
        ev[i] = M;
        M = h_mulmod(M, from->q[i] % m, m);
    }
    if (full) *full = M;
}
static u64 half_mod(const tfb_ctx* c, u64 m) {  // floor(Q/2) mod m via its mixed-radix digits
    u64 h = 0;
    for (int i = (int)c->L - 1; i >= 0; i--) h = (u64)(((u128)h * (c->q[i] % m) + c->halfmr[i] % m) % m);
    return h;
}

// smallest k such that prod_{j<k} p_j > Qb/(2Q) + 2  (bit-length estimate, conservative): |x'| <= Qb/2 whatever t is,
// because the product t x is reduced modulo Q_big before the lift
static int min_sub_basis(const tfb_ctx* cq, const tfb_ctx* cb, u64 t) {
    (void)t;
    long double need = 2.0L;  // slack bits
    for (u32 j = 0; j < cb->L; j++) need += log2l((long double)cb->q[j]);
    for (u32 i = 0; i < cq->L; i++) need -= log2l((long double)cq->q[i]);
    long double have = 0;
    for (u32 k = 0; k < cb->L; k++) {
        have += log2l((long double)cb->q[k]);
        if (have > need) return (int)k + 1;
    }
    return (int)cb->L;
}

template <int LF, int LT>
static int run_switch(tfb_ctx* from, tfb_ctx* to, const u64* in, u64* out, u64 polys, cudaStream_t st) {
    typedef SwitchTab<LF, LT> Tab;
    static std::map<std::pair<u64, u64>, Tab> cache;  // keyed by context uids (never reused)
    std::lock_guard<std::mutex> lk(g_fast_mu);
    auto key = std::make_pair(from->uid, to->uid);
    auto it = cache.find(key);
    if (it == cache.end()) {
        Tab t;
        fill_garner<LF>(t.g, from);
        for (int j = 0; j < LT; j++) {
            fill_eval(t.ev + j * LF, LF, from, to->q[j], &t.qmod[j]);
            t.pct[j] = h_prime_const(to->q[j]);
        }
        it = cache.emplace(key, t).first;
    }
    const u64 total = polys * from->N;
    const unsigned tb = 128;
    const u64 nb = (total + tb - 1) / tb;
    { ProfScope ps(PC_BASE_SWITCH, st); switch_fast_kernel<LF, LT><<<(unsigned)nb, tb, 0, st>>>(in, out, from->logN, total, it->second); }
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}

template <int L, int LB, int KB>
static int run_contract(tfb_ctx* cq, tfb_ctx* cb, u64 t, const u64* in, u64* out, u64 polys, cudaStream_t st) {
    typedef ContractTab<L, LB, KB> Tab;
    static std::map<std::tuple<u64, u64, u64>, Tab> cache;  // keyed by context uids (never reused)
    std::lock_guard<std::mutex> lk(g_fast_mu);
    auto key = std::make_tuple(cq->uid, cb->uid, t);
    auto it = cache.find(key);
    if (it == cache.end()) {
        Tab T;
        fill_garner<LB>(T.gb, cb);
        fill_garner<L>(T.gq, cq);
        for (int i = 0; i < L; i++) {
            const u64 qi = cq->q[i];
            fill_eval(T.ev_bq + i * LB, LB, cb, qi, &T.pb_mod_q[i]);
            T.h_q[i] = half_mod(cq, qi);
        }
        for (int j = 0; j < LB; j++) T.t_b[j] = h_tw(t % cb->q[j], cb->q[j]);
        for (int j = 0; j < KB; j++) {
            const u64 pj = cb->q[j];
            u64 Qm;
            fill_eval(T.ev_qb + j * L, L, cq, pj, &Qm);
            T.h_b[j] = half_mod(cq, pj);
            T.qinv_b[j] = h_tw(h_invmod(Qm, pj), pj);
        }
        it = cache.emplace(key, T).first;
    }
    const u64 total = polys * cq->N;
    const unsigned tb = 128;
    const u64 nb = (total + tb - 1) / tb;
    { ProfScope ps(PC_BFV_CONTRACT, st); contract_fast_kernel<L, LB, KB><<<(unsigned)nb, tb, 0, st>>>(in, out, cq->logN, total, it->second); }
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}

// ================================================================== joint basis
// tfb_bfv_mul's own extension basis.  The value of a BFV product does not depend on
// the big ring as long as the tensor integer x (|x| <= N Q^2 / 2 for canonical inputs)
// and y = round(t x / Q) (|y| <= t N Q / 2 + 1) are represented without wrap-around
// (bfv.jl:35-40,172-190 compute them over the integers).  So the fused multiply
// works over Q u P', P' = the first K primes of the caller's big ring with
// P' > 4 t N Q:  the Q-rows of the expanded operands are the inputs themselves,
// only K new residues are computed per coefficient, and the contraction needs one
// exact (non-centred) conversion r = (t x + h) mod Q -> P' and one centred
// conversion y: P' -> Q whose overflow count is unambiguous (|y| < P'/4).
//   expand_joint<L,K>   : [p][L][N] -> [p][L+K][N]   (rows 0..L-1 copied through)
//   contract_joint<L,K> : [p][L+K][N] -> [p][L][N]
// Second generation of the two kernels (what bounds them is the FMA-heavy pipe, profiles/r01_ncu_bfv_step.txt):
//   * Garner with the inverse folded into the constants: d_i = (a_i - sum_{j<i} d_j g_ij) ginv_i becomes ONE 128-bit
//     accumulation  a_i (c ginv_i) + sum_j d_j (-g_ij ginv_i)  and one reduction per digit (no Shoup products);
//     in the contraction a_i = t x_i + h is folded in as well (t ginv_i, h ginv_i);
//   * eta_j = (t x'_j + h - r mod p_j) comb_j likewise one accumulation x'_j (t comb_j) + h comb_j + sum_i d_i (-ev_ji comb_j);
//   * SP = every prime is 2^60 + e, e < 2^28 (the reference's nextprime(2^60+1) chains): 128-bit sums are reduced by two
//     Solinas folds at bit 60 (3 IMAD.WIDE + 3 IMAD) instead of a Shoup product plus a Barrett step (9 IMAD.WIDE + 6 IMAD).
// z mod q, canonical; SP: q = 2^60 + e and z < 2^126 (at most 16 products of residues below 2^61)
template <bool SP>
__device__ __forceinline__ u64 redj(const u128 z, const PrimeConst& pc, const u32 e) {
    if (!SP) return red128(z, pc);
    return red126_sp60((u64)(z >> 64), (u64)z, pc.q, e);
}

template <int LN>
struct GarnerF {
    u64 ng[LN > 1 ? tri(LN) : 1];  // ng[tri(i)+j] = -(prod_{k<j} q_k) ginv_i mod q_i, j < i
    u64 sc[LN];                    // multiplier of the residue: c ginv_i mod q_i
    u64 ad[LN];                    // constant term: h ginv_i mod q_i
    u64 half[LN];                  // mixed-radix digits of floor(Q/2)
    PrimeConst pc[LN];
    u32 e[LN];                     // q_i - 2^60 (SP)
};
// mixed-radix digits of the integer with residues sc^-1-scaled ... : d_i = (c r_i + h - sum_j d_j g_ij) ginv_i
template <int LN, bool SP, bool ADD>
__device__ __forceinline__ void garner_fused(const u64 (&r)[LN], u64 (&d)[LN], const GarnerF<LN>& g) {
#pragma unroll
    for (int i = 0; i < LN; i++) {
        if (!ADD && i == 0) {   // plain Garner (multiplier 1, no constant): the first digit is the first residue
            d[0] = r[0];
            continue;
        }
        u128 a = (u128)r[i] * g.sc[i];
        if (ADD) a += g.ad[i];
#pragma unroll
        for (int j = 0; j < i; j++) a += (u128)d[j] * g.ng[tri(i) + j];
        d[i] = redj<SP>(a, g.pc[i], g.e[i]);
    }
}
template <int LN>
__device__ __forceinline__ bool above_half_f(const u64 (&d)[LN], const GarnerF<LN>& g) {
    bool gt = false, decided = false;
#pragma unroll
    for (int i = LN - 1; i >= 0; i--) {
        const bool ne = d[i] != g.half[i];
        if (!decided && ne) gt = d[i] > g.half[i];
        decided = decided || ne;
    }
    return gt;
}

template <int L, int K>
struct ExpandJTab {
    GarnerF<L> g;
    u64 ev[K * L];   // (prod_{m<i} q_m) mod p_j
    u64 qmod[K];     // Q mod p_j
    PrimeConst pcb[K];
    u32 eb[K];       // p_j - 2^60 (SP)
    // fast route (SP): x = sum_i xi_i (Q/q_i) - v Q with xi_i = r_i (Q/q_i)^-1 mod q_i and v = floor(sum_i xi_i / q_i)
    u64 hinv[L];     // (Q/q_i)^-1 mod q_i
    u64 frac[L];     // floor(2^124 / q_i)
    u64 hev[K * L];  // (Q/q_i) mod p_j
    u64 nq[K];       // (-Q) mod p_j
};

// exact route: mixed-radix digits (Garner), sign from the digits of floor(Q/2), evaluation modulo every p_j
template <int L, int K, bool SP>
__device__ __noinline__ void expand_exact(const u64 (&r)[L], u64* __restrict__ out, const u64 base, const u32 logN, const ExpandJTab<L, K>& T) {
    u64 d[L];
    garner_fused<L, SP, false>(r, d, T.g);
    const bool neg = above_half_f<L>(d, T.g);
#pragma unroll 1
    for (int j = 0; j < K; j++) {
        u128 a = d[0];          // the weight of the first mixed-radix digit is 1
#pragma unroll
        for (int i = 1; i < L; i++) a += (u128)d[i] * T.ev[j * L + i];
        u64 v = redj<SP>(a, T.pcb[j], T.eb[j]);
        if (neg) v = sub_mod(v, T.qmod[j], T.pcb[j].q);
        out[base + ((u64)j << logN)] = v;
    }
}

template <int L, int K, bool SP>
__global__ void __launch_bounds__(128) expand_joint_kernel(const u64* __restrict__ in, u64* __restrict__ out, const u32 logN,
                                                           const u64 total, const u32 copyq, const __grid_constant__ ExpandJTab<L, K> T) {
    const u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const u64 p = idx >> logN;
    const u32 n = (u32)(idx & ((1u << logN) - 1));
    u64 r[L];
#pragma unroll
    for (int i = 0; i < L; i++) {
        r[i] = in[((p * L + i) << logN) + n];
        if (copyq) out[((p * (L + K) + i) << logN) + n] = r[i];
    }
    // copyq = 0: only the K new residues are written, [p][K][N] -- the forward transform reads the Q rows from the
    // caller's buffer (ntt_v3_kernels.cuh NttSrc), 8 of 17 row writes and their re-reads saved at the headline shape
    const u64 base = copyq ? ((p * (L + K) + L) << logN) + n : ((p * K) << logN) + n;
    if (!SP) {
        expand_exact<L, K, SP>(r, out, base, logN, T);
        return;
    }
    // Fast exact conversion.  sum_i xi_i / q_i = v + x/Q with x in [0,Q); the centred lift is x - Q iff x/Q > 1/2, so the
    // multiple of Q to remove is v' = floor(sum_i xi_i / q_i + 1/2).  Each fraction is taken to 60 bits, rounded DOWN by
    // less than 2 units: the 60-bit sum U under-estimates by less than 2L units, so floor(U / 2^60) is v' unless U sits
    // within 2L units below a multiple of 2^60 -- then (probability ~2^-55 per coefficient) the exact route decides.
    u64 xi[L];
    u64 U = 1ull << 59;
#pragma unroll
    for (int i = 0; i < L; i++) {
        xi[i] = redj<true>((u128)r[i] * T.hinv[i], T.g.pc[i], T.g.e[i]);
        U += __umul64hi(xi[i], T.frac[i]);
    }
    if ((U & ((1ull << 60) - 1)) >= (1ull << 60) - 2 * L - 2) {
        expand_exact<L, K, SP>(r, out, base, logN, T);
        return;
    }
    const u64 v = U >> 60;
#pragma unroll
    for (int j = 0; j < K; j++) {
        u128 a = (u128)v * T.nq[j];
#pragma unroll
        for (int i = 0; i < L; i++) a += (u128)xi[i] * T.hev[j * L + i];
        out[base + ((u64)j << logN)] = redj<true>(a, T.pcb[j], T.eb[j]);
    }
}

template <int L, int K>
struct ContractJTab {
    GarnerF<L> gq;           // sc = t ginv_i, ad = floor(Q/2) ginv_i: digits of r = (t x + floor(Q/2)) mod Q
    u64 xa[K];               // t comb_j mod p_j,  comb_j = Q^-1 (P'/p_j)^-1 mod p_j
    u64 ha[K];               // floor(Q/2) comb_j mod p_j
    u64 nev[K * L];          // -(prod_{m<i} q_m) comb_j mod p_j
    PrimeConst pcb[K];
    u32 eb[K];               // p_j - 2^60 (SP)
    u64 ev_bq[(K + 1) * L];  // [j*L + i] = (P'/p_j) mod q_i, j < K; [K*L + i] = (-P') mod q_i
    u32 sh[K];               // eta_j >> sh_j has at most 32 bits
    u32 R[K];                // floor(2^(58+sh_j) / p_j)
};

template <int L, int K, bool SP>
__global__ void __launch_bounds__(128) contract_joint_kernel(const u64* __restrict__ in, u64* __restrict__ out, const u32 logN,
                                                             const u64 total, const __grid_constant__ ContractJTab<L, K> T) {
    const u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const u64 p = idx >> logN;
    const u32 n = (u32)(idx & ((1u << logN) - 1));
    // r = (t x + h) mod Q as mixed-radix digits (exact, non-centred), a_i = t x_i + h folded into the Garner constants
    u64 x[L], d[L];
#pragma unroll
    for (int i = 0; i < L; i++) x[i] = in[((p * (L + K) + i) << logN) + n];
    garner_fused<L, SP, true>(x, d, T.gq);
    // y = (t x + h - r) / Q modulo p_j, pre-multiplied by (P'/p_j)^-1 for the CRT sum
    //   y = sum_j eta_j P'/p_j - w P',  w = round(sum_j eta_j / p_j)
    // accumulated on the fly into one 128-bit sum per q_i.  The loop over j is NOT unrolled:
    // fully unrolled the kernel is 68 KB of code and stalls 25% of the time on instruction
    // fetch (profiles/r01_ncu_bfv_step.txt); the constants are indexed in the parameter bank.
    u128 acc[L];
#pragma unroll
    for (int i = 0; i < L; i++) acc[i] = 0;
    u64 F = 1ull << 57;
#ifndef TFB_CONTRACT_UNROLL
#define TFB_CONTRACT_UNROLL 1
#endif
    constexpr int kUnroll = TFB_CONTRACT_UNROLL;
#pragma unroll kUnroll
    for (int j = 0; j < K; j++) {
        const PrimeConst pc = T.pcb[j];
        const u64 xj = in[((p * (L + K) + L + j) << logN) + n];
        u128 a = (u128)xj * T.xa[j] + T.ha[j];
        const u64* nv = T.nev + j * L;
#pragma unroll
        for (int i = 0; i < L; i++) a += (u128)d[i] * nv[i];
        const u64 eta = redj<SP>(a, pc, T.eb[j]);
        F += (u64)(u32)(eta >> T.sh[j]) * T.R[j];
        const u64* ev = T.ev_bq + j * L;
#pragma unroll
        for (int i = 0; i < L; i++) acc[i] += (u128)eta * ev[i];
    }
    // |y| < P'/4, so the 2^-21-accurate fixed-point sum F (scale 2^58) decides w without ambiguity
    const u64 w = F >> 58;
#pragma unroll
    for (int i = 0; i < L; i++) {
        acc[i] += (u128)w * T.ev_bq[K * L + i];
        out[((p * L + i) << logN) + n] = redj<SP>(acc[i], T.gq.pc[i], T.gq.e[i]);
    }
}

// number of leading primes of cb whose product exceeds 4 t N Q (0 if cb is too small
// or a prime is below 2^32, which the fixed-point rounding above assumes)
static int joint_basis_size(const tfb_ctx* cq, const tfb_ctx* cb, u64 t) {
    long double logQ = 0, logP = 0;
    for (u32 i = 0; i < cq->L; i++) logQ += log2l((long double)cq->q[i]);
    for (u32 k = 0; k < cb->L; k++) logP += log2l((long double)cb->q[k]);
    // The caller's big ring must itself hold t times every tensor value without wrap-around (P > t N Q^2): the
    // reference multiplies by t modulo P before the centred lift (signedmod.jl:24-32, bfv.jl:172-174), so below that
    // bound its result depends on P and only the conversion over the caller's own basis reproduces it.
    if (logP <= log2l((long double)t) + (long double)cq->logN + 2 * logQ + 0.01L) return 0;
    const long double need = 2.0L + 0.01L + log2l((long double)t) + (long double)cq->logN + logQ;
    long double have = 0;
    for (u32 k = 0; k < cb->L; k++) {
        if (cb->q[k] < (1ull << 32)) return 0;
        for (u32 i = 0; i < cq->L; i++) if (cq->q[i] == cb->q[k]) return 0;
        have += log2l((long double)cb->q[k]);
        if (have > need) return (int)k + 1;
    }
    return 0;
}

static bool is_sp_prime(u64 q) { return (q >> 60) == 1 && q - (1ull << 60) < (1ull << 28); }
// c = multiplier of the residues (1 or t), h = additive constant (0 or floor(Q/2)), both as values mod each q_i
template <int LN>
static void fill_garner_fused(GarnerF<LN>& g, const tfb_ctx* c, u64 mult, bool add_half) {
    for (int i = 0; i < LN; i++) {
        const u64 qi = c->q[i];
        u64 M = 1 % qi;
        u64 gm[LN > 0 ? LN : 1];
        for (int j = 0; j < i; j++) {
            gm[j] = M;
            M = h_mulmod(M, c->q[j] % qi, qi);
        }
        const u64 ginv = i ? h_invmod(M, qi) : 1 % qi;
        for (int j = 0; j < i; j++) {
            const u64 v = h_mulmod(gm[j], ginv, qi);
            g.ng[tri(i) + j] = v ? qi - v : 0;
        }
        g.sc[i] = h_mulmod(mult % qi, ginv, qi);
        g.ad[i] = add_half ? h_mulmod(half_mod(c, qi), ginv, qi) : 0;
        g.half[i] = c->halfmr[i];
        g.pc[i] = h_prime_const(qi);
        g.e[i] = (u32)(qi - (1ull << 60));
    }
}
static bool joint_sp(const tfb_ctx* cq, const tfb_ctx* cb, int K) {
    for (u32 i = 0; i < cq->L; i++) if (!is_sp_prime(cq->q[i])) return false;
    for (int j = 0; j < K; j++) if (!is_sp_prime(cb->q[j])) return false;
    return !g_force_generic_red;
}

template <int L, int K>
static int run_expand_joint(tfb_ctx* cq, tfb_ctx* cb, const u64* in, u64* out, u64 polys, cudaStream_t st, bool copyq) {
    typedef ExpandJTab<L, K> Tab;
    static std::map<std::pair<u64, u64>, Tab> cache;
    std::lock_guard<std::mutex> lk(g_fast_mu);
    auto key = std::make_pair(cq->uid, cb->uid);
    auto it = cache.find(key);
    if (it == cache.end()) {
        Tab t;
        fill_garner_fused<L>(t.g, cq, 1, false);
        for (int j = 0; j < K; j++) {
            fill_eval(t.ev + j * L, L, cq, cb->q[j], &t.qmod[j]);
            t.pcb[j] = h_prime_const(cb->q[j]);
            t.eb[j] = (u32)(cb->q[j] - (1ull << 60));
            t.nq[j] = t.qmod[j] ? cb->q[j] - t.qmod[j] : 0;
            for (int i = 0; i < L; i++) {   // (Q/q_i) mod p_j
                u64 M = 1 % cb->q[j];
                for (int m = 0; m < L; m++) if (m != i) M = h_mulmod(M, cq->q[m] % cb->q[j], cb->q[j]);
                t.hev[j * L + i] = M;
            }
        }
        for (int i = 0; i < L; i++) {
            const u64 qi = cq->q[i];
            u64 M = 1 % qi;                 // (Q/q_i) mod q_i
            for (int m = 0; m < L; m++) if (m != i) M = h_mulmod(M, cq->q[m] % qi, qi);
            t.hinv[i] = h_invmod(M, qi);
            t.frac[i] = qi > (1ull << 60) ? (u64)((((u128)1) << 124) / qi) : 0;   // only used on 2^60 + e primes (SP)
        }
        it = cache.emplace(key, t).first;
    }
    const u64 total = polys * cq->N;
    const unsigned tb = 128;
    const u64 nb = (total + tb - 1) / tb;
    {
        ProfScope ps(PC_BASE_SWITCH, st);
        if (joint_sp(cq, cb, K)) expand_joint_kernel<L, K, true><<<(unsigned)nb, tb, 0, st>>>(in, out, cq->logN, total, copyq ? 1u : 0u, it->second);
        else expand_joint_kernel<L, K, false><<<(unsigned)nb, tb, 0, st>>>(in, out, cq->logN, total, copyq ? 1u : 0u, it->second);
    }
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}

template <int L, int K>
static int run_contract_joint(tfb_ctx* cq, tfb_ctx* cb, u64 t, const u64* in, u64* out, u64 polys, cudaStream_t st) {
    typedef ContractJTab<L, K> Tab;
    static std::map<std::tuple<u64, u64, u64>, Tab> cache;
    std::lock_guard<std::mutex> lk(g_fast_mu);
    auto key = std::make_tuple(cq->uid, cb->uid, t);
    auto it = cache.find(key);
    if (it == cache.end()) {
        Tab T;
        fill_garner_fused<L>(T.gq, cq, t, true);
        for (int i = 0; i < L; i++) {
            const u64 qi = cq->q[i];
            u64 Pm = 1 % qi;   // P' mod q_i
            for (int j = 0; j < K; j++) Pm = h_mulmod(Pm, cb->q[j] % qi, qi);
            for (int j = 0; j < K; j++) {   // (P'/p_j) mod q_i
                u64 M = 1 % qi;
                for (int m = 0; m < K; m++) if (m != j) M = h_mulmod(M, cb->q[m] % qi, qi);
                T.ev_bq[j * L + i] = M;
            }
            T.ev_bq[K * L + i] = Pm ? qi - Pm : 0;
        }
        for (int j = 0; j < K; j++) {
            const u64 pj = cb->q[j];
            u64 Qm, ev[L];
            fill_eval(ev, L, cq, pj, &Qm);
            u64 M = 1 % pj;   // (P'/p_j) mod p_j
            for (int m = 0; m < K; m++) if (m != j) M = h_mulmod(M, cb->q[m] % pj, pj);
            const u64 comb = h_invmod(h_mulmod(Qm, M, pj), pj);
            T.xa[j] = h_mulmod(t % pj, comb, pj);
            T.ha[j] = h_mulmod(half_mod(cq, pj), comb, pj);
            for (int i = 0; i < L; i++) {
                const u64 v = h_mulmod(ev[i], comb, pj);
                T.nev[j * L + i] = v ? pj - v : 0;
            }
            T.pcb[j] = h_prime_const(pj);
            T.eb[j] = (u32)(pj - (1ull << 60));
            const u32 bits = 64 - (u32)__builtin_clzll(pj);
            T.sh[j] = bits > 32 ? bits - 32 : 0;
            T.R[j] = (u32)((((u128)1) << (58 + T.sh[j])) / pj);
        }
        it = cache.emplace(key, T).first;
    }
    const u64 total = polys * cq->N;
    const unsigned tb = 128;
    const u64 nb = (total + tb - 1) / tb;
    {
        ProfScope ps(PC_BFV_CONTRACT, st);
        if (joint_sp(cq, cb, K)) contract_joint_kernel<L, K, true><<<(unsigned)nb, tb, 0, st>>>(in, out, cq->logN, total, it->second);
        else contract_joint_kernel<L, K, false><<<(unsigned)nb, tb, 0, st>>>(in, out, cq->logN, total, it->second);
    }
    TFB_CUDA(cudaGetLastError());
    return TFB_OK;
}

#define JOINT_CASES(X) X(1, 2) X(1, 3) X(2, 3) X(2, 4) X(3, 4) X(3, 5) X(4, 5) X(4, 6) X(8, 9) X(8, 10)
// K > 0 when the joint-basis fast path applies to (cq, cb, t)
int fast_bfv_joint_k(const tfb_ctx* cq, const tfb_ctx* cb, u64 t) {
    if (!cq->conv_ok || cq->N != cb->N) return 0;
    const int K = joint_basis_size(cq, cb, t);
#define JK(A, B) if ((int)cq->L == A && K == B) return K;
    JOINT_CASES(JK)
#undef JK
    return 0;
}
int fast_expand_joint(tfb_ctx* cq, tfb_ctx* cb, int K, const u64* in, u64* out, u64 polys, cudaStream_t st, bool copyq) {
#define JE(A, B) if ((int)cq->L == A && K == B) return run_expand_joint<A, B>(cq, cb, in, out, polys, st, copyq);
    JOINT_CASES(JE)
#undef JE
    tfb_set_error("internal: no joint expand kernel");
    return TFB_EINVAL;
}
int fast_contract_joint(tfb_ctx* cq, tfb_ctx* cb, int K, u64 t, const u64* in, u64* out, u64 polys, cudaStream_t st) {
#define JC(A, B) if ((int)cq->L == A && K == B) return run_contract_joint<A, B>(cq, cb, t, in, out, polys, st);
    JOINT_CASES(JC)
#undef JC
    tfb_set_error("internal: no joint contract kernel");
    return TFB_EINVAL;
}

// returns 1 if a specialised kernel handled the call, 0 if the caller must use the
// generic path, <0 never; errors are reported through *rc
bool fast_base_switch(tfb_ctx* from, tfb_ctx* to, const u64* in, u64* out, u64 polys, cudaStream_t st, int* rc) {
    if (!from->conv_ok) return false;
#define SW(A, B) if (from->L == A && to->L == B) { *rc = run_switch<A, B>(from, to, in, out, polys, st); return true; }
    SW(8, 17) SW(2, 4) SW(3, 7) SW(1, 2) SW(4, 9)
#undef SW
    return false;
}
bool fast_bfv_contract(tfb_ctx* cq, tfb_ctx* cb, u64 t, const u64* in, u64* out, u64 polys, cudaStream_t st, int* rc) {
    if (!cq->conv_ok || !cb->conv_ok) return false;
    const int kb = min_sub_basis(cq, cb, t);
#define CT(A, B, K) if (cq->L == A && cb->L == B && kb <= K) { *rc = run_contract<A, B, K>(cq, cb, t, in, out, polys, st); return true; }
    CT(8, 17, 10) CT(8, 17, 17) CT(2, 4, 3) CT(2, 4, 4) CT(3, 7, 5) CT(3, 7, 7) CT(1, 2, 2) CT(4, 9, 6) CT(4, 9, 9)
#undef CT
    return false;
}
