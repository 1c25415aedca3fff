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
// Sub-basis trick in contract: y_abs = floor((t|x| + (Q-1)/2) / Q) is bounded by
// t*Qb/(2Q) + 1, so it is already determined by its residues modulo the first KB
// primes of the big basis (host picks the smallest KB with prod_{j<KB} p_j above
// that bound; the Garner tables of a prefix basis are prefixes of the full ones).
#include <map>
#include <tuple>

#include "engine.h"

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
template <int L, int LB, int KB>
struct ContractTab {
    GarnerC<LB> gb;
    GarnerC<L> gq;
    u64 ev_bq[L * LB];   // (prod_{m<k} p_m) mod q_i
    u64 pb_mod_q[L];     // Qb mod q_i
    tw_t t_q[L];         // t mod q_i
    u64 h_q[L];          // (Q-1)/2 mod q_i
    u64 ev_qb[KB * L];   // (prod_{m<i} q_m) mod p_j, j < KB
    tw_t t_b[KB];        // t mod p_j
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
    for (int j = 0; j < LB; j++) rb[j] = in[((p * LB + j) << logN) + n];
    garner_reg<LB, LB>(rb, db, T.gb);
    const bool neg = above_half_reg<LB>(db, T.gb);
    // a = t|x| + h modulo every q_i, then R = a mod Q in mixed radix over the q basis
    u64 aq[L], dq[L];
#pragma unroll
    for (int i = 0; i < L; i++) {
        const PrimeConst& pc = T.gq.pc[i];
        u64 v = eval_reg<LB, LB>(db, T.ev_bq + i * LB, pc);
        if (neg) v = sub_mod(T.pb_mod_q[i], v, pc.q);
        v = shoup_full(v, T.t_q[i].w, T.t_q[i].wp, pc.q);
        aq[i] = add_mod(v, T.h_q[i], pc.q);
    }
    garner_reg<L, L>(aq, dq, T.gq);
    // y_abs = (a - R)/Q modulo the first KB big primes
#pragma unroll
    for (int j = 0; j < KB; j++) {
        const PrimeConst& pc = T.gb.pc[j];
        const u64 xa = neg ? neg_mod(rb[j], pc.q) : rb[j];
        const u64 a = add_mod(shoup_full(xa, T.t_b[j].w, T.t_b[j].wp, pc.q), T.h_b[j], pc.q);
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

// smallest k such that prod_{j<k} p_j > t*Qb/(2Q) + 2  (bit-length estimate, conservative)
static int min_sub_basis(const tfb_ctx* cq, const tfb_ctx* cb, u64 t) {
    long double need = 2.0L;  // slack bits
    need += log2l((long double)t);
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
    auto key = std::make_tuple(cq->uid, cb->uid, t);
    auto it = cache.find(key);
    if (it == cache.end()) {
        Tab T;
        fill_garner<LB>(T.gb, cb);
        fill_garner<L>(T.gq, cq);
        for (int i = 0; i < L; i++) {
            const u64 qi = cq->q[i];
            fill_eval(T.ev_bq + i * LB, LB, cb, qi, &T.pb_mod_q[i]);
            T.t_q[i] = h_tw(t % qi, qi);
            T.h_q[i] = half_mod(cq, qi);
        }
        for (int j = 0; j < KB; j++) {
            const u64 pj = cb->q[j];
            u64 Qm;
            fill_eval(T.ev_qb + j * L, L, cq, pj, &Qm);
            T.t_b[j] = h_tw(t % pj, pj);
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
