/* CPU oracle (TEST INFRASTRUCTURE ONLY) -- C restatement of the ToyFHE.jl
 * power-of-two-cyclotomic / RNS hot path.  "restated reference (C), not Julia".
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (toyfhe.jl_b200/)
 * never does.  It is validated against oracle/toyfhe_oracle.py (Python big-int,
 * itself pinned to the reference's literal KATs) by tests/test_oracle_c.py.
 *
 * Algorithm follows the reference, not the GPU engine:
 *   - nntt  = PowMul_psi then natural-order radix-2 cyclic NTT over F_q
 *             (pow2_cyc_rings.jl:295-303), modmul = 128-bit product + %  (what
 *             GaloisFields' widemul+rem does);
 *   - inntt = cyclic NTT with w^-1, then x * N^-1 * psi^-i (pow2_cyc_rings.jl:308-318);
 *   - RNS ops are per-prime maps (crt.jl:120-134, 247-267);
 *   - base switch / multround / digit decomposition reconstruct every
 *     coefficient as a big integer first (crt.jl:105-112, bfv.jl:172-226,
 *     rlwe_she.jl:331-337) -- a small fixed-width bigint stands in for BigInt.
 * Twiddles are precomputed once per (q, N) -- kinder than the reference, which
 * rebuilds the plan per call (pow2_cyc_rings.jl:298-301).
 *
 * Parity status: primal-domain results are pinned through the Python oracle's
 * KATs; dual-domain ordering is "parity unpinned" (see toyfhe_oracle.py header).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef uint64_t u64;
typedef unsigned __int128 u128;

static inline u64 mulmod(u64 a, u64 b, u64 q) { return (u64)((u128)a * b % q); }
static inline u64 addmod(u64 a, u64 b, u64 q) { u64 s = a + b; return (s >= q || s < a) ? s - q : s; }
static inline u64 submod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }
static u64 powmod(u64 a, u64 e, u64 q) {
    u64 r = 1 % q;
    a %= q;
    while (e) { if (e & 1) r = mulmod(r, a, q); a = mulmod(a, a, q); e >>= 1; }
    return r;
}
static inline u64 invmod(u64 a, u64 q) { return powmod(a, q - 2, q); }

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------ plans */
typedef struct {
    u64 N, q, psi;
    int lg;
    u64 *psipow;   /* psi^i            */
    u64 *ipsipow;  /* N^-1 psi^-i      */
    u64 *wpow;     /* w^i, w=psi^2, i<N/2  */
    u64 *iwpow;    /* w^-i             */
    uint32_t *brev;
} plan_t;

plan_t *orc_plan_create(u64 N, u64 q, u64 psi) {
    plan_t *p = (plan_t *)calloc(1, sizeof(plan_t));
    p->N = N; p->q = q; p->psi = psi;
    int lg = 0; while ((1ull << lg) < N) lg++;
    p->lg = lg;
    p->psipow = (u64 *)malloc(N * 8); p->ipsipow = (u64 *)malloc(N * 8);
    p->wpow = (u64 *)malloc((N / 2 + 1) * 8); p->iwpow = (u64 *)malloc((N / 2 + 1) * 8);
    p->brev = (uint32_t *)malloc(N * 4);
    u64 ipsi = invmod(psi, q), w = mulmod(psi, psi, q), iw = mulmod(ipsi, ipsi, q);
    u64 ninv = invmod(N % q, q);
    u64 t = 1, ti = ninv;
    for (u64 i = 0; i < N; i++) { p->psipow[i] = t; p->ipsipow[i] = ti; t = mulmod(t, psi, q); ti = mulmod(ti, ipsi, q); }
    t = 1; ti = 1;
    for (u64 i = 0; i < N / 2 + 1; i++) { p->wpow[i] = t; p->iwpow[i] = ti; t = mulmod(t, w, q); ti = mulmod(ti, iw, q); }
    for (u64 i = 0; i < N; i++) { uint32_t r = 0; for (int b = 0; b < lg; b++) if (i >> b & 1) r |= 1u << (lg - 1 - b); p->brev[i] = r; }
    return p;
}
void orc_plan_destroy(plan_t *p) {
    if (!p) return;
    free(p->psipow); free(p->ipsipow); free(p->wpow); free(p->iwpow); free(p->brev); free(p);
}

/* natural-order in-place radix-2 DIT cyclic NTT; wp = table of root powers */
static void cyclic_ntt(const plan_t *p, u64 *a, const u64 *wp) {
    const u64 N = p->N, q = p->q;
    for (u64 i = 0; i < N; i++) { u64 j = p->brev[i]; if (i < j) { u64 t = a[i]; a[i] = a[j]; a[j] = t; } }
    for (u64 m = 2; m <= N; m <<= 1) {
        u64 half = m >> 1, step = N / m;
        for (u64 s = 0; s < N; s += m)
            for (u64 j = 0; j < half; j++) {
                u64 u = a[s + j], v = mulmod(a[s + j + half], wp[j * step], q);
                a[s + j] = addmod(u, v, q);
                a[s + j + half] = submod(u, v, q);
            }
    }
}

void orc_nntt(const plan_t *p, const u64 *in, u64 *out) {
    for (u64 i = 0; i < p->N; i++) out[i] = mulmod(in[i], p->psipow[i], p->q);
    cyclic_ntt(p, out, p->wpow);
}
void orc_inntt(const plan_t *p, const u64 *in, u64 *out) {
    if (out != in) memcpy(out, in, p->N * 8);
    cyclic_ntt(p, out, p->iwpow);
    for (u64 i = 0; i < p->N; i++) out[i] = mulmod(out[i], p->ipsipow[i], p->q);
}

/* ------------------------------------------------------------- RNS context */
typedef struct {
    u64 N; int L;
    u64 *q;
    plan_t **plan;
} rns_t;

rns_t *orc_rns_create(u64 N, int L, const u64 *q, const u64 *psi) {
    rns_t *r = (rns_t *)calloc(1, sizeof(rns_t));
    r->N = N; r->L = L;
    r->q = (u64 *)malloc(L * 8);
    r->plan = (plan_t **)malloc(L * sizeof(plan_t *));
    for (int i = 0; i < L; i++) { r->q[i] = q[i]; r->plan[i] = orc_plan_create(N, q[i], psi[i]); }
    return r;
}
void orc_rns_destroy(rns_t *r) {
    if (!r) return;
    for (int i = 0; i < r->L; i++) orc_plan_destroy(r->plan[i]);
    free(r->plan); free(r->q); free(r);
}

/* rows = number of [N] rows laid out [..][L][N]; prime of row r is r % L (crt.jl:247-267) */
void orc_rns_nntt(const rns_t *c, const u64 *in, u64 *out, long rows) {
#pragma omp parallel for schedule(static)
    for (long r = 0; r < rows; r++) orc_nntt(c->plan[r % c->L], in + r * c->N, out + r * c->N);
}
void orc_rns_inntt(const rns_t *c, const u64 *in, u64 *out, long rows) {
#pragma omp parallel for schedule(static)
    for (long r = 0; r < rows; r++) orc_inntt(c->plan[r % c->L], in + r * c->N, out + r * c->N);
}
/* op: 0 add, 1 sub, 2 mul (crt.jl:120-134) */
void orc_rns_binop(const rns_t *c, int op, const u64 *a, const u64 *b, u64 *out, long rows) {
#pragma omp parallel for schedule(static)
    for (long r = 0; r < rows; r++) {
        u64 q = c->q[r % c->L];
        const u64 *x = a + r * c->N, *y = b + r * c->N; u64 *o = out + r * c->N;
        for (u64 i = 0; i < c->N; i++)
            o[i] = op == 0 ? addmod(x[i], y[i], q) : op == 1 ? submod(x[i], y[i], q) : mulmod(x[i], y[i], q);
    }
}
void orc_rns_neg(const rns_t *c, const u64 *a, u64 *out, long rows) {
#pragma omp parallel for schedule(static)
    for (long r = 0; r < rows; r++) {
        u64 q = c->q[r % c->L];
        for (u64 i = 0; i < c->N; i++) { u64 v = a[r * c->N + i]; out[r * c->N + i] = v ? q - v : 0; }
    }
}
/* scalar given as its residues s[L] (pow2_cyc_rings.jl:177-185) */
void orc_rns_scalar_mul(const rns_t *c, const u64 *a, const u64 *s, u64 *out, long rows) {
#pragma omp parallel for schedule(static)
    for (long r = 0; r < rows; r++) {
        u64 q = c->q[r % c->L], sv = s[r % c->L];
        for (u64 i = 0; i < c->N; i++) out[r * c->N + i] = mulmod(a[r * c->N + i], sv, q);
    }
}
/* ring product of RNS polys: inntt(nntt(a) .* nntt(b)) (pow2_cyc_rings.jl:147-169) */
void orc_rns_ring_mul(const rns_t *c, const u64 *a, const u64 *b, u64 *out, long rows) {
#pragma omp parallel
    {
        u64 *ta = (u64 *)malloc(c->N * 8), *tb = (u64 *)malloc(c->N * 8);
#pragma omp for schedule(static)
        for (long r = 0; r < rows; r++) {
            const plan_t *p = c->plan[r % c->L];
            orc_nntt(p, a + r * c->N, ta); orc_nntt(p, b + r * c->N, tb);
            for (u64 i = 0; i < c->N; i++) ta[i] = mulmod(ta[i], tb[i], p->q);
            orc_inntt(p, ta, out + r * c->N);
        }
        free(ta); free(tb);
    }
}
/* apply_galois_element on primal rows (pow2_cyc_rings.jl:321-329) */
void orc_rns_galois(const rns_t *c, u64 g, const u64 *a, u64 *out, long rows) {
    const u64 N = c->N;
#pragma omp parallel for schedule(static)
    for (long r = 0; r < rows; r++) {
        u64 q = c->q[r % c->L];
        for (u64 i = 0; i < N; i++) {
            u64 gi = g * i, qq = gi / N, rr = gi % N, v = a[r * N + i];
            out[r * N + rr] = (qq & 1) ? (v ? q - v : 0) : v;
        }
    }
}
/* ciphertext tensor, no basis change (rlwe_she.jl:255-258):
 * c1,c2: [B][2][L][N] -> out [B][3][L][N] (primal in, primal out) */
void orc_ct_tensor(const rns_t *c, const u64 *c1, const u64 *c2, u64 *out, long batch) {
    const u64 N = c->N; const int L = c->L;
#pragma omp parallel
    {
        u64 *A0 = (u64 *)malloc(N * 8), *A1 = (u64 *)malloc(N * 8), *B0 = (u64 *)malloc(N * 8), *B1 = (u64 *)malloc(N * 8), *T = (u64 *)malloc(N * 8);
#pragma omp for schedule(static)
        for (long u = 0; u < batch * L; u++) {
            long b = u / L; int i = (int)(u % L);
            const plan_t *p = c->plan[i]; u64 q = p->q;
            orc_nntt(p, c1 + ((b * 2 + 0) * L + i) * N, A0); orc_nntt(p, c1 + ((b * 2 + 1) * L + i) * N, A1);
            orc_nntt(p, c2 + ((b * 2 + 0) * L + i) * N, B0); orc_nntt(p, c2 + ((b * 2 + 1) * L + i) * N, B1);
            for (u64 k = 0; k < N; k++) T[k] = mulmod(A0[k], B0[k], q);
            orc_inntt(p, T, out + ((b * 3 + 0) * L + i) * N);
            for (u64 k = 0; k < N; k++) T[k] = addmod(mulmod(A0[k], B1[k], q), mulmod(A1[k], B0[k], q), q);
            orc_inntt(p, T, out + ((b * 3 + 1) * L + i) * N);
            for (u64 k = 0; k < N; k++) T[k] = mulmod(A1[k], B1[k], q);
            orc_inntt(p, T, out + ((b * 3 + 2) * L + i) * N);
        }
        free(A0); free(A1); free(B0); free(B1); free(T);
    }
}
/* CKKS rescale (crt.jl:215-220): in [polys][L][N] -> out [polys][L-1][N] */
void orc_modswitch(const rns_t *c, const u64 *in, u64 *out, long polys) {
    const u64 N = c->N; const int L = c->L; const u64 qL = c->q[L - 1];
#pragma omp parallel for schedule(static)
    for (long u = 0; u < polys * (L - 1); u++) {
        long pidx = u / (L - 1); int i = (int)(u % (L - 1));
        u64 q = c->q[i], inv = invmod(qL % q, q);
        const u64 *ci = in + (pidx * L + i) * N, *cl = in + (pidx * L + L - 1) * N;
        u64 *o = out + (pidx * (L - 1) + i) * N;
        for (u64 k = 0; k < N; k++) o[k] = mulmod(inv, submod(ci[k], cl[k] % q, q), q);
    }
}
/* CRTExpand: multiply by P, append zero residue (crt.jl:35-40): in [polys][L][N] -> out [polys][L+1][N] */
void orc_crt_expand(const rns_t *c, u64 P, const u64 *in, u64 *out, long polys) {
    const u64 N = c->N; const int L = c->L;
#pragma omp parallel for schedule(static)
    for (long pidx = 0; pidx < polys; pidx++) {
        for (int i = 0; i < L; i++) {
            u64 q = c->q[i], pm = P % q;
            for (u64 k = 0; k < N; k++) out[(pidx * (L + 1) + i) * N + k] = mulmod(in[(pidx * L + i) * N + k], pm, q);
        }
        memset(out + (pidx * (L + 1) + L) * N, 0, N * 8);
    }
}

/* --------------------------------------------------------------- bigint */
#define BN 44 /* limbs: 2816 bits */
typedef struct { u64 w[BN]; int n; } big_t; /* n = used limbs (no leading zeros), value >= 0 */

static void big_set(big_t *a, u64 v) { memset(a->w, 0, sizeof(a->w)); a->w[0] = v; a->n = v ? 1 : 0; }
static void big_norm(big_t *a) { while (a->n > 0 && a->w[a->n - 1] == 0) a->n--; }
static int big_cmp(const big_t *a, const big_t *b) {
    if (a->n != b->n) return a->n < b->n ? -1 : 1;
    for (int i = a->n - 1; i >= 0; i--) if (a->w[i] != b->w[i]) return a->w[i] < b->w[i] ? -1 : 1;
    return 0;
}
static void big_mul_word(big_t *a, u64 m) { /* a *= m */
    u64 carry = 0;
    for (int i = 0; i < a->n; i++) { u128 t = (u128)a->w[i] * m + carry; a->w[i] = (u64)t; carry = (u64)(t >> 64); }
    if (carry) a->w[a->n++] = carry;
    big_norm(a);
}
static void big_addmul_word(big_t *a, const big_t *b, u64 m) { /* a += b*m */
    u64 carry = 0; int n = a->n > b->n ? a->n : b->n;
    for (int i = 0; i < n || carry; i++) {
        u128 t = (u128)(i < b->n ? b->w[i] : 0) * m + a->w[i] + carry;
        a->w[i] = (u64)t; carry = (u64)(t >> 64);
        if (i + 1 > a->n) a->n = i + 1;
    }
    if (n > a->n) a->n = n;
    big_norm(a);
}
static void big_add(big_t *a, const big_t *b) { big_addmul_word(a, b, 1); }
static void big_sub(big_t *a, const big_t *b) { /* a -= b, requires a >= b */
    u64 borrow = 0;
    for (int i = 0; i < a->n; i++) {
        u64 bi = i < b->n ? b->w[i] : 0, x = a->w[i], y = x - bi - borrow;
        borrow = (x < bi) || (x == bi && borrow) ? 1 : 0;
        a->w[i] = y;
    }
    big_norm(a);
}
static u64 big_mod_word(const big_t *a, u64 m) {
    u128 r = 0;
    for (int i = a->n - 1; i >= 0; i--) r = ((r << 64) | a->w[i]) % m;
    return (u64)r;
}
static void big_shr1(big_t *a) {
    for (int i = 0; i < a->n; i++) a->w[i] = (a->w[i] >> 1) | (i + 1 < a->n ? a->w[i + 1] << 63 : 0);
    big_norm(a);
}
/* Knuth algorithm D: q = floor(a / b), r = a mod b (b != 0) */
static void big_divrem(const big_t *a, const big_t *b, big_t *qo, big_t *ro) {
    big_set(qo, 0);
    if (big_cmp(a, b) < 0) { *ro = *a; return; }
    if (b->n == 1) {
        u128 r = 0; *qo = *a;
        for (int i = a->n - 1; i >= 0; i--) { u128 cur = (r << 64) | a->w[i]; qo->w[i] = (u64)(cur / b->w[0]); r = cur % b->w[0]; }
        big_norm(qo); big_set(ro, (u64)r); return;
    }
    int s = __builtin_clzll(b->w[b->n - 1]);
    int n = b->n, m = a->n - b->n;
    u64 vn[BN], un[BN + 1];
    for (int i = n - 1; i > 0; i--) vn[i] = s ? (b->w[i] << s) | (b->w[i - 1] >> (64 - s)) : b->w[i];
    vn[0] = b->w[0] << s;
    un[a->n] = s ? a->w[a->n - 1] >> (64 - s) : 0;
    for (int i = a->n - 1; i > 0; i--) un[i] = s ? (a->w[i] << s) | (a->w[i - 1] >> (64 - s)) : a->w[i];
    un[0] = a->w[0] << s;
    for (int j = m; j >= 0; j--) {
        u128 num = ((u128)un[j + n] << 64) | un[j + n - 1];
        u128 qhat = num / vn[n - 1], rhat = num % vn[n - 1];
        while ((qhat >> 64) || (u128)(u64)qhat * vn[n - 2] > ((rhat << 64) | un[j + n - 2])) {
            qhat--; rhat += vn[n - 1];
            if (rhat >> 64) break;
        }
        /* multiply and subtract */
        u64 borrow = 0, carry = 0;
        for (int i = 0; i < n; i++) {
            u128 pr = (u128)(u64)qhat * vn[i] + carry;
            carry = (u64)(pr >> 64);
            u64 sub = (u64)pr, x = un[i + j], y = x - sub - borrow;
            borrow = (x < sub) || (x == sub && borrow) ? 1 : 0;
            un[i + j] = y;
        }
        { u64 x = un[j + n], y = x - carry - borrow; borrow = (x < carry) || (x == carry && borrow) ? 1 : 0; un[j + n] = y; }
        u64 qd = (u64)qhat;
        if (borrow) { /* add back */
            qd--;
            u64 c2 = 0;
            for (int i = 0; i < n; i++) { u128 t = (u128)un[i + j] + vn[i] + c2; un[i + j] = (u64)t; c2 = (u64)(t >> 64); }
            un[j + n] += c2;
        }
        qo->w[j] = qd;
    }
    qo->n = m + 1; big_norm(qo);
    memset(ro->w, 0, sizeof(ro->w));
    for (int i = 0; i < n; i++) ro->w[i] = s ? (un[i] >> s) | (un[i + 1] << (64 - s)) : un[i];
    ro->n = n; big_norm(ro);
}

/* CRT basis helper: incremental reconstruction X in [0,Q) (crt.jl:105-112) */
typedef struct {
    int L; u64 q[BN];
    big_t M[BN];      /* M[i] = prod_{j<i} q_j */
    u64 Minv[BN];     /* (M[i] mod q_i)^-1 mod q_i */
    big_t Q, halfQ;   /* Q, Q>>1 */
} basis_t;

static void basis_init(basis_t *b, int L, const u64 *q) {
    b->L = L; big_t M; big_set(&M, 1);
    for (int i = 0; i < L; i++) {
        b->q[i] = q[i]; b->M[i] = M;
        b->Minv[i] = invmod(big_mod_word(&M, q[i]), q[i]);
        big_mul_word(&M, q[i]);
    }
    b->Q = M; b->halfQ = M; big_shr1(&b->halfQ);
}
static void basis_reconstruct(const basis_t *b, const u64 *res, big_t *X) {
    big_set(X, res[0]);
    for (int i = 1; i < b->L; i++) {
        u64 q = b->q[i], xm = big_mod_word(X, q);
        u64 t = mulmod(submod(res[i] % q, xm, q), b->Minv[i], q);
        big_addmul_word(X, &b->M[i], t);
    }
}
/* centred lift (signedmod.jl:12-19 / bfv.jl:202-220): returns sign (1 = negative), X <- |X'| */
static int basis_centre(const basis_t *b, big_t *X) {
    if (big_cmp(X, &b->halfQ) > 0) { big_t t = b->Q; big_sub(&t, X); *X = t; return 1; }
    return 0;
}

/* switch (bfv.jl:202-226): in [polys][Lf][N] (basis qf) -> out [polys][Lt][N] (basis qt) */
void orc_bfv_switch(u64 N, int Lf, const u64 *qf, int Lt, const u64 *qt, const u64 *in, u64 *out, long polys) {
    basis_t *bf = (basis_t *)malloc(sizeof(basis_t)); basis_init(bf, Lf, qf);
#pragma omp parallel for schedule(static)
    for (long u = 0; u < polys * (long)N; u++) {
        long p = u / N; u64 k = u % N;
        u64 res[BN]; big_t X;
        for (int i = 0; i < Lf; i++) res[i] = in[(p * Lf + i) * N + k];
        basis_reconstruct(bf, res, &X);
        int neg = basis_centre(bf, &X);
        for (int j = 0; j < Lt; j++) {
            u64 r = big_mod_word(&X, qt[j]);
            out[(p * Lt + j) * N + k] = (neg && r) ? qt[j] - r : r;
        }
    }
    free(bf);
}

/* mul_contract (bfv.jl:35-40): switch(R, multround(e, t, Q)) per coefficient, in [polys][Lb][N] over qb.
 * Transcribes, line by line:
 *   multround(SignedMod(x), t, Q) = div(e * t, Q, RoundNearestTiesAway)          bfv.jl:172-174, 188
 *   e * t = SignedMod{T}(e.x * T(t))   -- the product is taken IN THE CRT FIELD  signedmod.jl:24-28
 *           (residue-wise modulo every prime of R_big, i.e. modulo Q_big), only then
 *   convert(Integer, e) = centred lift modulo Q_big                              signedmod.jl:12-19
 *   div(.., Q, RoundNearestTiesAway), oftype(e, y) = y mod Q_big                 signedmod.jl:30-32, div_hacks.jl:120-135
 *   switch(R, .) = centred lift modulo Q_big again, then mod q_i                 bfv.jl:202-226
 * It equals rha(t*centre(x), Q) only while t*|x| < Q_big/2; beyond that the reference wraps modulo Q_big
 * (it does on test/bfv_crt.jl's 2+4-prime ring) and so does this.
 * out [polys][L][N] = y mod q_i */
void orc_bfv_contract(u64 N, int L, const u64 *q, int Lb, const u64 *qb, u64 t, const u64 *in, u64 *out, long polys) {
    basis_t *bq = (basis_t *)malloc(sizeof(basis_t)), *bb = (basis_t *)malloc(sizeof(basis_t));
    basis_init(bq, L, q); basis_init(bb, Lb, qb);
#pragma omp parallel for schedule(static)
    for (long u = 0; u < polys * (long)N; u++) {
        long p = u / N; u64 k = u % N;
        u64 res[BN]; big_t X, Y, R, twoR;
        for (int j = 0; j < Lb; j++) res[j] = mulmod(in[(p * Lb + j) * N + k] % qb[j], t % qb[j], qb[j]);   /* e.x * T(t) */
        basis_reconstruct(bb, res, &X);
        int neg = basis_centre(bb, &X);
        big_divrem(&X, &bq->Q, &Y, &R);
        /* RoundNearestTiesAway (div_hacks.jl:120-135): |r| >= Q/2 rounds away */
        twoR = R; big_add(&twoR, &R);
        if (big_cmp(&twoR, &bq->Q) >= 0) { big_t one; big_set(&one, 1); big_add(&Y, &one); }
        /* |y| <= Q_big/(2Q) + 1 < Q_big/2, so the second centred lift returns y itself and the final
         * residues are y mod q_i */
        for (int i = 0; i < L; i++) {
            u64 r = big_mod_word(&Y, q[i]);
            out[(p * L + i) * N + k] = (neg && r) ? q[i] - r : r;
        }
    }
    free(bq); free(bb);
}

/* BFV ciphertext multiply (rlwe_she.jl:247-262 with bfv.jl:34-40):
 * c1,c2 [B][2][L][N] -> out [B][3][L][N] */
void orc_bfv_mul(const rns_t *cq, const rns_t *cb, u64 t, const u64 *c1, const u64 *c2, u64 *out, long batch) {
    const u64 N = cq->N; const int L = cq->L, Lb = cb->L;
    u64 *e1 = (u64 *)malloc((size_t)batch * 2 * Lb * N * 8), *e2 = (u64 *)malloc((size_t)batch * 2 * Lb * N * 8);
    u64 *tz = (u64 *)malloc((size_t)batch * 3 * Lb * N * 8);
    orc_bfv_switch(N, L, cq->q, Lb, cb->q, c1, e1, batch * 2);
    orc_bfv_switch(N, L, cq->q, Lb, cb->q, c2, e2, batch * 2);
    orc_ct_tensor(cb, e1, e2, tz, batch);
    orc_bfv_contract(N, L, cq->q, Lb, cb->q, t, tz, out, batch * 3);
    free(e1); free(e2); free(tz);
}

/* number of base-2^w digits of Q (rlwe_she.jl:333) */
int orc_ndigits(int L, const u64 *q, int w) {
    basis_t *b = (basis_t *)malloc(sizeof(basis_t)); basis_init(b, L, q);
    int bits = (b->Q.n - 1) * 64 + (64 - __builtin_clzll(b->Q.w[b->Q.n - 1]));
    free(b);
    return (bits + w - 1) / w;
}

/* keyswitch digit polys (rlwe_she.jl:326-338).
 * cend [L][N] over q; digits embedded over target basis qt [Lt]; out [D][Lt][N].
 * w == 0: CRT digits (D = L), centred residue; else base-2^w digits of X in [0,Q). */
void orc_keyswitch_digits(u64 N, int L, const u64 *q, int Lt, const u64 *qt, int w, const u64 *cend, u64 *out) {
    if (w == 0) {
#pragma omp parallel for schedule(static)
        for (long u = 0; u < (long)L * Lt; u++) {
            int i = (int)(u / Lt), j = (int)(u % Lt);
            u64 qi = q[i], half = qi / 2, p = qt[j];
            for (u64 k = 0; k < N; k++) {
                u64 c = cend[(u64)i * N + k];
                u64 o;
                if (c > half) { u64 m = (qi - c) % p; o = m ? p - m : 0; } else o = c % p;
                out[((u64)i * Lt + j) * N + k] = o;
            }
        }
        return;
    }
    basis_t *b = (basis_t *)malloc(sizeof(basis_t)); basis_init(b, L, q);
    int D = orc_ndigits(L, q, w);
#pragma omp parallel for schedule(static)
    for (long k = 0; k < (long)N; k++) {
        u64 res[BN]; big_t X;
        for (int i = 0; i < L; i++) res[i] = cend[(u64)i * N + k];
        basis_reconstruct(b, res, &X);
        for (int d = 0; d < D; d++) {
            int bit = d * w, limb = bit / 64, off = bit % 64;
            u64 v = limb < BN ? X.w[limb] >> off : 0;
            if (off + w > 64 && limb + 1 < BN) v |= X.w[limb + 1] << (64 - off);
            v &= (w == 64) ? ~0ull : ((1ull << w) - 1);
            for (int j = 0; j < Lt; j++) out[((u64)d * Lt + j) * N + k] = v % qt[j];
        }
    }
    free(b);
}

/* keyswitch accumulation (rlwe_she.jl:340-344): c1 += masked_d * p_d, c2 += mask_d * p_d
 * c1,c2 [L][N] primal in/out; digits [D][L][N] primal; key [D][2][Lk][N] primal
 * (component 0 = mask, 1 = masked), key basis may have Lk >= L rows; ``which``[L]
 * selects the key residue row for each ciphertext row (downswitch_keyelement). */
void orc_keyswitch_accum(const rns_t *c, int D, int Lk, const int *which, const u64 *digits, const u64 *key, u64 *c1, u64 *c2) {
    const u64 N = c->N; const int L = c->L;
#pragma omp parallel
    {
        u64 *P = (u64 *)malloc(N * 8), *K = (u64 *)malloc(N * 8), *A1 = (u64 *)malloc(N * 8), *A2 = (u64 *)malloc(N * 8);
#pragma omp for schedule(static)
        for (int i = 0; i < L; i++) {
            const plan_t *p = c->plan[i]; u64 q = p->q;
            memset(A1, 0, N * 8); memset(A2, 0, N * 8);
            for (int d = 0; d < D; d++) {
                orc_nntt(p, digits + ((u64)d * L + i) * N, P);
                orc_nntt(p, key + (((u64)d * 2 + 0) * Lk + which[i]) * N, K);
                for (u64 k = 0; k < N; k++) A2[k] = addmod(A2[k], mulmod(K[k], P[k], q), q);
                orc_nntt(p, key + (((u64)d * 2 + 1) * Lk + which[i]) * N, K);
                for (u64 k = 0; k < N; k++) A1[k] = addmod(A1[k], mulmod(K[k], P[k], q), q);
            }
            orc_inntt(p, A1, A1); orc_inntt(p, A2, A2);
            for (u64 k = 0; k < N; k++) { c1[(u64)i * N + k] = addmod(c1[(u64)i * N + k], A1[k], q); c2[(u64)i * N + k] = addmod(c2[(u64)i * N + k], A2[k], q); }
        }
        free(P); free(K); free(A1); free(A2);
    }
}

/* exact CRT reconstruction to little-endian limbs, for tests: in [L][N] -> out [N][nl] */
void orc_rns_to_limbs(u64 N, int L, const u64 *q, const u64 *in, u64 *out, int nl) {
    basis_t *b = (basis_t *)malloc(sizeof(basis_t)); basis_init(b, L, q);
    for (u64 k = 0; k < N; k++) {
        u64 res[BN]; big_t X;
        for (int i = 0; i < L; i++) res[i] = in[(u64)i * N + k];
        basis_reconstruct(b, res, &X);
        for (int j = 0; j < nl; j++) out[k * nl + j] = j < BN ? X.w[j] : 0;
    }
    free(b);
}
