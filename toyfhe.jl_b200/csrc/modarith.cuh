// Word-size modular arithmetic for the sm_100a negacyclic-NTT engine.
//
// Replaces GaloisFields.jl PrimeField `* + - inv ^` (widemul + rem) on the hot
// path of pow2_cyc_rings.jl / crt.jl with Shoup (precomputed-quotient) and
// Barrett multiplication.  All residues are canonical in [0,q) at kernel
// boundaries; inside the butterfly ladders values are lazily kept in [0,4q)
// (Harvey), which needs q < 2^62.
#pragma once
#include <cstdint>

typedef uint64_t u64;
typedef uint32_t u32;
typedef unsigned __int128 u128;

#define TFB_MAX_L 64  // max RNS primes per context

#ifndef __CUDACC__
#define __align__(n) __attribute__((aligned(n)))
#endif

// (w, w') pair: w' = floor(w * 2^64 / q)
struct __align__(16) tw_t {
    u64 w, wp;
};

struct PrimeConst {
    u64 q;        // modulus
    u64 q2;       // 2q
    u64 br_hi;    // floor(2^128 / q) high word
    u64 br_lo;    // floor(2^128 / q) low word
    u64 c64;      // 2^64 mod q
    u64 c64p;     // Shoup companion floor(c64 * 2^64 / q)
};

#ifdef __CUDACC__
#define TFB_HD __host__ __device__ __forceinline__
#define TFB_D __host__ __device__ __forceinline__
#else
#define TFB_HD inline
#define TFB_D inline
#endif

// high 64 bits of a 64x64 product (device: IMAD.WIDE ladder; host: __int128 --
// the host path exists only so tests can emulate the kernels' index logic on CPU)
TFB_D u64 mulhi64(u64 a, u64 b) {
#ifdef __CUDA_ARCH__
    return __umul64hi(a, b);
#else
    return (u64)(((u128)a * b) >> 64);
#endif
}

// x*w mod q, lazily in [0,2q); valid for ANY 64-bit x (Harvey/Shoup).
// Device form: r = lo64(x*w + h*(-q)) as ONE multiply-accumulate chain over 32-bit
// limbs (2 IMAD.WIDE + 4 IMAD, no separate adds) -- measured 41 -> 36 SM cycles per
// warp-butterfly against the plain C expression (profiles/bfly_bench3_r1.txt).
TFB_D u64 shoup_lazy(u64 x, u64 w, u64 wp, u64 q) {
    u64 h = mulhi64(x, wp);
#ifdef __CUDA_ARCH__
    const u64 nq = 0 - q;
    u32 x0, x1, w0, w1, h0, h1, n0, n1, lo, hi;
    u64 acc;
    asm("mov.b64 {%0,%1}, %2;" : "=r"(x0), "=r"(x1) : "l"(x));
    asm("mov.b64 {%0,%1}, %2;" : "=r"(w0), "=r"(w1) : "l"(w));
    asm("mov.b64 {%0,%1}, %2;" : "=r"(h0), "=r"(h1) : "l"(h));
    asm("mov.b64 {%0,%1}, %2;" : "=r"(n0), "=r"(n1) : "l"(nq));
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(acc) : "r"(x0), "r"(w0));
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(h0), "r"(n0));
    asm("mov.b64 {%0,%1}, %2;" : "=r"(lo), "=r"(hi) : "l"(acc));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(x0), "r"(w1));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(x1), "r"(w0));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(h0), "r"(n1));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(h1), "r"(n0));
    asm("mov.b64 %0, {%1,%2};" : "=l"(acc) : "r"(lo), "r"(hi));
    return acc;
#else
    return x * w - h * q;
#endif
}
TFB_D u64 shoup_lazy(u64 x, tw_t t, u64 q) { return shoup_lazy(x, t.w, t.wp, q); }

// x*w mod q in [0,4q) for primes q = 2^b + e with 32 <= b and e < 2^32 (what the reference's
// nextprime(2^logq + 1) chains give, crt.jl:282-295); valid for ANY 64-bit x.
//   quotient: h~ = x1*p1 + floor((x1*p0 + x0*p1) / 2^32)  in [h-1, h]: the two cross products are summed as ONE 64-bit
//             multiply-accumulate chain whose carry joins the high word (1 IMAD.WIDE + 1 IMAD.HI with addend and carry
//             out for the cross terms, 1 IMAD.WIDE for x1*p1) -- the round-1 form took the two high halves separately
//             (2 IMAD.HI) and ptxas re-materialised a {0, t0} register pair per product (1 IMAD.MOV + 1 MOV each);
//             SASS per butterfly: 28.9 -> 27.1 FMA-heavy cycles, -1.8 % kernel time (tools/ntt_lab.cu, ABL bit 128);
//   tail:     x*w - h~*q = x*w - h~*e - (h~ << b)  (mod 2^64): one IMAD.WIDE less than a generic q.
// ne = 2^32 - e, shb = b - 32 (a run-time value: one shift and one 3-input add on the ALU pipe; as a
// compile-time constant ptxas turns it into a multiply on the FMA-heavy pipe, which bounds the kernels).
// Measured (IMAD.WIDE/IMAD.HI 4 cycles, IMAD 2 cycles per warp instruction, tools/bfly_bench4.cu).
TFB_D u64 shoup_lazy4(u64 x, u64 w, u64 wp, u64 q, u32 ne, u32 shb) {
#ifdef __CUDA_ARCH__
    u32 x0, x1, w0, w1, p0, p1, u0, u1, m1, c, t0, t1, h0, h1, lo, hi;
    u64 u, t, acc;
    asm("mov.b64 {%0,%1}, %2;" : "=r"(x0), "=r"(x1) : "l"(x));
    asm("mov.b64 {%0,%1}, %2;" : "=r"(w0), "=r"(w1) : "l"(w));
    asm("mov.b64 {%0,%1}, %2;" : "=r"(p0), "=r"(p1) : "l"(wp));
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(u) : "r"(x1), "r"(p0));
    asm("mov.b64 {%0,%1}, %2;" : "=r"(u0), "=r"(u1) : "l"(u));
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(x1), "r"(p1));
    asm("mov.b64 {%0,%1}, %2;" : "=r"(t0), "=r"(t1) : "l"(t));
    // m1:c = (x0*p1 + u) >> 32 with its carry; h = t + m1 + (c << 32)
    asm("{\n\t.reg .u32 d;\n\tmad.lo.cc.u32 d, %4, %5, %6;\n\tmadc.hi.cc.u32 %0, %4, %5, %7;\n\taddc.u32 %1, 0, 0;\n\t"
        "add.cc.u32 %2, %8, %0;\n\taddc.u32 %3, %9, %1;\n\t}"
        : "=&r"(m1), "=&r"(c), "=&r"(h0), "=&r"(h1) : "r"(x0), "r"(p1), "r"(u0), "r"(u1), "r"(t0), "r"(t1));
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(acc) : "r"(h0), "r"(ne));
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(x0), "r"(w0));
    asm("mov.b64 {%0,%1}, %2;" : "=r"(lo), "=r"(hi) : "l"(acc));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(x0), "r"(w1));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(x1), "r"(w0));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(h1), "r"(ne));
    hi = hi - h0 - (h0 << shb);
    asm("mov.b64 %0, {%1,%2};" : "=l"(acc) : "r"(lo), "r"(hi));
    return acc;
#else
    // host mirror (tests/emu): the same quotient estimate, so the lazy ranges checked on the CPU are the device's
    const u64 x1 = x >> 32, x0 = x & 0xffffffffu, p1 = wp >> 32, p0 = wp & 0xffffffffu;
    const u128 mid = (u128)x1 * p0 + (u128)x0 * p1;
    const u64 h = x1 * p1 + (u64)(mid >> 32);
    (void)ne; (void)shb;
    return x * w - h * q;
#endif
}

// conditional subtract: x in [0,2m) -> [0,m)
TFB_D u64 csub(u64 x, u64 m) { return x >= m ? x - m : x; }

TFB_D u64 shoup_full(u64 x, u64 w, u64 wp, u64 q) { return csub(shoup_lazy(x, w, wp, q), q); }

// a*b mod q for arbitrary a,b < 2^64 with a*b < q*2^64 (always true for a,b<q<2^63):
// Barrett with ratio = floor(2^128/q) (two words), result canonical.
TFB_D u64 barrett_mul(u64 a, u64 b, const PrimeConst& pc) {
    u64 z0 = a * b, z1 = mulhi64(a, b);
    // estimate floor(z * ratio / 2^128), low 64 bits only
    u64 c = mulhi64(z0, pc.br_lo);
    u64 t0 = z0 * pc.br_hi, t1 = mulhi64(z0, pc.br_hi);
    u64 s = t0 + c;
    u64 carry = s < t0;
    u64 r1 = t1 + carry;
    u64 u0 = z1 * pc.br_lo, u1 = mulhi64(z1, pc.br_lo);
    u64 s2 = s + u0;
    u64 carry2 = s2 < u0;
    u64 qhat = z1 * pc.br_hi + r1 + u1 + carry2;
    u64 r = z0 - qhat * pc.q;
    // qhat underestimates by at most 2
    r = csub(r, pc.q2);
    return csub(r, pc.q);
}

// reduce a 128-bit value (hi,lo) < q * 2^64 modulo q (canonical)
TFB_D u64 barrett_red128(u64 z1, u64 z0, const PrimeConst& pc) {
    u64 c = mulhi64(z0, pc.br_lo);
    u64 t0 = z0 * pc.br_hi, t1 = mulhi64(z0, pc.br_hi);
    u64 s = t0 + c;
    u64 carry = s < t0;
    u64 r1 = t1 + carry;
    u64 u0 = z1 * pc.br_lo, u1 = mulhi64(z1, pc.br_lo);
    u64 s2 = s + u0;
    u64 carry2 = s2 < u0;
    u64 qhat = z1 * pc.br_hi + r1 + u1 + carry2;
    u64 r = z0 - qhat * pc.q;
    r = csub(r, pc.q2);
    return csub(r, pc.q);
}

// x mod q for a single word x (canonical)
TFB_D u64 barrett_red64(u64 x, const PrimeConst& pc) {
    u64 qhat = mulhi64(x, pc.br_hi);  // floor(x*ratio/2^128) ~ hi(x * br_hi) (under by <=2)
    u64 r = x - qhat * pc.q;
    r = csub(r, pc.q2);
    return csub(r, pc.q);
}

// (z1 * 2^64 + z0) mod q for ANY 128-bit value (needs q < 2^62): the high word goes
// through a Shoup product with the constant 2^64 mod q, the low word through a
// one-word Barrett step; canonical result.
TFB_D u64 red128_any(u64 z1, u64 z0, const PrimeConst& pc) {
    const u64 s = shoup_lazy(z1, pc.c64, pc.c64p, pc.q);       // [0,2q)
    u64 r0 = z0 - mulhi64(z0, pc.br_hi) * pc.q;                // [0,3q)
    r0 = csub(r0, pc.q2);                                      // [0,2q)
    u64 t = s + r0;                                            // [0,4q)
    t = csub(t, pc.q2);
    return csub(t, pc.q);
}

// (z1 2^64 + z0) mod q, canonical, for q = 2^60 + e (e < 2^28) and z < 2^126: two Solinas folds at bit 60
// (z = lo - (z >> 60) e, twice) -- 3 IMAD.WIDE + 3 IMAD instead of the Shoup product plus Barrett step of red128_any
// (9 + 6); used where every prime of a ring has this shape (the reference's nextprime(2^60+1) chains, crt.jl:282-295).
TFB_D u64 red126_sp60(const u64 z1, const u64 z0, const u64 q, const u32 e) {
#ifdef __CUDA_ARCH__
    const u64 M60 = (1ull << 60) - 1;
    const u64 H0 = (z0 >> 60) | (z1 << 4);              // H = z >> 60 = H1 2^64 + H0, H1 < 4
    const u32 H1 = (u32)(z1 >> 60);
    u64 t0, t1, u0;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(t0) : "r"((u32)H0), "r"(e));
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(t1) : "r"((u32)(H0 >> 32)), "r"(e));
    const u64 mid = t1 + (t0 >> 32);                    // T = H e = Th 2^64 + Tl < 2^94
    const u64 Tl = (mid << 32) | (u32)t0;
    const u64 Th = (mid >> 32) + (u64)(H1 * e);
    const u64 Thi = (Tl >> 60) | (Th << 4);             // T >> 60 < 2^34
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(u0) : "r"((u32)Thi), "r"(e));
    const u64 U = u0 + ((u64)((u32)(Thi >> 32) * e) << 32);   // (T >> 60) e < 2^62
    u64 r = (z0 & M60) + (q - (Tl & M60)) + U;          // = z (mod q), in (0, 6q)
    r = (r & M60) + q - (u64)((u32)(r >> 60) * e);      // in (0, 2q)
    return csub(r, q);
#else
    (void)e;
    return (u64)((((u128)z1 << 64) | z0) % q);
#endif
}

// (hi:lo) += x y  and  (hi:lo) = x y  for x, y < 2^62 (residues and residue-sized constants: every modulus is below 2^62).
// With 30-bit upper halves the two cross products add up in 64 bits without a carry, so the 128-bit product is four
// 32x32->64 multiplies (one of them a multiply-add) and the accumulation seven 32-bit adds; the compiler's mul.lo + mul.hi
// pair computes the low partial products twice (about 24 vs 16 cycles of the integer-multiply pipe per product).
TFB_D void mac_wide62(u64& lo, u64& hi, const u64 x, const u64 y) {
#ifdef __CUDA_ARCH__
    asm("{\n\t"
        ".reg .u32 x0, x1, y0, y1, l0, l1, m0, m1, h0, h1, a0, a1, a2, a3;\n\t"
        ".reg .u64 l, m, h;\n\t"
        "mov.b64 {x0, x1}, %2;\n\t"
        "mov.b64 {y0, y1}, %3;\n\t"
        "mov.b64 {a0, a1}, %0;\n\t"
        "mov.b64 {a2, a3}, %1;\n\t"
        "mul.wide.u32 l, x0, y0;\n\t"
        "mul.wide.u32 m, x0, y1;\n\t"
        "mad.wide.u32 m, x1, y0, m;\n\t"
        "mul.wide.u32 h, x1, y1;\n\t"
        "mov.b64 {l0, l1}, l;\n\t"
        "mov.b64 {m0, m1}, m;\n\t"
        "mov.b64 {h0, h1}, h;\n\t"
        "add.cc.u32 a0, a0, l0;\n\t"
        "addc.cc.u32 a1, a1, l1;\n\t"
        "addc.cc.u32 a2, a2, h0;\n\t"
        "addc.u32 a3, a3, h1;\n\t"
        "add.cc.u32 a1, a1, m0;\n\t"
        "addc.cc.u32 a2, a2, m1;\n\t"
        "addc.u32 a3, a3, 0;\n\t"
        "mov.b64 %0, {a0, a1};\n\t"
        "mov.b64 %1, {a2, a3};\n\t"
        "}"
        : "+l"(lo), "+l"(hi)
        : "l"(x), "l"(y));
#else
    const u128 z = (((u128)hi << 64) | lo) + (u128)x * y;
    lo = (u64)z;
    hi = (u64)(z >> 64);
#endif
}
TFB_D void mul_wide62(u64& lo, u64& hi, const u64 x, const u64 y) {
#ifdef __CUDA_ARCH__
    asm("{\n\t"
        ".reg .u32 x0, x1, y0, y1, l0, l1, m0, m1, h0, h1;\n\t"
        ".reg .u64 l, m, h;\n\t"
        "mov.b64 {x0, x1}, %2;\n\t"
        "mov.b64 {y0, y1}, %3;\n\t"
        "mul.wide.u32 l, x0, y0;\n\t"
        "mul.wide.u32 m, x0, y1;\n\t"
        "mad.wide.u32 m, x1, y0, m;\n\t"
        "mul.wide.u32 h, x1, y1;\n\t"
        "mov.b64 {l0, l1}, l;\n\t"
        "mov.b64 {m0, m1}, m;\n\t"
        "mov.b64 {h0, h1}, h;\n\t"
        "add.cc.u32 l1, l1, m0;\n\t"
        "addc.cc.u32 h0, h0, m1;\n\t"
        "addc.u32 h1, h1, 0;\n\t"
        "mov.b64 %0, {l0, l1};\n\t"
        "mov.b64 %1, {h0, h1};\n\t"
        "}"
        : "=l"(lo), "=l"(hi)
        : "l"(x), "l"(y));
#else
    const u128 z = (u128)x * y;
    lo = (u64)z;
    hi = (u64)(z >> 64);
#endif
}

TFB_D u64 add_mod(u64 a, u64 b, u64 q) { return csub(a + b, q); }
TFB_D u64 sub_mod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }
TFB_D u64 neg_mod(u64 a, u64 q) { return a ? q - a : 0; }

// ----------------------------------------------------------------- host side
static inline u64 h_mulmod(u64 a, u64 b, u64 q) { return (u64)((u128)a * b % q); }
static inline u64 h_powmod(u64 a, u64 e, u64 q) {
    u64 r = 1 % q;
    a %= q;
    while (e) {
        if (e & 1) r = h_mulmod(r, a, q);
        a = h_mulmod(a, a, q);
        e >>= 1;
    }
    return r;
}
static inline u64 h_invmod(u64 a, u64 q) { return h_powmod(a, q - 2, q); }
static inline u64 h_shoup(u64 w, u64 q) { return (u64)(((u128)w << 64) / q); }
static inline tw_t h_tw(u64 w, u64 q) {
    tw_t t;
    t.w = w;
    t.wp = h_shoup(w, q);
    return t;
}
static inline PrimeConst h_prime_const(u64 q) {
    PrimeConst pc;
    pc.q = q;
    pc.q2 = 2 * q;
    // floor(2^128 / q): long division of 2^128 by q
    u128 hi = ((u128)1 << 64) / q;            // floor(2^64 / q)  (fits: q >= 2)
    u128 rem = ((u128)1 << 64) % q;
    u128 lo = (rem << 64) / q;
    pc.br_hi = (u64)hi;
    pc.br_lo = (u64)lo;
    pc.c64 = (u64)(((u128)1 << 64) % q);
    pc.c64p = h_shoup(pc.c64, q);
    return pc;
}
