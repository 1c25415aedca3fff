// TEST-ONLY: runs the per-thread phase bodies of the CUDA NTT kernels
// (toyfhe.jl_b200/csrc/ntt_core.cuh) thread-by-thread on the CPU, so the index
// logic (thread mapping, swizzle, twiddle indices, natural-order store) can be
// checked against the oracle in a container without a GPU.  Never shipped.
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../toyfhe.jl_b200/csrc/ntt_core.cuh"
#include "../../toyfhe.jl_b200/csrc/ntt_core3.cuh"
#include "../../toyfhe.jl_b200/csrc/tables.h"

static u32 log2floor(u64 q) { return 63 - (u32)__builtin_clzll(q); }

template <int R, int MODE>
static void run(int inverse, u64 q, u64 psi, u32 s0, const u64* in, u64* out) {
    u64 tab[16];
    for (int k = 0; k < 16; k++) tab[k] = q - (u64)k * (q - (1ull << log2floor(q)));
    const RedParams rp = MODE == 2 ? make_red2(q, log2floor(q), tab) : make_red(q, log2floor(q));
    typedef NttGeo<R> Geo;
    const u64 Nrow = (u64)Geo::N << s0;
    HostTables ht;
    build_tables(Nrow, q, psi, ht);
    std::vector<u64> smem(Geo::N), regs((size_t)Geo::T * 32);
    std::vector<tw_t> fwdc(Nrow), invc(Nrow);   // thread-order pass-3 copies, as tfb_ctx_create builds them
    permute_pass3(ht.fwd.data(), fwdc.data(), 10 + R + (int)s0);
    permute_pass3(ht.inv.data(), invc.data(), 10 + R + (int)s0);
    for (u32 blk = 0; blk < (1u << s0); blk++) {
        if (!inverse) {
            for (u32 t = 0; t < Geo::T; t++) fwd_phaseA<R, MODE>(&regs[t * 32], in + (u64)blk * Geo::N, smem.data(), ht.fwd.data(), rp, t, s0, blk);
            for (u32 t = 0; t < Geo::T; t++) fwd_phaseB<R, MODE>(&regs[t * 32], smem.data(), ht.fwd.data(), rp, t, s0, blk);
            for (u32 t = 0; t < Geo::T; t++) fwd_phaseC<R, MODE>(&regs[t * 32], out, smem.data(), fwdc.data(), rp, t, s0, blk);
        } else {
            for (u32 t = 0; t < Geo::T; t++) inv_phaseC<R>(&regs[t * 32], in, smem.data(), invc.data(), q, t, s0, blk);
            for (u32 t = 0; t < Geo::T; t++) inv_phaseB<R>(&regs[t * 32], smem.data(), ht.inv.data(), q, t, s0, blk);
            for (u32 t = 0; t < Geo::T; t++) inv_phaseA<R>(&regs[t * 32], out + (u64)blk * Geo::N, smem.data(), ht.inv.data(), q, t, s0, blk, ht.ninv, ht.ninv_w1);
        }
    }
}

// emulates the fast kernel on one row of length 2^(10+R+s0).  For s0>0 only the
// row-resident part is emulated: forward expects stages 1..s0 already applied to
// `in`; inverse leaves stages s0..1 (and the N^-1 scale) to the caller.
extern "C" int emu_ntt(int R, int mode, int inverse, u64 q, u64 psi, u32 s0, const u64* in, u64* out) {
#define RUN(RR) case RR: if (mode == 2) run<RR, 2>(inverse, q, psi, s0, in, out); else if (mode) run<RR, 1>(inverse, q, psi, s0, in, out); else run<RR, 0>(inverse, q, psi, s0, in, out); break;
    switch (R) {
        RUN(0) RUN(1) RUN(2) RUN(3) RUN(4)
        default: return 1;
    }
#undef RUN
    return 0;
}

// number of lazy-range violations seen by the MODE 2 butterflies since the last call (must stay 0)
extern "C" unsigned long long emu_overflow_count() {
    const unsigned long long v = g_emu_overflow;
    g_emu_overflow = 0;
    return v;
}

// bank-conflict census of the three shared-memory access patterns (8-byte words,
// 16 lanes per wavefront): returns the worst number of lanes of a half-warp that
// fall on the same 8-byte bank.
template <int R>
static int conflicts() {
    typedef NttGeo<R> Geo;
    int worst = 1;
    auto census = [&](u32* addr) {
        for (int h = 0; h < 2; h++) {
            int cnt[16] = {0};
            for (int l = 0; l < 16; l++) cnt[addr[h * 16 + l] % 16]++;
            for (int i = 0; i < 16; i++) worst = cnt[i] > worst ? cnt[i] : worst;
        }
    };
    u32 addr[32];
    for (u32 wbase = 0; wbase < Geo::T; wbase += 32) {
        for (u32 r = 0; r < 32; r++) {
            for (u32 l = 0; l < 32; l++) addr[l] = swz<R>(r, wbase + l);  // phase A write
            census(addr);
            for (u32 l = 0; l < 32; l++) { u32 t = wbase + l; addr[l] = swz<R>(t >> R, r * Geo::RS + (t & (Geo::RS - 1))); }  // phase B
            census(addr);
        }
        u32 w = wbase >> 5;
        for (u32 g = 0; g < Geo::G; g++)
            for (u32 c = 0; c < Geo::RS; c++) {
                for (u32 l = 0; l < 32; l++) addr[l] = swz<R>(brev_bits(l, 5), brev_bits(w * Geo::G + g, 5) * Geo::RS + c);
                census(addr);
            }
    }
    return worst;
}
extern "C" int emu_bank_conflicts(int R) {
    switch (R) {
        case 0: return conflicts<0>();
        case 1: return conflicts<1>();
        case 2: return conflicts<2>();
        case 3: return conflicts<3>();
        case 4: return conflicts<4>();
    }
    return -1;
}

// ---- third-generation kernels (ntt_core3.cuh): skewed row buffer, approximate quotient, fused reductions.
// Return the number of lazy-range violations (must be 0), -1 if the prime is not eligible.
static u32 floor_log2(u64 q) { u32 b = 0; while (b < 63 && (q >> (b + 1))) b++; return b; }
template <int R>
static long long emu3_fwd(u64 q, u64 psi, u32 s0, const u64* in, u64* out) {
    using namespace v3;
    typedef NttGeo<R> Geo;
    if (!prime_ok(q)) return -1;
    redent_t tab[16];
    fill_redtab(tab, q);
    const Red3 rp = make_red3(q, floor_log2(q), tab);
    const u64 Nrow = (u64)Geo::N << s0;
    HostTables ht;
    build_tables(Nrow, q, psi, ht);
    std::vector<tw_t> fwdc(Nrow);
    permute_pass3(ht.fwd.data(), fwdc.data(), 10 + R + (int)s0);
    std::vector<u64> smem(Lay<R>::ROW_WORDS), regs((size_t)Geo::T * 32);
    g_emu_overflow3 = 0;
    for (u32 blk = 0; blk < (1u << s0); blk++) {
        for (u32 a = 0; a < 32; a++) memcpy(&smem[slot<R>(a, 0)], in + (u64)blk * Geo::N + a * Geo::T, Geo::T * 8);   // the 32 bulk copies
        for (u32 t = 0; t < Geo::T; t++) pass1<R>(&regs[t * 32], smem.data(), ht.fwd.data(), rp, t, s0, blk);
        for (u32 t = 0; t < Geo::T; t++) pass2<R>(&regs[t * 32], smem.data(), ht.fwd.data(), rp, t, s0, blk);
        for (u32 t = 0; t < Geo::T; t++) pass3_load<R>(&regs[t * 32], smem.data(), t);
        for (u32 t = 0; t < Geo::T; t++) {
            if (s0 == 0) pass3_compute_store<R, true>(&regs[t * 32], out, fwdc.data(), rp, t, s0, blk);
            else pass3_compute_store<R, false>(&regs[t * 32], out, fwdc.data(), rp, t, s0, blk);
        }
    }
    return (long long)g_emu_overflow3;
}
template <int R>
static long long emu3_inv(u64 q, u64 psi, const u64* in, u64* out) {
    using namespace v3;
    typedef NttGeo<R> Geo;
    if (!prime_ok(q)) return -1;
    redent_t tab[16];
    u64 tab8[16];
    fill_redtab(tab, q);
    fill_redtab8(tab8, q);
    Red3 rp = make_red3(q, floor_log2(q), tab);
    rp.tab8 = tab8;
    HostTables ht;
    build_tables(Geo::N, q, psi, ht);
    std::vector<tw_t> invc(Geo::N);
    permute_pass3(ht.inv.data(), invc.data(), 10 + R);
    std::vector<u64> smem(Lay<R>::ROW_WORDS), regs((size_t)Geo::T * 32);
    g_emu_overflow3 = 0;
    memcpy(smem.data(), in, Geo::N * 8);   // the flat bulk copy
    for (u32 t = 0; t < Geo::T; t++) inv_pass3_load<R>(&regs[t * 32], smem.data(), t);
    for (u32 t = 0; t < Geo::T; t++) inv_pass3_compute_store<R>(&regs[t * 32], smem.data(), invc.data(), rp, t);
    for (u32 t = 0; t < Geo::T; t++) inv_pass2<R>(&regs[t * 32], smem.data(), ht.inv.data(), rp, t);
    for (u32 t = 0; t < Geo::T; t++) inv_pass1_load<R>(&regs[t * 32], smem.data(), t);
    for (u32 t = 0; t < Geo::T; t++) inv_pass1_compute_store<R>(&regs[t * 32], out, ht.inv.data(), rp, t, ht.ninv, ht.ninv_w1);
    return (long long)g_emu_overflow3;
}
extern "C" long long emu_ntt3_fwd(int R, u64 q, u64 psi, u32 s0, const u64* in, u64* out) {
    if (R == 4) return emu3_fwd<4>(q, psi, s0, in, out);
    if (s0 != 0) return -2;
    if (R == 3) return emu3_fwd<3>(q, psi, 0, in, out);
    if (R == 2) return emu3_fwd<2>(q, psi, 0, in, out);
    return -2;
}
extern "C" long long emu_ntt3_inv(int R, u64 q, u64 psi, const u64* in, u64* out) {
    if (R == 4) return emu3_inv<4>(q, psi, in, out);
    if (R == 3) return emu3_inv<3>(q, psi, in, out);
    if (R == 2) return emu3_inv<2>(q, psi, in, out);
    return -2;
}

// bank census of the skewed layout: 64-bit accesses per half-warp (passes 1, 2), 128-bit per quarter-warp (pass 3)
template <int R>
static int bank3() {
    using namespace v3;
    typedef NttGeo<R> Geo;
    int worst = 1;
    for (u32 wbase = 0; wbase < Geo::T; wbase += 32) {
        for (u32 r = 0; r < 32; r++)
            for (int h = 0; h < 2; h++) {
                int c1[16] = {0}, c2[16] = {0};
                for (int l = 0; l < 16; l++) {
                    const u32 t = wbase + h * 16 + l;
                    c1[slot<R>(r, t) % 16]++;
                    c2[slot<R>(t >> R, r * Geo::RS + (t & (Geo::RS - 1))) % 16]++;
                }
                for (int i = 0; i < 16; i++) { worst = c1[i] > worst ? c1[i] : worst; worst = c2[i] > worst ? c2[i] : worst; }
            }
        const u32 w = wbase >> 5;
        for (u32 g = 0; g < Geo::G; g++)
            for (u32 c = 0; c < Geo::RS; c += 2)
                for (int qw = 0; qw < 4; qw++) {
                    int cnt[8] = {0};
                    for (int l = 0; l < 8; l++) {
                        const u32 s = slot<R>(brev_bits(qw * 8 + l, 5), brev_bits(Geo::G * w + g, 5) * Geo::RS + c);
                        if (s % 2 || s + 1 >= Lay<R>::ROW_WORDS) return 99;   // 128-bit alignment / buffer bound
                        cnt[(s / 2) % 8]++;
                    }
                    for (int i = 0; i < 8; i++) worst = cnt[i] > worst ? cnt[i] : worst;
                }
    }
    // every position maps to its own slot inside the buffer
    std::vector<char> seen(Lay<R>::ROW_WORDS, 0);
    for (u32 a = 0; a < 32; a++)
        for (u32 i = 0; i < Geo::T; i++) {
            const u32 s = slot<R>(a, i);
            if (s >= Lay<R>::ROW_WORDS || seen[s]) return 98;
            seen[s] = 1;
        }
    return worst;
}
extern "C" int emu_bank_conflicts3(int R) { return R == 4 ? bank3<4>() : R == 3 ? bank3<3>() : R == 2 ? bank3<2>() : -1; }
