// TEST-ONLY: runs the per-thread phase bodies of the CUDA NTT kernels
// (toyfhe.jl_b200/csrc/ntt_core.cuh) thread-by-thread on the CPU, so the index
// logic (thread mapping, swizzle, twiddle indices, natural-order store) can be
// checked against the oracle in a container without a GPU.  Never shipped.
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../toyfhe.jl_b200/csrc/ntt_core.cuh"
#include "../../toyfhe.jl_b200/csrc/ntt_core2.cuh"
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

// ---- second-generation kernel (1024 threads x 16 residues, N = 2^14 sub-blocks)
// Execution order mirrors the CUDA kernel's synchronisation: middle passes run warp
// by warp (all 32 lanes load, then all 32 lanes store: only __syncwarp between),
// __syncthreads between passes.  A cross-warp hazard would show up as a mismatch.
template <int MODE>
static int run2(int inverse, u64 q, u64 psi, u32 s0, const u64* in, u64* out);
extern "C" int emu_ntt2(int mode, int inverse, u64 q, u64 psi, u32 s0, const u64* in, u64* out) {
    return mode ? run2<1>(inverse, q, psi, s0, in, out) : run2<0>(inverse, q, psi, s0, in, out);
}
template <int MODE>
static int run2(int inverse, u64 q, u64 psi, u32 s0, const u64* in, u64* out) {
    using namespace v2;
    const RedParams rp = make_red(q, log2floor(q));
    const u64 Nrow = (u64)N << s0;
    HostTables ht;
    build_tables(Nrow, q, psi, ht);
    std::vector<u64> smem(N), regs((size_t)T * 16);
    for (u32 blk = 0; blk < (1u << s0); blk++) {
        if (!inverse) {
            memcpy(smem.data(), in + (u64)blk * N, N * sizeof(u64));  // what the TMA bulk copy delivers
            for (int k = 0; k < 3; k++)
                for (u32 w = 0; w < T / 32; w++) {
                    for (u32 l = 0; l < 32; l++) { u32 t = w * 32 + l; PassCfg c = make_cfg(k, t, s0, blk); fwd_mid_load(&regs[t * 16], smem.data(), c); }
                    for (u32 l = 0; l < 32; l++) { u32 t = w * 32 + l; PassCfg c = make_cfg(k, t, s0, blk); fwd_mid_compute<MODE>(&regs[t * 16], ht.fwd.data(), c, rp); fwd_mid_store(&regs[t * 16], smem.data(), c); }
                }
            for (u32 t = 0; t < T; t++) fwd_last_load(&regs[t * 16], smem.data(), t);
            for (u32 t = 0; t < T; t++) fwd_last_compute_store<MODE>(&regs[t * 16], out, ht.fwd.data(), rp, t, s0, blk);
        } else {
            // s0 == 0: the row is first copied flat into shared memory; s0 > 0: gathered from global
            const u64* src = in;
            std::vector<u64> flat;
            if (s0 == 0) { flat.assign(in, in + N); src = flat.data(); }
            for (u32 t = 0; t < T; t++) inv_first_load(&regs[t * 16], src, t, s0, blk);
            for (u32 t = 0; t < T; t++) inv_first_compute_store(&regs[t * 16], smem.data(), ht.inv.data(), q, t, s0, blk);
            for (int k = 2; k >= 1; k--)
                for (u32 w = 0; w < T / 32; w++) {
                    for (u32 l = 0; l < 32; l++) { u32 t = w * 32 + l; PassCfg c = make_cfg(k, t, s0, blk); inv_mid_load(&regs[t * 16], smem.data(), c); }
                    for (u32 l = 0; l < 32; l++) { u32 t = w * 32 + l; PassCfg c = make_cfg(k, t, s0, blk); inv_mid_compute(&regs[t * 16], ht.inv.data(), c, q); inv_mid_store(&regs[t * 16], smem.data(), c); }
                }
            for (u32 t = 0; t < T; t++) { PassCfg c = make_cfg(0, t, s0, blk); inv_mid_load(&regs[t * 16], smem.data(), c); }
            for (u32 t = 0; t < T; t++) { PassCfg c = make_cfg(0, t, s0, blk); inv_final_compute_store(&regs[t * 16], out + (u64)blk * N, ht.inv.data(), c, q, t, s0, ht.ninv, ht.ninv_w1); }
        }
    }
    return 0;
}

// worst 8-byte-bank conflict degree over all shared-memory access patterns of v2
extern "C" int emu_bank_conflicts2() {
    using namespace v2;
    int worst = 1;
    auto census = [&](u32* addr) {
        for (int h = 0; h < 2; h++) {
            int cnt[16] = {0};
            for (int l = 0; l < 16; l++) cnt[addr[h * 16 + l] % 16]++;
            for (int i = 0; i < 16; i++) worst = cnt[i] > worst ? cnt[i] : worst;
        }
    };
    u32 addr[32];
    for (u32 w = 0; w < T / 32; w++) {
        for (int k = 0; k < 3; k++)
            for (int r = 0; r < 16; r++) {
                for (u32 l = 0; l < 32; l++) { PassCfg c = make_cfg(k, w * 32 + l, 0, 0); addr[l] = (u32)(r >> 2) * c.S4 + c.wl[r & 3]; }
                census(addr);
                for (u32 l = 0; l < 32; l++) { PassCfg c = make_cfg(k, w * 32 + l, 0, 0); addr[l] = (u32)(r >> 2) * c.S4 + c.ws[r & 3]; }
                census(addr);
            }
        for (int g = 0; g < 4; g++)
            for (int e = 0; e < 4; e++) {
                for (u32 l = 0; l < 32; l++) addr[l] = last_slot(last_rest(w * 32 + l, g), e);
                census(addr);
                for (u32 l = 0; l < 32; l++) addr[l] = (brev_bits((u32)e, 2) << 12) | ((u32)g * T + w * 32 + l);  // inverse: flat natural read
                census(addr);
            }
    }
    return worst;
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

// rows of 2^15 as a pair of sub-blocks (cluster of two CTAs exchanging through distributed shared memory)
extern "C" long long emu_ntt3_pair_fwd(u64 q, u64 psi, const u64* in, u64* out) {
    using namespace v3;
    constexpr int R = 4;
    typedef NttGeo<R> Geo;
    if (!prime_ok(q)) return -1;
    redent_t tab[16];
    fill_redtab(tab, q);
    const Red3 rp = make_red3(q, floor_log2(q), tab);
    const u64 Nrow = (u64)Geo::N * 2;
    HostTables ht;
    build_tables(Nrow, q, psi, ht);
    std::vector<tw_t> fwdc(Nrow);
    permute_pass3(ht.fwd.data(), fwdc.data(), 15);
    std::vector<u64> smem[2], regs[2];
    g_emu_overflow3 = 0;
    for (u32 r = 0; r < 2; r++) {
        smem[r].assign(Lay<R>::ROW_WORDS, 0);
        regs[r].assign((size_t)Geo::T * 32, 0);
        for (u32 a = 0; a < 32; a++) memcpy(&smem[r][slot<R>(a, 0)], in + (u64)r * Geo::N + a * Geo::T, Geo::T * 8);
    }
    for (u32 r = 0; r < 2; r++)
        for (u32 t = 0; t < Geo::T; t++) {
            pass1_cross_load<R>(&regs[r][t * 32], smem[r].data(), smem[1 - r].data(), r, ht.fwd[1], rp, t);
            pass1_cross_levels(&regs[r][t * 32], ht.fwd.data(), rp, 1, r);
        }
    for (u32 r = 0; r < 2; r++) {
        for (u32 t = 0; t < Geo::T; t++) pass1_store<R>(&regs[r][t * 32], smem[r].data(), t);
        for (u32 t = 0; t < Geo::T; t++) pass2<R>(&regs[r][t * 32], smem[r].data(), ht.fwd.data(), rp, t, 1, r);
        for (u32 t = 0; t < Geo::T; t++) pass3_load<R>(&regs[r][t * 32], smem[r].data(), t);
        for (u32 t = 0; t < Geo::T; t++) pass3_compute_store<R, false>(&regs[r][t * 32], out, fwdc.data(), rp, t, 1, r);
    }
    return (long long)g_emu_overflow3;
}
extern "C" long long emu_ntt3_pair_inv(u64 q, u64 psi, const u64* in, u64* out) {
    using namespace v3;
    constexpr int R = 4;
    typedef NttGeo<R> Geo;
    if (!prime_ok(q)) return -1;
    redent_t tab[16];
    u64 tab8[16];
    fill_redtab(tab, q);
    fill_redtab8(tab8, q);
    Red3 rp = make_red3(q, floor_log2(q), tab);
    rp.tab8 = tab8;
    const u64 Nrow = (u64)Geo::N * 2;
    HostTables ht;
    build_tables(Nrow, q, psi, ht);
    std::vector<tw_t> invc(Nrow);
    permute_pass3(ht.inv.data(), invc.data(), 15);
    std::vector<u64> smem[2], regs[2];
    g_emu_overflow3 = 0;
    for (u32 r = 0; r < 2; r++) {
        smem[r].assign(Lay<R>::ROW_WORDS, 0);
        regs[r].assign((size_t)Geo::T * 32, 0);
        memcpy(smem[r].data(), in + (u64)r * Geo::N, Geo::N * 8);   // one contiguous half each
    }
    for (u32 r = 0; r < 2; r++)
        for (u32 t = 0; t < Geo::T; t++) inv_pass3_load_pair<R>(&regs[r][t * 32], smem[r].data(), smem[1 - r].data(), r, t);
    for (u32 r = 0; r < 2; r++) {
        for (u32 t = 0; t < Geo::T; t++) inv_pass3_compute_store<R>(&regs[r][t * 32], smem[r].data(), invc.data(), rp, t, r);
        for (u32 t = 0; t < Geo::T; t++) inv_pass2<R>(&regs[r][t * 32], smem[r].data(), ht.inv.data(), rp, t, 1, r);
        for (u32 t = 0; t < Geo::T; t++) inv_pass1_load<R>(&regs[r][t * 32], smem[r].data(), t);
        for (u32 t = 0; t < Geo::T; t++) inv_pass1_levels_all(&regs[r][t * 32], ht.inv.data(), rp, 1, r);
        for (u32 t = 0; t < Geo::T; t++) pass1_store<R>(&regs[r][t * 32], smem[r].data(), t);
    }
    for (u32 r = 0; r < 2; r++)
        for (u32 t = 0; t < Geo::T; t++) {
            inv_cross_combine<R>(&regs[r][t * 32], smem[1 - r].data(), r, rp, t);
            inv_cross_finish<R>(&regs[r][t * 32], out + (u64)r * Geo::N, r ? ht.ninv_w1 : ht.ninv, rp, t);
        }
    return (long long)g_emu_overflow3;
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
