// TEST-ONLY: runs the per-thread phase bodies of the CUDA NTT kernels
// (toyfhe.jl_b200/csrc/ntt_core.cuh) thread-by-thread on the CPU, so the index
// logic (thread mapping, swizzle, twiddle indices, natural-order store) can be
// checked against the oracle in a container without a GPU.  Never shipped.
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../toyfhe.jl_b200/csrc/ntt_core.cuh"
#include "../../toyfhe.jl_b200/csrc/tables.h"

template <int R>
static void run(int inverse, u64 q, u64 psi, u32 s0, const u64* in, u64* out) {
    typedef NttGeo<R> Geo;
    const u64 Nrow = (u64)Geo::N << s0;
    HostTables ht;
    build_tables(Nrow, q, psi, ht);
    std::vector<u64> smem(Geo::N), regs((size_t)Geo::T * 32);
    for (u32 blk = 0; blk < (1u << s0); blk++) {
        if (!inverse) {
            for (u32 t = 0; t < Geo::T; t++) fwd_phaseA<R>(&regs[t * 32], in + (u64)blk * Geo::N, smem.data(), ht.fwd.data(), q, t, s0, blk);
            for (u32 t = 0; t < Geo::T; t++) fwd_phaseB<R>(&regs[t * 32], smem.data(), ht.fwd.data(), q, t, s0, blk);
            for (u32 t = 0; t < Geo::T; t++) fwd_phaseC<R>(&regs[t * 32], out, smem.data(), ht.fwd.data(), q, t, s0, blk);
        } else {
            for (u32 t = 0; t < Geo::T; t++) inv_phaseC<R>(&regs[t * 32], in, smem.data(), ht.inv.data(), q, t, s0, blk);
            for (u32 t = 0; t < Geo::T; t++) inv_phaseB<R>(&regs[t * 32], smem.data(), ht.inv.data(), q, t, s0, blk);
            for (u32 t = 0; t < Geo::T; t++) inv_phaseA<R>(&regs[t * 32], out + (u64)blk * Geo::N, smem.data(), ht.inv.data(), q, t, s0, blk, ht.ninv, ht.ninv_w1);
        }
    }
}

// emulates the fast kernel on one row of length 2^(10+R+s0).  For s0>0 only the
// row-resident part is emulated: forward expects stages 1..s0 already applied to
// `in`; inverse leaves stages s0..1 (and the N^-1 scale) to the caller.
extern "C" int emu_ntt(int R, int inverse, u64 q, u64 psi, u32 s0, const u64* in, u64* out) {
    switch (R) {
        case 0: run<0>(inverse, q, psi, s0, in, out); break;
        case 1: run<1>(inverse, q, psi, s0, in, out); break;
        case 2: run<2>(inverse, q, psi, s0, in, out); break;
        case 3: run<3>(inverse, q, psi, s0, in, out); break;
        case 4: run<4>(inverse, q, psi, s0, in, out); break;
        default: return 1;
    }
    return 0;
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
