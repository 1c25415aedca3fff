// Third-generation persistent row kernels (ntt_v3_kernels.cuh) for N = 2^12 and 2^13 -- the ring degrees of the
// reference's BFV tests (test/bfv_crt.jl: 2048..4096) and of examples/encrypted_mnist (N = 2^13, 60/40-bit chain).
// Two (N = 2^13) or four (N = 2^12) CTAs are resident per SM.  Same arithmetic and index maps as the N = 2^14
// instantiation in ntt_kernels3.cu; this translation unit exists so the instantiations compile in parallel.
#include "ntt_v3_kernels.cuh"

int ntt4_setup_device() {
    int rc = v3k::setup_s<2>();
    if (rc) return rc;
    return v3k::setup_s<3>();
}

// Returns -1 when the third-generation kernels do not apply (the caller falls back to ntt_kernels.cu).
int launch_ntt_s(tfb_ctx* c, const u64* in, u64* out, u64 rows, bool inverse, cudaStream_t st) {
    if (!c->v3_ok || g_ntt_force_harvey || g_ntt_max_mode < 2) return -1;
    if (c->logN == 12) return v3k::launch_s<2>(c, in, out, rows, inverse, st);
    if (c->logN == 13) return v3k::launch_s<3>(c, in, out, rows, inverse, st);
    return -1;
}

int launch_ntt_s_bcast(tfb_ctx* c, const u64* in, u64* out, u64 polys, cudaStream_t st) {
    if (c->logN == 12) return v3k::launch_s<2>(c, in, out, polys * c->L, false, st, c->L);
    if (c->logN == 13) return v3k::launch_s<3>(c, in, out, polys * c->L, false, st, c->L);
    return -1;
}

int launch_ntt_s_gather(tfb_ctx* c, const void* src, u64* out, u64 rows, cudaStream_t st) {
    const v3k::NttSrc* s = (const v3k::NttSrc*)src;
    if (c->logN == 12) return v3k::launch_s<2>(c, nullptr, out, rows, false, st, 1, s);
    if (c->logN == 13) return v3k::launch_s<3>(c, nullptr, out, rows, false, st, 1, s);
    return -1;
}

int launch_ntt_s_crt(tfb_ctx* c, tfb_ctx* r, const u64* cend, u64 ct_stride, u64* dig, u32 k0, u32 dn, u64 batch, cudaStream_t st) {
    if (r->logN == 12) return v3k::launch_crt<2>(c, r, cend, ct_stride, dig, k0, dn, batch, st);
    if (r->logN == 13) return v3k::launch_crt<3>(c, r, cend, ct_stride, dig, k0, dn, batch, st);
    return -1;
}

int launch_ntt_s_pow2(tfb_ctx* r, const u64* limbs, u32 nl, u32 w, u64* dig, u32 k0, u32 dn, u64 batch, cudaStream_t st) {
    if (r->logN == 12) return v3k::launch_pow2<2>(r, limbs, nl, w, dig, k0, dn, batch, st);
    if (r->logN == 13) return v3k::launch_pow2<3>(r, limbs, nl, w, dig, k0, dn, batch, st);
    return -1;
}
