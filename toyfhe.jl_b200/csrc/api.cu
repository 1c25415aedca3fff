// C-ABI of libtoyfhe_b200.so (see include/toyfhe_b200.h for the contract and the
// reference methods each entry point replaces).
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <new>

#include "engine.h"
#include "ntt_core3.cuh"
#include "tables.h"

static thread_local std::string g_err;
void tfb_set_error(const std::string& msg) { g_err = msg; }
int tfb_cuda_fail(cudaError_t e, const char* what) {
    g_err = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
    return TFB_ECUDA;
}

// Every entry point that takes a context runs with that context's device current and restores the caller's device on
// return (one process may drive several GPUs; contexts are destroyed from garbage collectors at arbitrary points).
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
// Calls that use a context's scratch (ws / stage / io, cached joint contexts) on DIFFERENT streams are ordered on the device:
// when a call arrives on another stream than the previous one, an event is recorded on the PREVIOUS stream (it covers
// everything submitted there so far, in particular the previous call's last kernel) and the new stream waits on it.
// Calls that stay on one stream -- the normal case -- cost nothing: stream order already serialises them.
struct ScratchGuard {
    tfb_ctx* c;
    cudaStream_t st;
    ScratchGuard(tfb_ctx* c_, void* stream) : c(c_), st((cudaStream_t)stream) {
        if (c->scratch_used && c->scratch_stream != st) {
            if (!c->scratch_done) cudaEventCreateWithFlags(&c->scratch_done, cudaEventDisableTiming);
            if (c->scratch_done && cudaEventRecord(c->scratch_done, c->scratch_stream) == cudaSuccess) cudaStreamWaitEvent(st, c->scratch_done, 0);
            else cudaGetLastError();   // the previous stream no longer exists: its work has completed or been abandoned by the caller
        }
    }
    ~ScratchGuard() {
        c->scratch_stream = st;
        c->scratch_used = true;
    }
    ScratchGuard(const ScratchGuard&) = delete;
    ScratchGuard& operator=(const ScratchGuard&) = delete;
};
#define TFB_CAT2(a, b) a##b
#define TFB_CAT(a, b) TFB_CAT2(a, b)
#define CHECK_CTX(c)                                                \
    if (!(c)) { tfb_set_error("null context"); return TFB_EINVAL; } \
    DeviceGuard TFB_CAT(dg_, __COUNTER__)((c)->device)
#define CHECK_ROWS(c, rows)                                                                         \
    do {                                                                                            \
        if ((rows) % (c)->L) { tfb_set_error("rows must be a multiple of the number of primes"); return TFB_EINVAL; } \
    } while (0)
#define CHECK_PTR(p)                                                       \
    do {                                                                   \
        if (!(p)) { tfb_set_error("null buffer: " #p); return TFB_EINVAL; } \
    } while (0)

static int grow(void** p, size_t* have, size_t bytes) {
    if (*have >= bytes) return TFB_OK;
    if (*p) {
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) return tfb_cuda_fail(e, "cudaDeviceSynchronize");
        cudaFree(*p);
        *p = nullptr;
        *have = 0;
    }
    size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(p, want);
    if (e != cudaSuccess) {
        e = cudaMalloc(p, bytes);
        want = bytes;
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        tfb_set_error("out of device memory for engine scratch");
        return TFB_ENOMEM;
    }
    *have = want;
    return TFB_OK;
}
int ws_reserve(tfb_ctx* c, size_t bytes) { return grow(&c->ws, &c->ws_bytes, bytes); }
int stage_reserve(tfb_ctx* c, size_t bytes) { return grow(&c->stage, &c->stage_bytes, bytes); }

// ------------------------------------------------------------------ profiling
#include <map>
#include <mutex>
static bool g_prof_on = false;
static std::mutex g_prof_mu;
struct ProfRec { int cls; cudaEvent_t a, b; };
static std::vector<ProfRec> g_prof_recs;
static std::vector<cudaEvent_t> g_prof_pool;
static unsigned long long g_prof_count[PC_COUNT];
static double g_prof_ms[PC_COUNT];
static cudaEvent_t prof_event() {
    if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
ProfScope::ProfScope(int cls_, cudaStream_t st_) : cls(cls_), st(st_), stop(nullptr) {
    tfb_count_launch();
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    cudaEvent_t a = prof_event();
    stop = prof_event();
    cudaEventRecord(a, st);
    g_prof_recs.push_back({cls, a, stop});
}
ProfScope::~ProfScope() {
    if (stop) cudaEventRecord(stop, st);
}
static const char* kProfNames[PC_COUNT] = {"ntt_fwd_row", "ntt_inv_row", "ntt_other", "elementwise", "tensor_dual",
                                           "base_switch", "bfv_contract", "ks_digits", "ks_accum", "ks_finish", "level"};

extern "C" {

int tfb_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_on = on != 0;
    return TFB_OK;
}
// drains the recorded launches (synchronises their events) into per-class totals;
// counts/ms are arrays of tfb_profile_classes() entries and are ADDED to.
int tfb_profile_read(unsigned long long* counts, double* ms, int reset) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& r : g_prof_recs) {
        cudaError_t e = cudaEventSynchronize(r.b);
        if (e != cudaSuccess) return tfb_cuda_fail(e, "cudaEventSynchronize");
        float t = 0;
        cudaEventElapsedTime(&t, r.a, r.b);
        g_prof_count[r.cls]++;
        g_prof_ms[r.cls] += t;
        g_prof_pool.push_back(r.a);
        g_prof_pool.push_back(r.b);
    }
    g_prof_recs.clear();
    for (int i = 0; i < PC_COUNT; i++) {
        if (counts) counts[i] = g_prof_count[i];
        if (ms) ms[i] = g_prof_ms[i];
        if (reset) { g_prof_count[i] = 0; g_prof_ms[i] = 0; }
    }
    return TFB_OK;
}
int tfb_profile_classes(void) { return PC_COUNT; }
int tfb_debug_ntt_version(int v) {
    g_ntt_version = (v == 1 || v == 3) ? v : 3;
    return TFB_OK;
}
int tfb_debug_ntt_max_mode(int m) {
    g_ntt_max_mode = m < 0 ? 0 : (m > 2 ? 2 : m);
    return TFB_OK;
}
int tfb_debug_ntt_cross(int on) {
    extern bool g_ntt_cross;
    g_ntt_cross = on != 0;
    return TFB_OK;
}
int tfb_debug_ntt_force_harvey(int on) {
    extern bool g_ntt_force_harvey;
    g_ntt_force_harvey = on != 0;
    return TFB_OK;
}
int tfb_debug_force_generic(int on) {
    extern bool g_force_generic, g_force_generic_red;
    g_force_generic = on == 1;        // 1: generic runtime-L conversion kernels
    g_force_generic_red = on == 2;    // 2: specialised kernels, but Shoup/Barrett reductions even on 2^60 + e primes
    return TFB_OK;
}
const char* tfb_profile_class_name(int i) { return (i >= 0 && i < PC_COUNT) ? kProfNames[i] : ""; }

const char* tfb_last_error(void) { return g_err.c_str(); }
int tfb_version(void) { return 100; }
unsigned long long tfb_kernel_launches(void) { return tfb_launch_count(); }

int tfb_minimal_primitive_root(uint64_t q, uint64_t n, uint64_t* out) {
    if (!out || !h_is_prime(q) || !h_minimal_primitive_root(q, n, out)) {
        tfb_set_error("no primitive n-th root: need prime q with n | q-1, n a power of two");
        return TFB_EINVAL;
    }
    return TFB_OK;
}

int tfb_prime_chain(uint32_t N, const int32_t* logqs, uint32_t n, uint64_t* q_out, uint64_t* psi_out) {
    if (!logqs || !q_out || !N || (N & (N - 1))) { tfb_set_error("prime chain: bad arguments"); return TFB_EINVAL; }
    // ascending-logq generation order (stable), results in the caller's order (crt.jl:283-291)
    std::vector<uint32_t> perm(n);
    for (uint32_t i = 0; i < n; i++) perm[i] = i;
    for (uint32_t i = 1; i < n; i++)
        for (uint32_t j = i; j > 0 && logqs[perm[j - 1]] > logqs[perm[j]]; j--) std::swap(perm[j - 1], perm[j]);
    u64 last = 0;
    const u64 step = 2ull * N;
    for (uint32_t k = 0; k < n; k++) {
        const int lq = logqs[perm[k]];
        if (lq < 2 || lq > 61) { tfb_set_error("prime chain: logq must be in 2..61"); return TFB_EUNSUPPORTED; }
        u64 p = (1ull << lq) + 1;
        if (last + step > p) p = last + step;
        while (!h_is_prime(p)) {
            p += step;
            if (p >> 62) { tfb_set_error("prime chain: ran past 2^62"); return TFB_EUNSUPPORTED; }
        }
        last = p;
        q_out[perm[k]] = p;
    }
    if (psi_out)
        for (uint32_t i = 0; i < n; i++)
            if (!h_minimal_primitive_root(q_out[i], step, &psi_out[i])) { tfb_set_error("prime chain: no 2N-th root"); return TFB_EINVAL; }
    return TFB_OK;
}

int tfb_ndigits(const uint64_t* q, uint32_t L, uint32_t w, uint32_t* out) {
    if (!q || !out || !L || !w) { tfb_set_error("ndigits: bad arguments"); return TFB_EINVAL; }
    // bit length of Q = prod q_i with a little-endian limb product
    std::vector<u64> X(1, 1);
    for (uint32_t i = 0; i < L; i++) {
        u64 carry = 0;
        for (size_t k = 0; k < X.size(); k++) {
            u128 t = (u128)X[k] * q[i] + carry;
            X[k] = (u64)t;
            carry = (u64)(t >> 64);
        }
        if (carry) X.push_back(carry);
    }
    size_t bits = (X.size() - 1) * 64 + (64 - __builtin_clzll(X.back()));
    *out = (uint32_t)((bits + w - 1) / w);
    return TFB_OK;
}

static inline u32 logN_of(u64 N) { u32 l = 0; while ((1ull << l) < N) l++; return l; }
int tfb_ctx_create(int device, uint32_t N, uint32_t L, const uint64_t* q, const uint64_t* psi, tfb_ctx** out) {
    if (!out || !q || !psi) { tfb_set_error("ctx_create: null argument"); return TFB_EINVAL; }
    *out = nullptr;
    if (N < 2 || (N & (N - 1))) { tfb_set_error("ctx_create: N must be a power of two >= 2"); return TFB_EINVAL; }
    if (N > (1u << 16)) { tfb_set_error("ctx_create: N > 2^16 is not supported"); return TFB_EUNSUPPORTED; }
    if (L < 1 || L > TFB_MAX_L) { tfb_set_error("ctx_create: L must be in 1..64"); return TFB_EINVAL; }
    for (uint32_t i = 0; i < L; i++) {
        if (q[i] >> 62) { tfb_set_error("ctx_create: modulus must be < 2^62"); return TFB_EUNSUPPORTED; }
        if (q[i] < 3 || (q[i] - 1) % (2ull * N)) { tfb_set_error("ctx_create: need q = 1 (mod 2N)"); return TFB_EINVAL; }
        if (!h_is_prime(q[i])) { tfb_set_error("ctx_create: modulus is not prime"); return TFB_EINVAL; }
        if (psi[i] >= q[i]) { tfb_set_error("ctx_create: psi must be < q"); return TFB_EINVAL; }
        // is_primitive_root(psi, 2N) (pow2_cyc_rings.jl:22,31) -- we also require psi^N = -1
        if (h_powmod(psi[i], N, q[i]) != q[i] - 1) { tfb_set_error("ctx_create: psi is not a primitive 2N-th root of unity"); return TFB_EINVAL; }
        for (uint32_t j = 0; j < i; j++)
            if (q[j] == q[i]) { tfb_set_error("ctx_create: repeated modulus"); return TFB_EINVAL; }
    }
    {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); tfb_set_error("ctx_create: no such CUDA device"); return TFB_EINVAL; }
    }
    DeviceGuard dg(device);
    tfb_ctx* c = new tfb_ctx();
    static std::atomic<u64> next_uid{1};
    c->uid = next_uid.fetch_add(1);
    c->device = device;
    c->N = N;
    c->L = L;
    c->logN = 0;
    while ((1u << c->logN) < N) c->logN++;
    c->q.assign(q, q + L);
    c->psi.assign(psi, psi + L);
    c->d_fwd = c->d_inv = nullptr;
    c->d_pp = nullptr;
    c->d_ginv = nullptr;
    c->d_halfmr = nullptr;
    c->ws = c->stage = c->io = nullptr;
    c->d_ckks_pos = nullptr;
    c->scratch_done = nullptr;
    c->scratch_stream = nullptr;
    c->scratch_used = false;
    c->ws_bytes = c->stage_bytes = c->io_bytes = 0;
    c->conv_ok = false;
    c->num_sms = 0;
    c->ntt_mode = 1;
    c->v3_ok = true;
    bool all60 = true;
    // [0, L*N): natural psi^brev(k) tables; [L*N, 2*L*N): thread-order copies for pass 3 (tables.h permute_pass3)
    std::vector<tw_t> fwd((size_t)2 * L * N), inv((size_t)2 * L * N);
    std::vector<PrimeParams> pp(L);
    for (uint32_t i = 0; i < L; i++) {
        HostTables ht;
        build_tables(N, q[i], psi[i], ht);
        memcpy(&fwd[(size_t)i * N], ht.fwd.data(), (size_t)N * sizeof(tw_t));
        memcpy(&inv[(size_t)i * N], ht.inv.data(), (size_t)N * sizeof(tw_t));
        permute_pass3(ht.fwd.data(), &fwd[(size_t)(L + i) * N], (int)logN_of(N));
        permute_pass3(ht.inv.data(), &inv[(size_t)(L + i) * N], (int)logN_of(N));
        pp[i].pc = ht.pc;
        pp[i].ninv = ht.ninv;
        pp[i].ninv_w1 = ht.ninv_w1;
        u32 sh = 63 - (u32)__builtin_clzll(q[i]);
        pp[i].sh = sh;
        pp[i].pad_ = 0;
        const u64 e = q[i] - (1ull << sh);
        const bool lazy_ok = (e <= ((1ull << sh) >> 4)) && ((u128)q[i] * 15 < ((u128)1 << 64));
        if (!lazy_ok) c->ntt_mode = 0;
        if (sh != 60 || e >= (1ull << 28)) all60 = false;
        if (!lazy_ok || !v3::prime_ok(q[i])) c->v3_ok = false;
    }
    if (c->ntt_mode == 1 && all60) c->ntt_mode = 2;   // approximate-quotient ladder (ntt_core.cuh MODE 2)
    int rc = TFB_OK;
    cudaError_t e;
    if ((e = cudaMalloc(&c->d_fwd, fwd.size() * sizeof(tw_t))) != cudaSuccess ||
        (e = cudaMalloc(&c->d_inv, inv.size() * sizeof(tw_t))) != cudaSuccess ||
        (e = cudaMalloc(&c->d_pp, L * sizeof(PrimeParams))) != cudaSuccess ||
        (e = cudaMemcpy(c->d_fwd, fwd.data(), fwd.size() * sizeof(tw_t), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(c->d_inv, inv.data(), inv.size() * sizeof(tw_t), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(c->d_pp, pp.data(), L * sizeof(PrimeParams), cudaMemcpyHostToDevice)) != cudaSuccess)
        rc = tfb_cuda_fail(e, "ctx_create table upload");
    if (!rc) rc = build_garner(c);
    if (!rc) rc = ntt_setup_device();
    if (!rc) rc = ntt3_setup_device();
    if (!rc) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->num_sms = prop.multiProcessorCount;
    }
    if (rc) {
        std::string keep = g_err;
        tfb_ctx_destroy(c);
        g_err = keep;
        return rc;
    }
    *out = c;
    return TFB_OK;
}

static void forget_joint(const tfb_ctx* c);
static void forget_pipe(const tfb_ctx* c);
int tfb_ctx_destroy(tfb_ctx* c) {
    if (!c) return TFB_OK;
    DeviceGuard dg(c->device);
    tfb_forget_ctx_pairs(c);
    forget_joint(c);
    forget_pipe(c);
    cudaFree(c->d_fwd);
    cudaFree(c->d_inv);
    cudaFree(c->d_pp);
    cudaFree(c->d_ginv);
    cudaFree(c->d_halfmr);
    cudaFree(c->ws);
    cudaFree(c->stage);
    cudaFree(c->io);
    cudaFree(c->d_ckks_pos);
    if (c->scratch_done) cudaEventDestroy(c->scratch_done);
    delete c;
    return TFB_OK;
}

int tfb_ctx_info(const tfb_ctx* c, uint32_t* N, uint32_t* L, uint64_t* q, uint64_t* psi) {
    CHECK_CTX(c);
    if (N) *N = c->N;
    if (L) *L = c->L;
    if (q) memcpy(q, c->q.data(), c->L * sizeof(u64));
    if (psi) memcpy(psi, c->psi.data(), c->L * sizeof(u64));
    return TFB_OK;
}

int tfb_malloc(tfb_ctx* c, size_t bytes, void** dptr) {
    CHECK_CTX(c);
    CHECK_PTR(dptr);
    cudaError_t e = cudaMalloc(dptr, bytes ? bytes : 1);
    if (e != cudaSuccess) { cudaGetLastError(); tfb_set_error("out of device memory"); return TFB_ENOMEM; }
    return TFB_OK;
}
int tfb_free(tfb_ctx* c, void* dptr) {
    CHECK_CTX(c);
    TFB_CUDA(cudaFree(dptr));
    return TFB_OK;
}
int tfb_memcpy_h2d(tfb_ctx* c, void* dst, const void* src, size_t bytes, void* stream) {
    CHECK_CTX(c);
    TFB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return TFB_OK;
}
int tfb_memcpy_d2h(tfb_ctx* c, void* dst, const void* src, size_t bytes, void* stream) {
    CHECK_CTX(c);
    TFB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return TFB_OK;
}
int tfb_sync(tfb_ctx* c, void* stream) {
    CHECK_CTX(c);
    TFB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return TFB_OK;
}

// ------------------------------------------------------------------ transforms
int tfb_ntt_fwd(tfb_ctx* c, const uint64_t* in, uint64_t* out, uint64_t rows, void* stream) {
    CHECK_CTX(c); CHECK_ROWS(c, rows);
    ScratchGuard TFB_CAT(sg_, __COUNTER__)(c, stream);
    if (!rows) return TFB_OK;
    CHECK_PTR(in); CHECK_PTR(out);
    return launch_ntt(c, in, out, rows, false, (cudaStream_t)stream);
}
int tfb_ntt_inv(tfb_ctx* c, const uint64_t* in, uint64_t* out, uint64_t rows, void* stream) {
    CHECK_CTX(c); CHECK_ROWS(c, rows);
    ScratchGuard TFB_CAT(sg_, __COUNTER__)(c, stream);
    if (!rows) return TFB_OK;
    CHECK_PTR(in); CHECK_PTR(out);
    return launch_ntt(c, in, out, rows, true, (cudaStream_t)stream);
}

static int binop(tfb_ctx* c, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, uint64_t rows, void* stream) {
    CHECK_CTX(c); CHECK_ROWS(c, rows);
    if (!rows) return TFB_OK;
    CHECK_PTR(a); CHECK_PTR(b); CHECK_PTR(out);
    return launch_binop(c, op, a, b, out, rows, (cudaStream_t)stream);
}
int tfb_add(tfb_ctx* c, const uint64_t* a, const uint64_t* b, uint64_t* out, uint64_t rows, void* s) { return binop(c, 0, a, b, out, rows, s); }
int tfb_sub(tfb_ctx* c, const uint64_t* a, const uint64_t* b, uint64_t* out, uint64_t rows, void* s) { return binop(c, 1, a, b, out, rows, s); }
int tfb_mul(tfb_ctx* c, const uint64_t* a, const uint64_t* b, uint64_t* out, uint64_t rows, void* s) { return binop(c, 2, a, b, out, rows, s); }
int tfb_neg(tfb_ctx* c, const uint64_t* a, uint64_t* out, uint64_t rows, void* stream) {
    CHECK_CTX(c); CHECK_ROWS(c, rows);
    if (!rows) return TFB_OK;
    CHECK_PTR(a); CHECK_PTR(out);
    return launch_neg(c, a, out, rows, (cudaStream_t)stream);
}
int tfb_scalar_mul(tfb_ctx* c, const uint64_t* a, const uint64_t* s_residues, uint64_t* out, uint64_t rows, void* stream) {
    CHECK_CTX(c); CHECK_ROWS(c, rows);
    if (!rows) return TFB_OK;
    CHECK_PTR(a); CHECK_PTR(out); CHECK_PTR(s_residues);
    return launch_scalar_mul(c, a, s_residues, out, rows, (cudaStream_t)stream);
}

int tfb_mul_plain(tfb_ctx* c, const uint64_t* a, const uint64_t* plain, uint64_t* out, uint64_t polys, int accumulate, void* stream) {
    CHECK_CTX(c);
    if (!polys) return TFB_OK;
    CHECK_PTR(a); CHECK_PTR(plain); CHECK_PTR(out);
    if (c->N < 2) { tfb_set_error("mul_plain: N must be at least 2"); return TFB_EINVAL; }
    return launch_mul_plain(c, a, plain, out, polys, accumulate != 0, (cudaStream_t)stream);
}

int tfb_add_plain(tfb_ctx* c, const uint64_t* a, const uint64_t* plain, uint64_t* out, uint64_t polys, uint64_t stride_words, void* stream) {
    CHECK_CTX(c);
    if (!polys) return TFB_OK;
    CHECK_PTR(a); CHECK_PTR(plain); CHECK_PTR(out);
    if (c->N < 2 || stride_words % 2 || stride_words < (uint64_t)c->L * c->N) { tfb_set_error("add_plain: stride must be even and at least one polynomial"); return TFB_EINVAL; }
    return launch_add_plain(c, a, plain, out, polys, stride_words, (cudaStream_t)stream);
}

int tfb_lincomb(tfb_ctx* c, const uint64_t* in, uint64_t in_stride_words, uint32_t J, const uint64_t* weights, uint32_t C, uint64_t* out,
                uint64_t polys, void* stream) {
    CHECK_CTX(c);
    if (!polys) return TFB_OK;
    CHECK_PTR(in); CHECK_PTR(weights); CHECK_PTR(out);
    if (J < 1 || J > 63 || C < 1 || C > 4) { tfb_set_error("lincomb: need 1 <= J <= 63 terms and 1 <= C <= 4 outputs"); return TFB_EINVAL; }
    if (in_stride_words < polys * c->L * c->N) { tfb_set_error("lincomb: input stride shorter than one input"); return TFB_EINVAL; }
    return launch_lincomb(c, in, in_stride_words, J, weights, C, out, polys, (cudaStream_t)stream);
}

int tfb_ring_mul(tfb_ctx* c, const uint64_t* a, const uint64_t* b, uint64_t* out, uint64_t rows, void* stream) {
    CHECK_CTX(c); CHECK_ROWS(c, rows);
    ScratchGuard TFB_CAT(sg_, __COUNTER__)(c, stream);
    if (!rows) return TFB_OK;
    CHECK_PTR(a); CHECK_PTR(b); CHECK_PTR(out);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t words = (size_t)rows * c->N;
    // the long-row NTT path uses c->ws itself, so keep our temporaries in `stage`
    int rc = stage_reserve(c, 2 * words * sizeof(u64));
    if (rc) return rc;
    u64* A = (u64*)c->stage;
    u64* B = A + words;
    if ((rc = launch_ntt(c, a, A, rows, false, st))) return rc;
    if ((rc = launch_ntt(c, b, B, rows, false, st))) return rc;
    if ((rc = launch_binop(c, 2, A, B, A, rows, st))) return rc;
    return launch_ntt(c, A, out, rows, true, st);
}

int tfb_galois(tfb_ctx* c, uint64_t g, const uint64_t* in, uint64_t* out, uint64_t rows, void* stream) {
    CHECK_CTX(c); CHECK_ROWS(c, rows);
    if (!rows) return TFB_OK;
    CHECK_PTR(in); CHECK_PTR(out);
    return launch_galois(c, g, in, out, rows, (cudaStream_t)stream);
}

int tfb_rescale(tfb_ctx* c, const uint64_t* in, uint64_t* out, uint64_t polys, void* stream) {
    CHECK_CTX(c);
    if (!polys) return TFB_OK;
    CHECK_PTR(in); CHECK_PTR(out);
    if (in == out) { tfb_set_error("tfb_rescale cannot run in place"); return TFB_EINVAL; }
    return launch_rescale(c, in, out, polys, (cudaStream_t)stream);
}
int tfb_crt_expand(tfb_ctx* c, uint64_t P, const uint64_t* in, uint64_t* out, uint64_t polys, void* stream) {
    CHECK_CTX(c);
    if (!polys) return TFB_OK;
    CHECK_PTR(in); CHECK_PTR(out);
    if (in == out) { tfb_set_error("tfb_crt_expand cannot run in place"); return TFB_EINVAL; }
    return launch_crt_expand(c, P, in, out, polys, (cudaStream_t)stream);
}

// ------------------------------------------------------------ ciphertext multiply
static int ct_tensor_dev(tfb_ctx* c, const u64* c1, const u64* c2, u64* out, u64 batch, cudaStream_t st) {
    const size_t poly = (size_t)c->L * c->N;
    int rc = stage_reserve(c, 4 * batch * poly * sizeof(u64));
    if (rc) return rc;
    u64* A = (u64*)c->stage;
    u64* B = c1 == c2 ? A : A + 2 * batch * poly;   // squaring (c*c, rlwe_she.jl:264-266 with one operand): transform once
    if (c1 == c2) {
        if ((rc = launch_ntt(c, c1, A, 2 * batch * c->L, false, st))) return rc;
    } else {
        // both operands in ONE forward launch (rows gathered from the two buffers); two launches where that kernel family does not apply
        rc = launch_ntt_gather(c, c1, c2, (u32)(2 * batch), c->L, nullptr, A, 4 * batch, st);
        if (rc == -1) {
            if ((rc = launch_ntt(c, c1, A, 2 * batch * c->L, false, st))) return rc;
            rc = launch_ntt(c, c2, B, 2 * batch * c->L, false, st);
        }
        if (rc) return rc;
    }
    if ((rc = launch_tensor_dual(c, A, B, out, batch, st))) return rc;
    return launch_ntt(c, out, out, 3 * batch * c->L, true, st);
}

int tfb_ct_tensor(tfb_ctx* c, const uint64_t* c1, const uint64_t* c2, uint64_t* out, uint64_t batch, void* stream) {
    CHECK_CTX(c);
    ScratchGuard TFB_CAT(sg_, __COUNTER__)(c, stream);
    if (!batch) return TFB_OK;
    CHECK_PTR(c1); CHECK_PTR(c2); CHECK_PTR(out);
    return ct_tensor_dev(c, c1, c2, out, batch, (cudaStream_t)stream);
}

int tfb_bfv_switch(tfb_ctx* from, tfb_ctx* to, const uint64_t* in, uint64_t* out, uint64_t polys, void* stream) {
    CHECK_CTX(from); CHECK_CTX(to);
    if (!polys) return TFB_OK;
    CHECK_PTR(in); CHECK_PTR(out);
    if (in == out) { tfb_set_error("tfb_bfv_switch cannot run in place"); return TFB_EINVAL; }
    return launch_base_switch(from, to, in, out, polys, (cudaStream_t)stream);
}
int tfb_bfv_contract(tfb_ctx* cq, tfb_ctx* cb, uint64_t t, const uint64_t* in, uint64_t* out, uint64_t polys, void* stream) {
    CHECK_CTX(cq); CHECK_CTX(cb);
    if (!polys) return TFB_OK;
    if (t == 0) { tfb_set_error("bfv_contract: plaintext modulus is zero"); return TFB_EINVAL; }
    CHECK_PTR(in); CHECK_PTR(out);
    if (in == out) { tfb_set_error("tfb_bfv_contract cannot run in place"); return TFB_EINVAL; }
    return launch_bfv_contract(cq, cb, t, in, out, polys, (cudaStream_t)stream);
}

int tfb_bfv_encode(tfb_ctx* c, uint64_t t, const uint64_t* delta, uint32_t nl, const uint64_t* m, uint64_t* out, uint64_t polys, void* stream) {
    CHECK_CTX(c);
    if (!polys) return TFB_OK;
    CHECK_PTR(delta); CHECK_PTR(m); CHECK_PTR(out);
    return launch_bfv_encode(c, t, delta, nl, m, out, polys, (cudaStream_t)stream);
}
int tfb_bfv_decode(tfb_ctx* c, uint64_t t, const uint64_t* delta, uint32_t nl, const uint64_t* b, uint64_t* out, uint64_t polys, void* stream) {
    CHECK_CTX(c);
    if (!polys) return TFB_OK;
    CHECK_PTR(delta); CHECK_PTR(b); CHECK_PTR(out);
    return launch_bfv_decode(c, t, delta, nl, b, out, polys, (cudaStream_t)stream);
}

int tfb_centered_mod(tfb_ctx* c, uint64_t t, const uint64_t* b, uint64_t* out, uint64_t polys, void* stream) {
    CHECK_CTX(c);
    if (!polys) return TFB_OK;
    CHECK_PTR(b); CHECK_PTR(out);
    return launch_centered_mod(c, t, b, out, polys, (cudaStream_t)stream);
}
int tfb_ckks_encode(tfb_ctx* c, double scale, const double* slots, uint64_t* out, uint64_t polys, void* stream) {
    CHECK_CTX(c);
    ScratchGuard TFB_CAT(sg_, __COUNTER__)(c, stream);
    if (!polys) return TFB_OK;
    CHECK_PTR(slots); CHECK_PTR(out);
    return launch_ckks_encode(c, scale, slots, out, polys, (cudaStream_t)stream);
}
int tfb_ckks_decode(tfb_ctx* c, double scale, const uint64_t* in, double* slots, uint64_t polys, void* stream) {
    CHECK_CTX(c);
    ScratchGuard TFB_CAT(sg_, __COUNTER__)(c, stream);
    if (!polys) return TFB_OK;
    CHECK_PTR(in); CHECK_PTR(slots);
    return launch_ckks_decode(c, scale, in, slots, polys, (cudaStream_t)stream);
}
int tfb_sample_uniform(tfb_ctx* c, uint64_t seed, uint32_t stream_id, uint64_t* out, uint64_t polys, void* stream) {
    CHECK_CTX(c);
    if (!polys) return TFB_OK;
    CHECK_PTR(out);
    return launch_sample_uniform(c, seed, stream_id, out, polys, (cudaStream_t)stream);
}
int tfb_sample_gaussian(tfb_ctx* c, double sigma, uint64_t seed, uint32_t stream_id, uint64_t* out, uint64_t polys, void* stream) {
    CHECK_CTX(c);
    if (!polys) return TFB_OK;
    CHECK_PTR(out);
    return launch_sample_gaussian(c, sigma, seed, stream_id, out, polys, (cudaStream_t)stream);
}

// batch chunk so that the R_big intermediates (7 polys per pair) stay bounded
static u64 bfv_chunk(const tfb_ctx* cb, u64 batch) {
    const size_t per = 7 * (size_t)cb->L * cb->N * sizeof(u64);
    u64 ch = (u64)((size_t)(3ull << 30) / per);
    if (ch < 1) ch = 1;
    return ch < batch ? ch : batch;
}

// joint context Q u P' (P' = first K primes of cb) owned by the library, cached per (cq, cb, K)
struct JointKey {
    u64 a, b;
    int k;
    bool operator<(const JointKey& o) const { return a != o.a ? a < o.a : (b != o.b ? b < o.b : k < o.k); }
};
static std::map<JointKey, tfb_ctx*> g_joint;
static std::mutex g_joint_mu;
static void forget_joint(const tfb_ctx* c) {
    std::vector<tfb_ctx*> dead;
    {
        std::lock_guard<std::mutex> lk(g_joint_mu);
        for (auto it = g_joint.begin(); it != g_joint.end();) {
            if (it->first.a == c->uid || it->first.b == c->uid) {
                dead.push_back(it->second);
                it = g_joint.erase(it);
            } else
                ++it;
        }
    }
    for (tfb_ctx* j : dead) tfb_ctx_destroy(j);
}
static int joint_ctx(tfb_ctx* cq, tfb_ctx* cb, int K, tfb_ctx** out) {
    std::lock_guard<std::mutex> lk(g_joint_mu);
    JointKey key{cq->uid, cb->uid, K};
    auto it = g_joint.find(key);
    if (it != g_joint.end()) { *out = it->second; return TFB_OK; }
    std::vector<u64> q(cq->q), psi(cq->psi);
    q.insert(q.end(), cb->q.begin(), cb->q.begin() + K);
    psi.insert(psi.end(), cb->psi.begin(), cb->psi.begin() + K);
    tfb_ctx* j = nullptr;
    int rc = tfb_ctx_create(cq->device, cq->N, (uint32_t)q.size(), q.data(), psi.data(), &j);
    if (rc) return rc;
    g_joint[key] = j;
    *out = j;
    return TFB_OK;
}

extern bool g_force_generic;
extern int g_ntt_version;
int tfb_bfv_mul(tfb_ctx* cq, tfb_ctx* cb, uint64_t t, const uint64_t* c1, const uint64_t* c2, uint64_t* out, uint64_t batch, void* stream) {
    CHECK_CTX(cq); CHECK_CTX(cb);
    ScratchGuard TFB_CAT(sg_, __COUNTER__)(cq, stream);
    ScratchGuard TFB_CAT(sg_, __COUNTER__)(cb, stream);
    if (!batch) return TFB_OK;
    CHECK_PTR(c1); CHECK_PTR(c2); CHECK_PTR(out);
    if (cq->N != cb->N) { tfb_set_error("bfv_mul: ring degrees differ"); return TFB_EINVAL; }
    if (t == 0) { tfb_set_error("bfv_mul: plaintext modulus is zero"); return TFB_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t polyq = (size_t)cq->L * cq->N;
    int rc;
    const int K = g_force_generic ? 0 : fast_bfv_joint_k(cq, cb, t);
    if (K > 0) {
        // own extension basis Q u P' (rns_fast.cu "joint basis"): same integers, same result
        tfb_ctx* cj = nullptr;
        if ((rc = joint_ctx(cq, cb, K, &cj))) return rc;
        const size_t polyj = (size_t)cj->L * cj->N;
        const u64 ch = bfv_chunk(cj, batch);
        const size_t polyk = (size_t)K * cj->N;      // the K new residues of one polynomial
        // E (4 operand polynomials per pair, dual) | T (3 products per pair) | X (the expansions' new residues only)
        if ((rc = ws_reserve(cj, (7 * ch * polyj + 4 * ch * polyk) * sizeof(u64)))) return rc;
        u64* E1 = (u64*)cj->ws;
        u64* T = E1 + 4 * ch * polyj;
        u64* X = T + 3 * ch * polyj;
        const bool gather = cj->v3_ok && cj->logN >= 12 && cj->logN <= 14 && g_ntt_version == 3 && !g_ntt_force_harvey && g_ntt_max_mode >= 2;
        for (u64 b0 = 0; b0 < batch; b0 += ch) {
            const u64 nb = batch - b0 < ch ? batch - b0 : ch;
            const bool square = c1 == c2;            // c*c: expand and transform the operand once
            u64* E2 = square ? E1 : E1 + 2 * nb * polyj;
            const u64* a1 = c1 + b0 * 2 * polyq;
            const u64* a2 = c2 + b0 * 2 * polyq;
            if (gather) {
                // the expansions write only their K new rows; the transform takes the Q rows from the caller's ciphertexts
                u64* X2 = square ? X : X + 2 * nb * polyk;
                if ((rc = fast_expand_joint(cq, cb, K, a1, X, 2 * nb, st, false))) return rc;
                if (!square && (rc = fast_expand_joint(cq, cb, K, a2, X2, 2 * nb, st, false))) return rc;
                rc = launch_ntt_gather(cj, a1, a2, (u32)(2 * nb), cq->L, X, E1, (square ? 2 : 4) * nb, st);
                if (rc) return rc == -1 ? (tfb_set_error("internal: gather transform unavailable"), TFB_EINVAL) : rc;
                if ((rc = launch_tensor_dual(cj, E1, E2, T, nb, st))) return rc;
                if ((rc = launch_ntt(cj, T, T, 3 * nb * cj->L, true, st))) return rc;
                if ((rc = fast_contract_joint(cq, cb, K, t, T, out + b0 * 3 * polyq, 3 * nb, st))) return rc;
                continue;
            }
            if ((rc = fast_expand_joint(cq, cb, K, a1, E1, 2 * nb, st))) return rc;
            if (!square && (rc = fast_expand_joint(cq, cb, K, a2, E2, 2 * nb, st))) return rc;
            if ((rc = launch_ntt(cj, E1, E1, (square ? 2 : 4) * nb * cj->L, false, st))) return rc;   // E1 and E2 are contiguous
            if ((rc = launch_tensor_dual(cj, E1, E2, T, nb, st))) return rc;
            if ((rc = launch_ntt(cj, T, T, 3 * nb * cj->L, true, st))) return rc;
            if ((rc = fast_contract_joint(cq, cb, K, t, T, out + b0 * 3 * polyq, 3 * nb, st))) return rc;
        }
        return TFB_OK;
    }
    const size_t polyb = (size_t)cb->L * cb->N;
    const u64 ch = bfv_chunk(cb, batch);
    rc = ws_reserve(cb, 7 * ch * polyb * sizeof(u64));
    if (rc) return rc;
    u64* E1 = (u64*)cb->ws;
    u64* T = E1 + 4 * ch * polyb;
    for (u64 b0 = 0; b0 < batch; b0 += ch) {
        const u64 nb = batch - b0 < ch ? batch - b0 : ch;
        u64* E2 = E1 + 2 * nb * polyb;
        if ((rc = launch_base_switch(cq, cb, c1 + b0 * 2 * polyq, E1, 2 * nb, st))) return rc;
        if ((rc = launch_base_switch(cq, cb, c2 + b0 * 2 * polyq, E2, 2 * nb, st))) return rc;
        if ((rc = launch_ntt(cb, E1, E1, 4 * nb * cb->L, false, st))) return rc;   // E1 and E2 are contiguous
        if ((rc = launch_tensor_dual(cb, E1, E2, T, nb, st))) return rc;
        if ((rc = launch_ntt(cb, T, T, 3 * nb * cb->L, true, st))) return rc;
        if ((rc = launch_bfv_contract(cq, cb, t, T, out + b0 * 3 * polyq, 3 * nb, st))) return rc;
    }
    return TFB_OK;
}

// ------------------------------------------------------------------- keyswitch
int tfb_keyswitch_digits(tfb_ctx* c, tfb_ctx* target, uint32_t w, const uint64_t* cend, uint64_t* out, uint64_t batch, void* stream) {
    CHECK_CTX(c); CHECK_CTX(target);
    if (!batch) return TFB_OK;
    CHECK_PTR(cend); CHECK_PTR(out);
    uint32_t D = c->L;
    if (w) {
        int rc = tfb_ndigits(c->q.data(), c->L, w, &D);
        if (rc) return rc;
    }
    return launch_ks_digits(c, target, (int)w, cend, (u64)c->L * c->N, out, 0, D, batch, (cudaStream_t)stream);
}

// digit polynomials k0 .. k0+dn-1 of cend, in the NTT domain of ring r: dig [batch][dn][r->L][N].  Base-2^w digits are
// small integers (below every prime when 2^w <= min q): they are written ONCE as [batch][dn][N] and the forward kernel
// reads each row under all r->L primes (launch_ntt_bcast) instead of materialising r->L identical copies first.
static int ks_digits_dual(tfb_ctx* c, tfb_ctx* r, uint32_t w, const u64* cend, u64 ct_stride, u64* dig, u32 k0, u32 dn, u64 batch,
                          cudaStream_t st) {
    int rc;
    if (w == 0) {   // CRT digits of 2^15-position rows: extraction fused with the first global level of the transform
        rc = launch_ks_crt_ntt(c, r, cend, ct_stride, dig, k0, dn, batch, st);
        if (rc != -1) return rc;
        // N = 2^12 .. 2^14: digits formed in the transform's load phase (no digit rows written and re-read)
        rc = g_force_generic ? -1 : launch_ntt_crt(c, r, cend, ct_stride, dig, k0, dn, batch, st);
        if (rc != -1) return rc;
    }
    bool small = w > 0 && w < 62;
    for (u32 i = 0; small && i < r->L; i++) small = (1ull << w) <= r->q[i];
    if (small && !g_force_generic && r->v3_ok && r->logN >= 12 && r->logN <= 14 && c->N == r->N) {
        // base-2^w digits cut out of the binary limbs of the integers while the transform loads its row
        const size_t need = (size_t)batch * c->L * c->N * sizeof(u64);
        if ((rc = ws_reserve(r, need))) return rc;
        u64* limbs = (u64*)r->ws;
        if ((rc = launch_ks_limbs(c, cend, ct_stride, limbs, batch, st))) return rc;   // (once per digit chunk: chunks are 4 GiB of digit rows)
        rc = launch_ntt_pow2(r, limbs, c->L, w, dig, k0, dn, batch, st);
        if (rc != -1) return rc;
    }
    if (small && r->L > 1 && r->v3_ok && r->logN >= 12 && r->logN <= 14) {
        const size_t need = (size_t)batch * dn * r->N * sizeof(u64);
        if ((rc = ws_reserve(r, need))) return rc;
        u64* compact = (u64*)r->ws;
        if ((rc = launch_ks_digits(c, r, (int)w, cend, ct_stride, compact, k0, dn, batch, st, true))) return rc;
        rc = launch_ntt_bcast(r, compact, dig, (u64)batch * dn, st);
        if (rc != -1) return rc;
    }
    if ((rc = launch_ks_digits(c, r, (int)w, cend, ct_stride, dig, k0, dn, batch, st))) return rc;
    return launch_ntt(r, dig, dig, (u64)batch * dn * r->L, false, st);
}

int tfb_keyswitch(tfb_ctx* c, tfb_ctx* ext, uint32_t w, const uint64_t* key_dual, uint32_t D, const uint64_t* ct,
                  uint32_t comps, uint64_t* out, uint64_t batch, void* stream) {
    CHECK_CTX(c);
    ScratchGuard TFB_CAT(sg_, __COUNTER__)(c, stream);
    ScratchGuard TFB_CAT(sg_, __COUNTER__)(ext ? ext : c, stream);
    if (!batch) return TFB_OK;
    CHECK_PTR(key_dual); CHECK_PTR(ct); CHECK_PTR(out);
    if (comps != 2 && comps != 3) { tfb_set_error("keyswitch: ciphertext must have 2 or 3 components"); return TFB_EINVAL; }
    uint32_t Dneed = c->L;
    if (w) {
        int rc = tfb_ndigits(c->q.data(), c->L, w, &Dneed);
        if (rc) return rc;
    }
    if (D < Dneed) { tfb_set_error("keyswitch: evaluation key has too few digit components"); return TFB_EINVAL; }
    tfb_ctx* r = ext ? ext : c;  // ring the accumulation happens in
    if (ext) {
        if (ext->N != c->N || ext->L != c->L + 1) { tfb_set_error("keyswitch: raised ring must be the ciphertext primes plus one special prime"); return TFB_EINVAL; }
        for (u32 i = 0; i < c->L; i++)
            if (ext->q[i] != c->q[i]) { tfb_set_error("keyswitch: raised ring must start with the ciphertext primes"); return TFB_EINVAL; }
    }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t polyr = (size_t)r->L * r->N, polyc = (size_t)c->L * c->N;
    // digit chunk so the materialised digit polynomials stay below 4 GiB (of 180 GB): fewer, longer accumulation passes
    u32 dch = (u32)((size_t)(4ull << 30) / (batch * polyr * sizeof(u64)));
    if (dch < 1) dch = 1;
    if (dch > Dneed) dch = Dneed;
    int rc = stage_reserve(r, (2 * batch * polyr + (size_t)dch * batch * polyr) * sizeof(u64));
    if (rc) return rc;
    u64* acc = (u64*)r->stage;
    u64* dig = acc + 2 * batch * polyr;
    const u64* cend = ct + (size_t)(comps - 1) * polyc;
    for (u32 k0 = 0; k0 < Dneed; k0 += dch) {
        const u32 dn = Dneed - k0 < dch ? Dneed - k0 : dch;
        if ((rc = ks_digits_dual(c, r, w, cend, (u64)comps * polyc, dig, k0, dn, batch, st))) return rc;
        if ((rc = launch_ks_accum(r, k0, dn, dig, key_dual, acc, k0 ? 1 : 0, batch, st))) return rc;
    }
    if ((rc = launch_ntt(r, acc, acc, 2 * batch * r->L, true, st))) return rc;
    if (ext) return launch_ks_finish_raised(c, ext, ct, comps, acc, out, batch, st);
    return launch_ks_finish(c, ct, comps, acc, out, batch, st);
}

// everything of the sharded keyswitch up to the inverse transform: *acc_out [batch][2][Ls][N] (primal) = this rank's rows of
// sum_k digit_k * key_k; the caller adds the ciphertext's own components (launch_ks_finish, or the peer-push epilogue)
static int keyswitch_shard_acc(tfb_ctx* c, tfb_ctx* r, uint32_t first, uint32_t w, const uint64_t* key_dual, uint32_t D, const uint64_t* ct,
                        uint32_t comps, uint64_t batch, cudaStream_t st, u64** acc_out) {
    CHECK_PTR(key_dual); CHECK_PTR(ct);
    if (comps != 2 && comps != 3) { tfb_set_error("keyswitch: ciphertext must have 2 or 3 components"); return TFB_EINVAL; }
    if (r->N != c->N || r->L == 0 || (u64)first + r->L > c->L) { tfb_set_error("keyswitch_shard: shard primes out of range"); return TFB_EINVAL; }
    for (u32 i = 0; i < r->L; i++)
        if (r->q[i] != c->q[first + i]) { tfb_set_error("keyswitch_shard: shard ring must hold the primes first.. of the ciphertext ring"); return TFB_EINVAL; }
    uint32_t Dneed = c->L;
    if (w) {
        int rc = tfb_ndigits(c->q.data(), c->L, w, &Dneed);
        if (rc) return rc;
    }
    if (D < Dneed) { tfb_set_error("keyswitch: evaluation key has too few digit components"); return TFB_EINVAL; }
    const size_t polyr = (size_t)r->L * r->N, polyc = (size_t)c->L * c->N;
    u32 dch = (u32)((size_t)(4ull << 30) / (batch * polyr * sizeof(u64)));
    if (dch < 1) dch = 1;
    if (dch > Dneed) dch = Dneed;
    int rc = stage_reserve(r, (2 * batch * polyr + (size_t)dch * batch * polyr) * sizeof(u64));
    if (rc) return rc;
    u64* acc = (u64*)r->stage;
    u64* dig = acc + 2 * batch * polyr;
    const u64* cend = ct + (size_t)(comps - 1) * polyc;
    for (u32 k0 = 0; k0 < Dneed; k0 += dch) {
        const u32 dn = Dneed - k0 < dch ? Dneed - k0 : dch;
        if ((rc = ks_digits_dual(c, r, w, cend, (u64)comps * polyc, dig, k0, dn, batch, st))) return rc;   // digits of the WHOLE integer under the shard's primes
        if ((rc = launch_ks_accum(r, k0, dn, dig, key_dual, acc, k0 ? 1 : 0, batch, st))) return rc;
    }
    if ((rc = launch_ntt(r, acc, acc, 2 * batch * r->L, true, st))) return rc;
    *acc_out = acc;
    return TFB_OK;
}

int tfb_keyswitch_shard(tfb_ctx* c, tfb_ctx* r, uint32_t first, uint32_t w, const uint64_t* key_dual, uint32_t D, const uint64_t* ct,
                        uint32_t comps, uint64_t* out, uint64_t batch, void* stream) {
    CHECK_CTX(c); CHECK_CTX(r);
    ScratchGuard TFB_CAT(sg_, __COUNTER__)(c, stream);
    ScratchGuard TFB_CAT(sg_, __COUNTER__)(r, stream);
    if (!batch) return TFB_OK;
    CHECK_PTR(out);
    cudaStream_t st = (cudaStream_t)stream;
    u64* acc;
    int rc = keyswitch_shard_acc(c, r, first, w, key_dual, D, ct, comps, batch, st, &acc);
    if (rc) return rc;
    return launch_ks_finish(r, ct, comps, acc, out, batch, st, c->L, first);
}

// ---------------------------------------------------------- peer exchange (sharded keyswitch, one process per GPU)
// Local allocation: two result slots (a call writes slot epoch & 1 on every rank, so a rank that runs one call ahead never
// overwrites rows a slower peer is still reading) + the flag words + the CTA counter and the error word.
struct tfb_xchg {
    tfb_ctx* ctx;
    u32 rank, world;
    size_t slot_words;
    u64* local;                       // cudaMalloc: [2][slot_words] results, then TFB_MAX_PEERS flag words, then count, err
    u64* peer[TFB_MAX_PEERS];         // the same allocation of every rank, mapped here (peer[rank] = local)
    bool ipc[TFB_MAX_PEERS];
    u64 epoch;
};
static u64* xchg_flags(const tfb_xchg* x, u32 p) { return x->peer[p] + 2 * x->slot_words; }

int tfb_xchg_create(tfb_ctx* c, uint32_t rank, uint32_t world, uint64_t slot_bytes, tfb_xchg** out) {
    CHECK_CTX(c); CHECK_PTR(out);
    if (world < 1 || world > TFB_MAX_PEERS || rank >= world || slot_bytes == 0 || slot_bytes % 8) { tfb_set_error("xchg_create: need rank < world <= 8 and a slot size in whole words"); return TFB_EINVAL; }
    tfb_xchg* x = new (std::nothrow) tfb_xchg();
    if (!x) return TFB_ENOMEM;
    x->ctx = c; x->rank = rank; x->world = world; x->slot_words = slot_bytes / 8; x->epoch = 0;
    for (u32 p = 0; p < TFB_MAX_PEERS; p++) { x->peer[p] = nullptr; x->ipc[p] = false; }
    const size_t bytes = 2 * slot_bytes + (TFB_MAX_PEERS + 2) * sizeof(u64);
    cudaError_t e = cudaMalloc((void**)&x->local, bytes);
    if (e != cudaSuccess) { delete x; return tfb_cuda_fail(e, "cudaMalloc(exchange buffer)"); }
    e = cudaMemset(x->local, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();      // zeroed before any peer can learn the address
    if (e != cudaSuccess) { cudaFree(x->local); delete x; return tfb_cuda_fail(e, "cudaMemset(exchange buffer)"); }
    x->peer[rank] = x->local;
    *out = x;
    return TFB_OK;
}
int tfb_xchg_export(tfb_xchg* x, uint8_t handle[64]) {
    if (!x) { tfb_set_error("null exchange"); return TFB_EINVAL; }
    CHECK_CTX(x->ctx); CHECK_PTR(handle);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    TFB_CUDA(cudaIpcGetMemHandle(&h, x->local));
    memcpy(handle, &h, 64);
    return TFB_OK;
}
int tfb_xchg_attach_ipc(tfb_xchg* x, uint32_t peer, const uint8_t handle[64]) {
    if (!x) { tfb_set_error("null exchange"); return TFB_EINVAL; }
    CHECK_CTX(x->ctx); CHECK_PTR(handle);
    if (peer >= x->world || peer == x->rank || x->peer[peer]) { tfb_set_error("xchg_attach: bad or repeated peer"); return TFB_EINVAL; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void* p = nullptr;
    TFB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    x->peer[peer] = (u64*)p;
    x->ipc[peer] = true;
    return TFB_OK;
}
/* same process (threads, or several exchanges in one test process): the peer's local base pointer itself */
int tfb_xchg_attach_ptr(tfb_xchg* x, uint32_t peer, void* base) {
    if (!x) { tfb_set_error("null exchange"); return TFB_EINVAL; }
    CHECK_CTX(x->ctx); CHECK_PTR(base);
    if (peer >= x->world || peer == x->rank || x->peer[peer]) { tfb_set_error("xchg_attach: bad or repeated peer"); return TFB_EINVAL; }
    cudaPointerAttributes at;
    TFB_CUDA(cudaPointerGetAttributes(&at, base));
    if (at.type != cudaMemoryTypeDevice) { tfb_set_error("xchg_attach: not a device pointer"); return TFB_EINVAL; }
    if (at.device != x->ctx->device) {
        int can = 0;
        TFB_CUDA(cudaDeviceCanAccessPeer(&can, x->ctx->device, at.device));
        if (!can) { tfb_set_error("xchg_attach: no peer access between the two devices"); return TFB_EUNSUPPORTED; }
        const cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return tfb_cuda_fail(e, "cudaDeviceEnablePeerAccess");
        (void)cudaGetLastError();
    }
    x->peer[peer] = (u64*)base;
    return TFB_OK;
}
int tfb_xchg_local(tfb_xchg* x, void** base) {
    if (!x || !base) { tfb_set_error("null exchange"); return TFB_EINVAL; }
    *base = x->local;
    return TFB_OK;
}
/* the caller makes sure (barrier) that no peer still writes into this rank's buffer */
int tfb_xchg_destroy(tfb_xchg* x) {
    if (!x) return TFB_OK;
    CHECK_CTX(x->ctx);
    cudaDeviceSynchronize();
    for (u32 p = 0; p < x->world; p++)
        if (x->ipc[p]) cudaIpcCloseMemHandle(x->peer[p]);
    cudaFree(x->local);
    delete x;
    return TFB_OK;
}

int tfb_keyswitch_shard_push(tfb_ctx* c, tfb_ctx* r, uint32_t first, uint32_t w, const uint64_t* key_dual, uint32_t D, const uint64_t* ct,
                             uint32_t comps, tfb_xchg* x, uint64_t** result, uint64_t batch, void* stream) {
    CHECK_CTX(c); CHECK_CTX(r);
    if (!x || x->ctx->device != c->device) { tfb_set_error("keyswitch_shard_push: exchange belongs to another device"); return TFB_EINVAL; }
    CHECK_PTR(result);
    ScratchGuard TFB_CAT(sg_, __COUNTER__)(c, stream);
    ScratchGuard TFB_CAT(sg_, __COUNTER__)(r, stream);
    if (!batch) { tfb_set_error("keyswitch_shard_push: empty batch (every rank must take part in the exchange)"); return TFB_EINVAL; }
    if (batch * 2 * c->L * c->N > x->slot_words) { tfb_set_error("keyswitch_shard_push: batch larger than the exchange slot"); return TFB_EINVAL; }
    for (u32 p = 0; p < x->world; p++)
        if (!x->peer[p]) { tfb_set_error("keyswitch_shard_push: a peer is not attached"); return TFB_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    u64* acc;
    int rc = keyswitch_shard_acc(c, r, first, w, key_dual, D, ct, comps, batch, st, &acc);
    if (rc) return rc;
    const u64 epoch = ++x->epoch;
    u64 *outs[TFB_MAX_PEERS], *flags[TFB_MAX_PEERS];
    for (u32 p = 0; p < x->world; p++) { outs[p] = x->peer[p] + (epoch & 1) * x->slot_words; flags[p] = xchg_flags(x, p); }
    u32* tail = (u32*)(x->local + 2 * x->slot_words + TFB_MAX_PEERS);
    rc = launch_ks_finish_push(r, ct, comps, acc, batch, st, c->L, first, outs, flags, tail, tail + 2, x->rank, x->world, epoch);
    if (rc) return rc;
    *result = x->local + (epoch & 1) * x->slot_words;
    return TFB_OK;
}
/* 1 when a wait of an earlier push timed out (a peer never arrived); synchronises the stream's device first */
int tfb_xchg_check(tfb_xchg* x, int* timed_out) {
    if (!x || !timed_out) { tfb_set_error("null exchange"); return TFB_EINVAL; }
    CHECK_CTX(x->ctx);
    u32 e = 0;
    TFB_CUDA(cudaDeviceSynchronize());
    TFB_CUDA(cudaMemcpy(&e, (u32*)(x->local + 2 * x->slot_words + TFB_MAX_PEERS) + 2, sizeof(u32), cudaMemcpyDeviceToHost));
    *timed_out = (int)e;
    return TFB_OK;
}

// ---------------------------------------------------------- host-buffer variants
#define H2D(dst, src, words) TFB_CUDA(cudaMemcpyAsync(dst, src, (words) * sizeof(u64), cudaMemcpyHostToDevice, st))
#define D2H(dst, src, words) TFB_CUDA(cudaMemcpyAsync(dst, src, (words) * sizeof(u64), cudaMemcpyDeviceToHost, st))

// host-call I/O buffer: a third growable allocation per context (ws and stage are
// used by the device-side composites the host variants call into)
static int io_buf(tfb_ctx* c, size_t words, u64** p) {
    int rc = grow(&c->io, &c->io_bytes, words * sizeof(u64));
    if (rc) return rc;
    *p = (u64*)c->io;
    return TFB_OK;
}

int tfb_ntt_fwd_host(tfb_ctx* c, const uint64_t* in, uint64_t* out, uint64_t rows, void* stream) {
    CHECK_CTX(c); CHECK_ROWS(c, rows);
    if (!rows) return TFB_OK;
    CHECK_PTR(in); CHECK_PTR(out);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t words = (size_t)rows * c->N;
    u64* d;
    int rc = io_buf(c, words, &d);
    if (rc) return rc;
    H2D(d, in, words);
    if ((rc = launch_ntt(c, d, d, rows, false, st))) return rc;
    D2H(out, d, words);
    TFB_CUDA(cudaStreamSynchronize(st));
    return TFB_OK;
}
int tfb_ntt_inv_host(tfb_ctx* c, const uint64_t* in, uint64_t* out, uint64_t rows, void* stream) {
    CHECK_CTX(c); CHECK_ROWS(c, rows);
    if (!rows) return TFB_OK;
    CHECK_PTR(in); CHECK_PTR(out);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t words = (size_t)rows * c->N;
    u64* d;
    int rc = io_buf(c, words, &d);
    if (rc) return rc;
    H2D(d, in, words);
    if ((rc = launch_ntt(c, d, d, rows, true, st))) return rc;
    D2H(out, d, words);
    TFB_CUDA(cudaStreamSynchronize(st));
    return TFB_OK;
}
int tfb_ring_mul_host(tfb_ctx* c, const uint64_t* a, const uint64_t* b, uint64_t* out, uint64_t rows, void* stream) {
    CHECK_CTX(c); CHECK_ROWS(c, rows);
    if (!rows) return TFB_OK;
    CHECK_PTR(a); CHECK_PTR(b); CHECK_PTR(out);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t words = (size_t)rows * c->N;
    u64* d;
    int rc = io_buf(c, 2 * words, &d);
    if (rc) return rc;
    H2D(d, a, words);
    H2D(d + words, b, words);
    if ((rc = tfb_ring_mul(c, d, d + words, d, rows, stream))) return rc;
    D2H(out, d, words);
    TFB_CUDA(cudaStreamSynchronize(st));
    return TFB_OK;
}
int tfb_ct_tensor_host(tfb_ctx* c, const uint64_t* c1, const uint64_t* c2, uint64_t* out, uint64_t batch, void* stream) {
    CHECK_CTX(c);
    if (!batch) return TFB_OK;
    CHECK_PTR(c1); CHECK_PTR(c2); CHECK_PTR(out);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t poly = (size_t)c->L * c->N;
    u64* d;
    int rc = io_buf(c, 7 * batch * poly, &d);
    if (rc) return rc;
    u64 *d1 = d, *d2 = d + 2 * batch * poly, *dout = d + 4 * batch * poly;
    H2D(d1, c1, 2 * batch * poly);
    H2D(d2, c2, 2 * batch * poly);
    if ((rc = ct_tensor_dev(c, d1, d2, dout, batch, st))) return rc;
    D2H(out, dout, 3 * batch * poly);
    TFB_CUDA(cudaStreamSynchronize(st));
    return TFB_OK;
}
// Host-buffer BFV multiply, software-pipelined over chunks of the batch: copy-in, compute and
// copy-out run on three internal streams with double-buffered device slots, so the two PCIe
// directions and the kernels overlap (the one-shot version spent ~80% of its time in the copies).
struct HostPipe {
    cudaStream_t in = nullptr, cmp = nullptr, out = nullptr;
    cudaEvent_t h2d_done[2] = {nullptr, nullptr}, cmp_done[2] = {nullptr, nullptr}, d2h_done[2] = {nullptr, nullptr};
    cudaEvent_t entry = nullptr, exit_ = nullptr;
    bool ok = false;
};
static std::map<u64, HostPipe> g_pipes;   // by context uid
static std::mutex g_pipes_mu;
static int host_pipe(tfb_ctx* c, HostPipe** out) {
    std::lock_guard<std::mutex> lk(g_pipes_mu);
    HostPipe& p = g_pipes[c->uid];
    if (!p.ok) {
        TFB_CUDA(cudaStreamCreateWithFlags(&p.in, cudaStreamNonBlocking));
        TFB_CUDA(cudaStreamCreateWithFlags(&p.cmp, cudaStreamNonBlocking));
        TFB_CUDA(cudaStreamCreateWithFlags(&p.out, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            TFB_CUDA(cudaEventCreateWithFlags(&p.h2d_done[i], cudaEventDisableTiming));
            TFB_CUDA(cudaEventCreateWithFlags(&p.cmp_done[i], cudaEventDisableTiming));
            TFB_CUDA(cudaEventCreateWithFlags(&p.d2h_done[i], cudaEventDisableTiming));
        }
        TFB_CUDA(cudaEventCreateWithFlags(&p.entry, cudaEventDisableTiming));
        TFB_CUDA(cudaEventCreateWithFlags(&p.exit_, cudaEventDisableTiming));
        p.ok = true;
    }
    *out = &p;
    return TFB_OK;
}
static void forget_pipe(const tfb_ctx* c) {
    std::lock_guard<std::mutex> lk(g_pipes_mu);
    auto it = g_pipes.find(c->uid);
    if (it == g_pipes.end()) return;
    HostPipe& p = it->second;
    if (p.ok) {
        cudaStreamDestroy(p.in); cudaStreamDestroy(p.cmp); cudaStreamDestroy(p.out);
        for (int i = 0; i < 2; i++) { cudaEventDestroy(p.h2d_done[i]); cudaEventDestroy(p.cmp_done[i]); cudaEventDestroy(p.d2h_done[i]); }
        cudaEventDestroy(p.entry); cudaEventDestroy(p.exit_);
    }
    g_pipes.erase(it);
}

int tfb_bfv_mul_host(tfb_ctx* cq, tfb_ctx* cb, uint64_t t, const uint64_t* c1, const uint64_t* c2, uint64_t* out, uint64_t batch, void* stream) {
    CHECK_CTX(cq); CHECK_CTX(cb);
    if (!batch) return TFB_OK;
    CHECK_PTR(c1); CHECK_PTR(c2); CHECK_PTR(out);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t poly = (size_t)cq->L * cq->N;
    // chunk: about 32 MiB of input per copy, at least one pair
    static const size_t chunk_bytes = [] {   // TFB_HOST_CHUNK_MIB: input bytes per pipelined copy (tuning knob, default 32 MiB)
        const char* e = getenv("TFB_HOST_CHUNK_MIB");
        const long v = e ? atol(e) : 0;
        return (size_t)(v > 0 && v <= 1024 ? v : 32) << 20;
    }();
    u64 ch = (u64)(chunk_bytes / (4 * poly * sizeof(u64)));
    if (ch < 1) ch = 1;
    if (ch > batch) ch = batch;
    HostPipe* P;
    int rc = host_pipe(cq, &P);
    if (rc) return rc;
    u64* d;
    if ((rc = io_buf(cq, 2 * 7 * ch * poly, &d))) return rc;
    u64* slot_in[2] = {d, d + 4 * ch * poly};
    u64* slot_out[2] = {d + 8 * ch * poly, d + 11 * ch * poly};
    TFB_CUDA(cudaEventRecord(P->entry, st));
    TFB_CUDA(cudaStreamWaitEvent(P->in, P->entry, 0));
    TFB_CUDA(cudaStreamWaitEvent(P->cmp, P->entry, 0));
    TFB_CUDA(cudaStreamWaitEvent(P->out, P->entry, 0));
    u64 idx = 0;
    for (u64 b0 = 0; b0 < batch; b0 += ch, idx++) {
        const u64 nb = batch - b0 < ch ? batch - b0 : ch;
        const int s = (int)(idx & 1);
        if (idx >= 2) TFB_CUDA(cudaStreamWaitEvent(P->in, P->cmp_done[s], 0));   // slot's previous compute has read it
        TFB_CUDA(cudaMemcpyAsync(slot_in[s], c1 + b0 * 2 * poly, 2 * nb * poly * sizeof(u64), cudaMemcpyHostToDevice, P->in));
        TFB_CUDA(cudaMemcpyAsync(slot_in[s] + 2 * nb * poly, c2 + b0 * 2 * poly, 2 * nb * poly * sizeof(u64), cudaMemcpyHostToDevice, P->in));
        TFB_CUDA(cudaEventRecord(P->h2d_done[s], P->in));
        TFB_CUDA(cudaStreamWaitEvent(P->cmp, P->h2d_done[s], 0));
        if (idx >= 2) TFB_CUDA(cudaStreamWaitEvent(P->cmp, P->d2h_done[s], 0));  // slot's previous result has left
        if ((rc = tfb_bfv_mul(cq, cb, t, slot_in[s], slot_in[s] + 2 * nb * poly, slot_out[s], nb, (void*)P->cmp))) return rc;
        TFB_CUDA(cudaEventRecord(P->cmp_done[s], P->cmp));
        TFB_CUDA(cudaStreamWaitEvent(P->out, P->cmp_done[s], 0));
        TFB_CUDA(cudaMemcpyAsync(out + b0 * 3 * poly, slot_out[s], 3 * nb * poly * sizeof(u64), cudaMemcpyDeviceToHost, P->out));
        TFB_CUDA(cudaEventRecord(P->d2h_done[s], P->out));
    }
    TFB_CUDA(cudaEventRecord(P->exit_, P->out));
    TFB_CUDA(cudaStreamWaitEvent(st, P->exit_, 0));
    TFB_CUDA(cudaStreamSynchronize(P->out));
    TFB_CUDA(cudaStreamSynchronize(P->cmp));
    TFB_CUDA(cudaStreamSynchronize(P->in));
    TFB_CUDA(cudaStreamSynchronize(st));
    return TFB_OK;
}
int tfb_rescale_host(tfb_ctx* c, const uint64_t* in, uint64_t* out, uint64_t polys, void* stream) {
    CHECK_CTX(c);
    if (!polys) return TFB_OK;
    CHECK_PTR(in); CHECK_PTR(out);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t pin = (size_t)c->L * c->N, pout = (size_t)(c->L - 1) * c->N;
    u64* d;
    int rc = io_buf(c, polys * (pin + pout), &d);
    if (rc) return rc;
    H2D(d, in, polys * pin);
    if ((rc = launch_rescale(c, d, d + polys * pin, polys, st))) return rc;
    D2H(out, d + polys * pin, polys * pout);
    TFB_CUDA(cudaStreamSynchronize(st));
    return TFB_OK;
}
int tfb_bfv_encode_host(tfb_ctx* c, uint64_t t, const uint64_t* delta, uint32_t nl, const uint64_t* m, uint64_t* out, uint64_t polys, void* stream) {
    CHECK_CTX(c);
    if (!polys) return TFB_OK;
    CHECK_PTR(delta); CHECK_PTR(m); CHECK_PTR(out);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t pin = (size_t)c->N, pout = (size_t)c->L * c->N;
    u64* d;
    int rc = io_buf(c, polys * (pin + pout), &d);
    if (rc) return rc;
    H2D(d, m, polys * pin);
    if ((rc = launch_bfv_encode(c, t, delta, nl, d, d + polys * pin, polys, st))) return rc;
    D2H(out, d + polys * pin, polys * pout);
    TFB_CUDA(cudaStreamSynchronize(st));
    return TFB_OK;
}
int tfb_bfv_decode_host(tfb_ctx* c, uint64_t t, const uint64_t* delta, uint32_t nl, const uint64_t* b, uint64_t* out, uint64_t polys, void* stream) {
    CHECK_CTX(c);
    if (!polys) return TFB_OK;
    CHECK_PTR(delta); CHECK_PTR(b); CHECK_PTR(out);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t pin = (size_t)c->L * c->N, pout = (size_t)c->N;
    u64* d;
    int rc = io_buf(c, polys * (pin + pout), &d);
    if (rc) return rc;
    H2D(d, b, polys * pin);
    if ((rc = launch_bfv_decode(c, t, delta, nl, d, d + polys * pin, polys, st))) return rc;
    D2H(out, d + polys * pin, polys * pout);
    TFB_CUDA(cudaStreamSynchronize(st));
    return TFB_OK;
}

}  // extern "C"

