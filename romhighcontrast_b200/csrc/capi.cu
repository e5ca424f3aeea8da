// extern "C" surface of libromhc.so (declared in include/romhc.h).
#include "../../include/romhc.h"
#include "romhc_internal.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <new>
#include <thread>
#include <vector>

namespace romhc {
extern int g_gram_variant;
extern int g_tn_variant;
extern int g_sweep_variant;

static thread_local char g_err[1024] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* last_error() { return g_err; }

void Context::release() {
    if (scratch) cudaFree(scratch);
    if (ws_base) cudaFree(ws_base);
    if (ws_flags) cudaFree(ws_flags);
    if (h_flags) cudaFreeHost(h_flags);
    for (int* p : tile_rowv) cudaFree(p);
    for (int* p : tile_colv) cudaFree(p);
    tile_rowv.clear(); tile_colv.clear(); tile_ready = false; kernels_configured = false;
    for (auto& kv : tile_rinfo_cache) cudaFree(kv.second);
    tile_rinfo_cache.clear();
    scratch = nullptr; ws_base = nullptr; ws_flags = nullptr; h_flags = nullptr;
    scratch_bytes = 0; ws_K = 0;
    free_host_stage();
    for (int i = 0; i < 5; ++i) { if (rg_buf[i]) cudaFree(rg_buf[i]); rg_buf[i] = nullptr; rg_cap[i] = 0; }
}

void Context::free_host_stage() {
    for (int i = 0; i < 2; ++i) {
        cudaFree(hstage.y[i]); cudaFree(hstage.x[i]); cudaFree(hstage.u[i]); cudaFree(hstage.it[i]); cudaFree(hstage.rel[i]);
        hstage.y[i] = hstage.x[i] = hstage.u[i] = hstage.rel[i] = nullptr; hstage.it[i] = nullptr;
        if (hstage.done[i]) cudaEventDestroy(hstage.done[i]);
        if (hstage.copied[i]) cudaEventDestroy(hstage.copied[i]);
        hstage.done[i] = hstage.copied[i] = nullptr;
    }
    for (int i = 0; i < 2; ++i) { if (hstage.bounce[i]) cudaFreeHost(hstage.bounce[i]); hstage.bounce[i] = nullptr; }
    hstage.bounce_cap = 0;
    if (hstage.copy2) cudaStreamDestroy(hstage.copy2);
    hstage.copy2 = nullptr;
    for (int i = 0; i < 2; ++i) {
        if (hstage.xfer_dev[i]) cudaFree(hstage.xfer_dev[i]);
        hstage.xfer_dev[i] = nullptr;
        if (hstage.xfer_done[i]) cudaEventDestroy(hstage.xfer_done[i]);
        if (hstage.xfer_ready[i]) cudaEventDestroy(hstage.xfer_ready[i]);
        hstage.xfer_done[i] = hstage.xfer_ready[i] = nullptr;
    }
    hstage.xfer_cap = 0;
    if (hstage.it_pin) cudaFreeHost(hstage.it_pin);
    if (hstage.rel_pin) cudaFreeHost(hstage.rel_pin);
    hstage.it_pin = nullptr; hstage.rel_pin = nullptr; hstage.pin_cap = 0;
    if (hstage.compute) cudaStreamDestroy(hstage.compute);
    if (hstage.copy) cudaStreamDestroy(hstage.copy);
    hstage.compute = hstage.copy = nullptr;
    hstage.cap = 0;
}

int Context::ensure_pinned_stats(int64_t K) {
    HostStage& s = hstage;
    if (K <= s.pin_cap) return ROMHC_OK;
    if (s.it_pin) cudaFreeHost(s.it_pin);
    if (s.rel_pin) cudaFreeHost(s.rel_pin);
    s.it_pin = nullptr; s.rel_pin = nullptr; s.pin_cap = 0;
    CK(cudaMallocHost((void**)&s.it_pin, size_t(K) * 4));
    CK(cudaMallocHost((void**)&s.rel_pin, size_t(K) * 8));
    s.pin_cap = K;
    return ROMHC_OK;
}

int Context::ensure_host_stage(int64_t chunk) {
    if (chunk <= hstage.cap) return ROMHC_OK;
    free_host_stage();
    const LevelGeo& g = levels[0];
    const int nb = nrb * ncb;
    const int64_t D = int64_t(g.R - 1) * (g.C - 1);
    CK(cudaStreamCreateWithFlags(&hstage.compute, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&hstage.copy, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        CK(cudaMalloc(&hstage.y[i], size_t(chunk) * nb * 8));
        CK(cudaMalloc(&hstage.x[i], size_t(chunk) * g.Dp * 8));
        CK(cudaMalloc(&hstage.u[i], size_t(chunk) * D * 8));
        CK(cudaMalloc(&hstage.rel[i], size_t(chunk) * 8));
        CK(cudaMalloc(&hstage.it[i], size_t(chunk) * 4));
        CK(cudaEventCreateWithFlags(&hstage.done[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&hstage.copied[i], cudaEventDisableTiming));
    }
    hstage.cap = chunk;
    return ROMHC_OK;
}

}  // namespace romhc

using namespace romhc;

struct romhc_context { Context c; };

#define H(h) (&(h)->c)
#define ST(s) ((cudaStream_t)(s))
#define CHECK_H(h) do { if (!(h)) { set_error("null context"); return ROMHC_ERR_ARG; } \
                        cudaError_t e__ = cudaSetDevice((h)->c.device); \
                        if (e__ != cudaSuccess) { set_error("cudaSetDevice: %s", cudaGetErrorString(e__)); return ROMHC_ERR_CUDA; } } while (0)

extern "C" {

int romhc_version(void) { return 100; }
const char* romhc_last_error(void) { return last_error(); }
int64_t romhc_launch_count(void) { return (int64_t)g_launches.load(); }

int romhc_create(int nrb, int ncb, int N, int device, romhc_handle* out) {
    if (!out) { set_error("out is null"); return ROMHC_ERR_ARG; }
    *out = nullptr;
    if (nrb < 1 || ncb < 1 || N < 1 || nrb * ncb > ROMHC_MAX_BLOCKS) {
        set_error("bad geometry (%d, %d), N=%d (need 1 <= nrb*ncb <= %d)", nrb, ncb, N, ROMHC_MAX_BLOCKS);
        return ROMHC_ERR_ARG;
    }
    if (nrb * N < 2 || ncb * N < 2) { set_error("mesh has no interior vertex"); return ROMHC_ERR_ARG; }
    if ((long long)(nrb * (long long)N + 1) * (ncb * (long long)N + 8) > (1LL << 30)) {
        set_error("mesh too large"); return ROMHC_ERR_ARG;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error("no CUDA device available (%s): the ROMHighContrast B200 path has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return ROMHC_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { set_error("device %d out of range (0..%d)", device, ndev - 1); return ROMHC_ERR_ARG; }
    CK(cudaSetDevice(device));
    romhc_context* h = new (std::nothrow) romhc_context();
    if (!h) { set_error("out of host memory"); return ROMHC_ERR_ARG; }
    h->c.nrb = nrb; h->c.ncb = ncb; h->c.N = N; h->c.device = device;
    h->c.build_levels();
    *out = h;
    return ROMHC_OK;
}

int romhc_destroy(romhc_handle h) {
    if (!h) return ROMHC_OK;
    cudaSetDevice(h->c.device);
    h->c.release();
    delete h;
    return ROMHC_OK;
}

int romhc_set_option(romhc_handle h, const char* name, double value) {
    if (!h || !name) { set_error("null argument"); return ROMHC_ERR_ARG; }
    Context* c = H(h);
    if (!strcmp(name, "rtol")) c->rtol = value;
    else if (!strcmp(name, "maxit")) c->maxit = (int)value;
    else if (!strcmp(name, "coarse_sweeps")) { c->coarse_sweeps = std::max(1, (int)value); c->build_levels(); }
    else if (!strcmp(name, "nu")) { c->nu = std::max(1, std::min(4, (int)value)); }
    else if (!strcmp(name, "nu_mid")) { c->nu_mid = std::max(0, std::min(4, (int)value)); }
    else if (!strcmp(name, "nu_tail")) { c->nu_tail = std::max(1, std::min(8, (int)value)); c->build_levels(); }
    else if (!strcmp(name, "tile")) c->use_tile = value != 0.0;
    else if (!strcmp(name, "bridge")) {
        // the hierarchy changes: drop everything sized by it (workspace, vertex-class and row-info tables)
        cudaSetDevice(c->device);
        cudaDeviceSynchronize();
        c->use_bridge = value != 0.0;
        if (c->ws_base) cudaFree(c->ws_base);
        c->ws_base = nullptr; c->ws_K = 0;
        for (int* q : c->tile_rowv) cudaFree(q);
        for (int* q : c->tile_colv) cudaFree(q);
        c->tile_rowv.clear(); c->tile_colv.clear(); c->tile_ready = false; c->kernels_configured = false;
        for (auto& kv : c->tile_rinfo_cache) cudaFree(kv.second);
        c->tile_rinfo_cache.clear();
        c->build_levels();
    }
    else if (!strcmp(name, "gram_variant")) romhc::g_gram_variant = (int)value;
    else if (!strcmp(name, "tn_variant")) romhc::g_tn_variant = (int)value;
    else if (!strcmp(name, "fused")) c->use_fused = value != 0.0;
    else if (!strcmp(name, "ws_guard_poke")) {                   // negative control of the detector: damage `value` bytes of a zone
        if (c->ws_base && !c->ws_gaps.empty()) {
            cudaSetDevice(c->device);
            const auto& gp = c->ws_gaps[c->ws_gaps.size() / 2];
            cudaMemset((double*)c->ws_base + gp.first, 0, (size_t)std::max(0.0, std::min(value, double(gp.second * 8))));
        }
    }
    else if (!strcmp(name, "ws_guard")) {
        cudaSetDevice(c->device);
        cudaDeviceSynchronize();
        c->ws_guard = std::max(0, std::min(1 << 20, (int)value));
        if (c->ws_base) cudaFree(c->ws_base);                    // the next solve lays the workspace out again
        c->ws_base = nullptr; c->ws_K = 0; c->ws_gaps.clear();
    }
    else if (!strcmp(name, "defer_x")) c->defer_x = value != 0.0;
    else if (!strcmp(name, "papply_pers")) c->papply_pers = std::max(0, std::min(2, (int)value));
    else if (!strcmp(name, "sweep")) { c->use_sweep = value != 0.0; if (value != 0.0) romhc::g_sweep_variant = (int)value >= 2 ? 2 : 1; }
    else if (!strcmp(name, "proj_variant")) c->proj_variant = (int)value;
    else if (!strcmp(name, "z32")) c->use_z32 = std::max(0, std::min(3, (int)value));
    else if (!strcmp(name, "tile_persistent")) c->tile_persistent = value != 0.0;
    else if (!strcmp(name, "tile_prefetch")) c->tile_prefetch = value != 0.0;
    else if (!strcmp(name, "tile_ty")) c->tile_ty_cap = std::max(4, std::min(64, ((int)value) & ~3));
    else if (!strcmp(name, "threads")) c->strip_threads = ((int)value >= 512) ? 512 : 256;
    else if (!strcmp(name, "strip_kb")) c->strip_budget = (size_t)std::max(16.0, std::min(227.0, value)) * 1024;
    else if (!strcmp(name, "workspace_gb")) c->ws_budget_bytes = (size_t)(value * double(1 << 30));
    else if (!strcmp(name, "profile")) { c->prof_on = value != 0.0; for (int i = 0; i < PROF_NKIND; ++i) { c->prof_ms[i] = 0; c->prof_n[i] = 0; } }
    else if (!strcmp(name, "host_chunks")) c->host_chunks = std::max(1, std::min(64, (int)value));
    else if (!strcmp(name, "check_every")) c->check_every = std::max(1, (int)value);
    else if (!strcmp(name, "min_check_iter")) c->min_check_iter = std::max(1, (int)value);
    else { set_error("unknown option '%s'", name); return ROMHC_ERR_ARG; }
    return ROMHC_OK;
}

int romhc_get_info(romhc_handle h, int64_t* info) {
    if (!h || !info) { set_error("null argument"); return ROMHC_ERR_ARG; }
    const Context* c = H(h);
    const LevelGeo& g = c->levels[0];
    memset(info, 0, 16 * sizeof(int64_t));
    info[0] = int64_t(g.R - 1) * (g.C - 1); info[1] = g.Dp; info[2] = g.P; info[3] = g.R; info[4] = g.C;
    info[5] = (int64_t)c->levels.size(); info[6] = c->tail_level; info[7] = c->coarse_D; info[8] = c->coarse_direct;
    info[9] = c->nrb; info[10] = c->ncb; info[11] = c->N; info[12] = (int64_t)c->solve_bytes_per_system();
    info[13] = (int64_t)c->tail_smem; info[14] = c->bridge_level;
    info[15] = c->bridge_level >= 0 ? c->levels[c->bridge_level + 1].N : 0;
    return ROMHC_OK;
}

int romhc_get_profile(romhc_handle h, double* ms8, int64_t* n8) {
    if (!h || !ms8 || !n8) { set_error("null argument"); return ROMHC_ERR_ARG; }
    for (int i = 0; i < 8; ++i) { ms8[i] = i < PROF_NKIND ? H(h)->prof_ms[i] : 0.0; n8[i] = i < PROF_NKIND ? H(h)->prof_n[i] : 0; }
    return ROMHC_OK;
}

int romhc_check_guards(romhc_handle h, int64_t* n_bad) {
    if (!h || !n_bad) { set_error("null argument"); return ROMHC_ERR_ARG; }
    cudaSetDevice(H(h)->device);
    return H(h)->check_guards(n_bad);
}

int romhc_malloc(void** p, size_t bytes) { CK(cudaMalloc(p, bytes ? bytes : 8)); return ROMHC_OK; }
int romhc_free(void* p) { if (p) CK(cudaFree(p)); return ROMHC_OK; }
int romhc_malloc_host(void** p, size_t bytes) { CK(cudaMallocHost(p, bytes ? bytes : 8)); return ROMHC_OK; }
int romhc_free_host(void* p) { if (p) CK(cudaFreeHost(p)); return ROMHC_OK; }
int romhc_memcpy_h2d(void* d, const void* s, size_t n, void* st) {
    CK(cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, ST(st))); return ROMHC_OK;
}
int romhc_memcpy_d2h(void* d, const void* s, size_t n, void* st) {
    CK(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, ST(st))); return ROMHC_OK;
}
int romhc_memset(void* d, int v, size_t n, void* st) { CK(cudaMemsetAsync(d, v, n, ST(st))); return ROMHC_OK; }
int romhc_stream_sync(void* st) { CK(cudaStreamSynchronize(ST(st))); return ROMHC_OK; }

int romhc_pack(romhc_handle h, const double* c, double* p, int64_t K, void* st) { CHECK_H(h); return H(h)->pack(c, p, K, ST(st)); }
int romhc_unpack(romhc_handle h, const double* p, double* c, int64_t K, void* st) { CHECK_H(h); return H(h)->unpack(p, c, K, ST(st)); }

int romhc_apply(romhc_handle h, const double* y, const double* u, double* out, int64_t K, void* st) {
    CHECK_H(h); return H(h)->apply(y, u, out, K, ST(st));
}
int romhc_energy_norm(romhc_handle h, const double* y, const double* u, int64_t K, double* out, void* st) {
    CHECK_H(h); return H(h)->energy(y, u, nullptr, nullptr, 0, out, K, 0, 1, ST(st));
}
int romhc_l2_norm(romhc_handle h, const double* u, int64_t K, double* out, void* st) {
    CHECK_H(h); return H(h)->energy(nullptr, u, nullptr, nullptr, 0, out, K, 1, 1, ST(st));
}
int romhc_error_norm(romhc_handle h, const double* U, const double* coef, const double* basis, int n, int64_t K,
                     double* out, void* st) {
    CHECK_H(h);
    if (n < 0 || n > 256) { set_error("error_norm: bad n"); return ROMHC_ERR_ARG; }
    if (n > 0 && K > 0 && H(h)->use_sweep) {
        const int rc = H(h)->error_sweep(U, coef, basis, n, out, K, ST(st));
        if (rc >= 0) return rc;                     // -1: wide mesh / large n -> the strip kernel below
    }
    return H(h)->energy(nullptr, U, n > 0 ? coef : nullptr, basis, n, out, K, 0, 1, ST(st));
}

int romhc_solve(romhc_handle h, const double* y, int64_t K, double* x, int* iters, double* relres, void* st,
                int64_t* stats4) {
    CHECK_H(h);
    SolveStats s{0, 0, 0};
    const int rc = H(h)->solve(y, K, x, iters, relres, ST(st), &s);
    if (stats4) { stats4[0] = s.launched_iterations; stats4[1] = s.chunks; stats4[2] = s.status; stats4[3] = (int64_t)H(h)->ws_bytes; }
    return rc;
}
int romhc_solve_rhs(romhc_handle h, const double* y, const double* rhs, int64_t K, double* x, int* iters, double* relres,
                    void* st, int64_t* stats4) {
    CHECK_H(h);
    if (!rhs) { set_error("solve_rhs: rhs is null"); return ROMHC_ERR_ARG; }
    SolveStats s{0, 0, 0};
    const int rc = H(h)->solve(y, K, x, iters, relres, ST(st), &s, rhs);
    if (stats4) { stats4[0] = s.launched_iterations; stats4[1] = s.chunks; stats4[2] = s.status; stats4[3] = (int64_t)H(h)->ws_bytes; }
    return rc;
}
int romhc_precond(romhc_handle h, const double* y, const double* r, double* z, int64_t K, void* st) {
    CHECK_H(h); return H(h)->precond(y, r, z, K, ST(st));
}

int romhc_project_operators(romhc_handle h, const double* basis, int n, double* Ahat, double* bhat, void* st) {
    CHECK_H(h); return H(h)->project_operators(basis, n, Ahat, bhat, ST(st));
}
int romhc_reduced_solve(const double* y, int nb, const double* Ahat, const double* rhs, int rps, int n, int64_t K,
                        double* C, int* info, void* st) {
    return reduced_solve(y, nb, Ahat, rhs, rps, n, K, C, info, ST(st));
}

int romhc_gemm_nt(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int64_t M,
                  int64_t N, int64_t Kd, int sym, void* st) { return gemm_nt(A, lda, B, ldb, C, ldc, M, N, Kd, sym, ST(st)); }
int romhc_gemm_nn(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int64_t M,
                  int64_t N, int64_t Kd, void* st) { return gemm_nn(A, lda, B, ldb, C, ldc, M, N, Kd, ST(st)); }
int romhc_gemm_tn(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int64_t M,
                  int64_t N, int64_t Kd, void* st) { return gemm_tn(A, lda, B, ldb, C, ldc, M, N, Kd, ST(st)); }
int romhc_column_mean(const double* X, int64_t ld, int64_t K, int64_t D, double* mean, void* st) {
    return column_mean(X, ld, K, D, mean, ST(st));
}
int romhc_center_rows(double* X, int64_t ld, int64_t K, int64_t D, const double* mean, void* st) {
    return center_rows(X, ld, K, D, mean, ST(st));
}

int romhc_evaluate(romhc_handle h, const double* pts, int m, const double* u, int64_t K, double* out, void* st) {
    CHECK_H(h); return H(h)->evaluate(pts, m, u, K, out, ST(st));
}
int romhc_interp_weights(romhc_handle h, const double* pts, int m, int* idx3, double* w3, void* st) {
    CHECK_H(h); return H(h)->interp_weights(pts, m, idx3, w3, ST(st));
}
int romhc_row_norms(const double* X, int64_t ld, int64_t K, int64_t D, double* out, void* st) {
    return row_norms(X, ld, K, D, out, ST(st));
}
int romhc_tsqr_r(const double* W, int64_t ld, int b, int64_t Dp, double* R, void* st) {
    if (!W || !R) { set_error("null argument"); return ROMHC_ERR_ARG; }
    return tsqr_r(W, ld, b, Dp, R, ST(st));
}
int romhc_row_dots(const double* X, int64_t ldx, const double* Y, int64_t ldy, int64_t K, int64_t D, double* out, void* st) {
    return row_dots(X, ldx, Y, ldy, K, D, out, ST(st));
}
int romhc_estimator(const double* c, int64_t K, int n, const double* ab, int nb, int invert, double* out, void* st) {
    return estimator_contract(c, K, n, ab, nb, invert, out, ST(st));
}
int romhc_argmax(const double* v, int64_t K, int64_t* idx, double* val, void* st) { return argmax_first(v, K, idx, val, ST(st)); }
int romhc_poly_features(const double* basis, int64_t ld, int n, int64_t D, const int* terms, int nterms, int degree,
                        double* out, int64_t ldo, void* st) {
    if (!basis || !terms || !out) { set_error("null argument"); return ROMHC_ERR_ARG; }
    return poly_features(basis, ld, n, D, terms, nterms, degree, out, ldo, ST(st));
}

// ---- host-buffer entry points --------------------------------------------------------------------------------------------
// Pageable destinations: a device-to-host copy into pageable memory blocks the calling thread behind the whole copy and
// would serialise the pipeline, so the solutions go to pinned bounce buffers first and a stream-ordered host callback
// moves them on with a few threads while the next chunk is being solved.
struct HostCopyTask { const char* src; char* dst; size_t bytes; int nthreads; };
// memcpy threads per rank for pageable destinations: the host cores are shared by the ranks of the box (torchrun exports
// LOCAL_WORLD_SIZE), so 8 ranks x 8 threads would oversubscribe a 32-core host; ROMHC_COPY_THREADS overrides
static int host_copy_threads() {
    if (const char* e = getenv("ROMHC_COPY_THREADS")) { const int v = atoi(e); if (v > 0) return std::min(v, 64); }
    int ranks = 1;
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, atoi(e));
    const int hw = std::max(1, int(std::thread::hardware_concurrency()));
    return std::max(1, std::min(8, hw / (2 * ranks)));
}
static void CUDART_CB host_copy_callback(void* p) {
    HostCopyTask* t = static_cast<HostCopyTask*>(p);
    const int nt = std::max(1, t->nthreads);
    const size_t per = ((t->bytes + nt - 1) / nt + 4095) & ~size_t(4095);
    std::vector<std::thread> th;
    for (int i = 1; i < nt; ++i) {
        const size_t o = per * i;
        if (o >= t->bytes) break;
        th.emplace_back([=]() { memcpy(t->dst + o, t->src + o, std::min(per, t->bytes - o)); });
    }
    memcpy(t->dst, t->src, std::min(per, t->bytes));
    for (auto& x : th) x.join();
    delete t;
}

int romhc_generate_solutions_host(romhc_handle h, const double* y_host, int64_t K, double* U_host, int* iters_host,
                                  double* relres_host) {
    CHECK_H(h);
    if (K <= 0) return ROMHC_OK;
    Context* c = H(h);
    const LevelGeo& g = c->levels[0];
    const int nb = c->nrb * c->ncb;
    const int64_t D = int64_t(g.R - 1) * (g.C - 1);
    // Chunked, double-buffered pipeline: the D2H copy of chunk i (copy stream) overlaps the solve of chunk i+1
    // (compute stream).  Staging buffers persist in the context.
    const size_t per = 2 * size_t(g.Dp + D) * 8 + c->solve_bytes_per_system();
    int64_t chunk = std::max<int64_t>(1, std::min<int64_t>({(int64_t)(c->ws_budget_bytes / per), (int64_t)32768, K}));
    if (K >= 4096) chunk = std::min<int64_t>(chunk, (K + c->host_chunks - 1) / c->host_chunks);
    // the schedule below merges a short remainder into the last chunk: staging capacity = chunk + small / 2
    const int64_t small = std::min<int64_t>(chunk, std::max<int64_t>(512, chunk / 4));
    const int64_t stage_cap = chunk + small / 2;
    int rc = c->ensure_host_stage(stage_cap);
    if (rc) return rc;
    HostStage& s = c->hstage;
    rc = c->ensure_pinned_stats(K);
    if (rc) return rc;
    // is the destination pinned (or managed)?  cudaPointerGetAttributes reports plain malloc memory as unregistered
    bool pinned_dst = false;
    {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, U_host) == cudaSuccess)
            pinned_dst = at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
        cudaGetLastError();
    }
    if (!pinned_dst) {
        const size_t need = size_t(stage_cap) * D * 8;
        if (need > s.bounce_cap) {
            for (int i = 0; i < 2; ++i) { if (s.bounce[i]) cudaFreeHost(s.bounce[i]); s.bounce[i] = nullptr; }
            s.bounce_cap = 0;
            for (int i = 0; i < 2; ++i) CK(cudaMallocHost((void**)&s.bounce[i], need));
            s.bounce_cap = need;
        }
        if (!s.copy2) CK(cudaStreamCreateWithFlags(&s.copy2, cudaStreamNonBlocking));
    }
    const int copy_threads = host_copy_threads();
    // Chunk schedule: full chunks first, then halved ones -- only the LAST chunk's D2H copy is exposed (nothing left
    // to overlap it with), so it should be small; chunks below ~600 systems would under-fill the persistent kernels.
    // Every kc the schedule produces is <= stage_cap, the capacity the staging buffers were sized for.
    auto pipeline = [&]() -> int {
        int64_t nchunk = 0;
        for (int64_t k0 = 0; k0 < K; ++nchunk) {
            const int64_t left = K - k0;
            int64_t kc = std::min<int64_t>(chunk, left);
            if (K >= 4096 && left <= chunk + small && left > small) kc = std::max<int64_t>(small, (left + 1) / 2);
            if (left - kc < small / 2) kc = left;                  // no tiny remainder
            kc = std::min<int64_t>(kc, stage_cap);                 // never beyond the staged capacity
            const int slot = int(nchunk & 1);
            if (nchunk >= 2) CK(cudaStreamWaitEvent(s.compute, s.copied[slot], 0));
            CK(cudaMemcpyAsync(s.y[slot], y_host + k0 * nb, size_t(kc) * nb * 8, cudaMemcpyHostToDevice, s.compute));
            int r = c->solve(s.y[slot], kc, s.x[slot], s.it[slot], s.rel[slot], s.compute, nullptr);
            if (r) return r;
            r = c->unpack(s.x[slot], s.u[slot], kc, s.compute);
            if (r) return r;
            CK(cudaEventRecord(s.done[slot], s.compute));
            // pageable destination: slot 0 and slot 1 use different copy streams so that the callback of one chunk and the
            // D2H copy of the next run side by side
            cudaStream_t cs = (!pinned_dst && slot) ? s.copy2 : s.copy;
            CK(cudaStreamWaitEvent(cs, s.done[slot], 0));
            if (pinned_dst) {
                CK(cudaMemcpyAsync(U_host + k0 * D, s.u[slot], size_t(kc) * D * 8, cudaMemcpyDeviceToHost, cs));
            } else {
                CK(cudaMemcpyAsync(s.bounce[slot], s.u[slot], size_t(kc) * D * 8, cudaMemcpyDeviceToHost, cs));
            }
            // per-system statistics go to pinned staging first: a D2H copy into the caller's (usually pageable) arrays would
            // block the host behind the big solution copy and serialise the pipeline
            if (iters_host) CK(cudaMemcpyAsync(s.it_pin + k0, s.it[slot], size_t(kc) * 4, cudaMemcpyDeviceToHost, cs));
            if (relres_host) CK(cudaMemcpyAsync(s.rel_pin + k0, s.rel[slot], size_t(kc) * 8, cudaMemcpyDeviceToHost, cs));
            CK(cudaEventRecord(s.copied[slot], cs));
            if (!pinned_dst) {
                HostCopyTask* t = new HostCopyTask{(const char*)s.bounce[slot], (char*)(U_host + k0 * D), size_t(kc) * D * 8, copy_threads};
                const cudaError_t el = cudaLaunchHostFunc(cs, host_copy_callback, t);
                if (el != cudaSuccess) { delete t; CK(el); }
            }
            k0 += kc;
        }
        return ROMHC_OK;
    };
    rc = pipeline();
    // Drain on EVERY path: queued D2H copies and host callbacks write into the caller's buffer, which the caller may
    // release as soon as this function returns -- also when it returns an error.
    const cudaError_t d0 = cudaStreamSynchronize(s.compute);
    const cudaError_t d1 = cudaStreamSynchronize(s.copy);
    const cudaError_t d2 = s.copy2 ? cudaStreamSynchronize(s.copy2) : cudaSuccess;
    if (rc == ROMHC_OK) { CK(d0); CK(d1); CK(d2); }
    if (rc == ROMHC_OK) {
        if (iters_host) memcpy(iters_host, s.it_pin, size_t(K) * 4);
        if (relres_host) memcpy(relres_host, s.rel_pin, size_t(K) * 8);
    }
    return rc;
}

// ---- pageable host memory <-> padded device layout, pipelined --------------------------------------------------------
// The reference's API hands (K, D) numpy arrays in and out of every call; a plain cudaMemcpy from / to pageable memory
// runs at the speed of one staging thread.  Here the rows move in chunks through two pinned bounce buffers: a
// stream-ordered host callback copies with several threads while the DMA engine moves the previous chunk, and the
// layout kernel (k_pack / k_unpack) runs per chunk on the caller's stream.  Pinned caller memory skips the bounce.
static bool host_ptr_is_pinned(const void* p) {
    cudaPointerAttributes at;
    bool pinned = false;
    if (cudaPointerGetAttributes(&at, p) == cudaSuccess) pinned = at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
    cudaGetLastError();
    return pinned;
}

static int ensure_xfer_stage(Context* c, size_t chunk_bytes, bool need_bounce) {
    HostStage& s = c->hstage;
    if (!s.copy) {
        CK(cudaStreamCreateWithFlags(&s.compute, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&s.copy, cudaStreamNonBlocking));
    }
    if (!s.copy2) CK(cudaStreamCreateWithFlags(&s.copy2, cudaStreamNonBlocking));
    if (chunk_bytes > s.xfer_cap) {
        for (int i = 0; i < 2; ++i) { if (s.xfer_dev[i]) cudaFree(s.xfer_dev[i]); s.xfer_dev[i] = nullptr; }
        s.xfer_cap = 0;
        for (int i = 0; i < 2; ++i) CK(cudaMalloc(&s.xfer_dev[i], chunk_bytes));
        s.xfer_cap = chunk_bytes;
    }
    for (int i = 0; i < 2; ++i) {
        if (!s.xfer_done[i]) CK(cudaEventCreateWithFlags(&s.xfer_done[i], cudaEventDisableTiming));
        if (!s.xfer_ready[i]) CK(cudaEventCreateWithFlags(&s.xfer_ready[i], cudaEventDisableTiming));
    }
    if (need_bounce && chunk_bytes > s.bounce_cap) {
        for (int i = 0; i < 2; ++i) { if (s.bounce[i]) cudaFreeHost(s.bounce[i]); s.bounce[i] = nullptr; }
        s.bounce_cap = 0;
        for (int i = 0; i < 2; ++i) CK(cudaMallocHost((void**)&s.bounce[i], chunk_bytes));
        s.bounce_cap = chunk_bytes;
    }
    return ROMHC_OK;
}

static const size_t XFER_CHUNK_BYTES = size_t(128) << 20;

int romhc_pack_host(romhc_handle h, const double* compact_host, double* padded_dev, int64_t K, void* stream) {
    CHECK_H(h);
    if (K <= 0) return ROMHC_OK;
    if (!compact_host || !padded_dev) { set_error("null argument"); return ROMHC_ERR_ARG; }
    Context* c = H(h);
    const LevelGeo& g = c->levels[0];
    const int64_t D = int64_t(g.R - 1) * (g.C - 1);
    const int64_t rows = std::max<int64_t>(1, std::min<int64_t>(K, int64_t(XFER_CHUNK_BYTES / (size_t(D) * 8))));
    const bool pinned = host_ptr_is_pinned(compact_host);
    int rc = ensure_xfer_stage(c, size_t(rows) * D * 8, !pinned); if (rc) return rc;
    HostStage& s = c->hstage;
    cudaStream_t st = ST(stream);
    const int copy_threads = host_copy_threads();
    auto pipeline = [&]() -> int {
        int64_t i = 0;
        for (int64_t k0 = 0; k0 < K; k0 += rows, ++i) {
            const int64_t kc = std::min(rows, K - k0);
            const int slot = int(i & 1);
            cudaStream_t cs = slot ? s.copy2 : s.copy;
            const size_t bytes = size_t(kc) * D * 8;
            CK(cudaStreamWaitEvent(cs, s.xfer_done[slot], 0));        // the layout kernel that last read this slot has finished
            const double* src = compact_host + k0 * D;
            if (!pinned) {
                HostCopyTask* t = new HostCopyTask{(const char*)src, (char*)s.bounce[slot], bytes, copy_threads};
                const cudaError_t el = cudaLaunchHostFunc(cs, host_copy_callback, t);
                if (el != cudaSuccess) { delete t; CK(el); }
                src = s.bounce[slot];
            }
            CK(cudaMemcpyAsync(s.xfer_dev[slot], src, bytes, cudaMemcpyHostToDevice, cs));
            CK(cudaEventRecord(s.xfer_ready[slot], cs));
            CK(cudaStreamWaitEvent(st, s.xfer_ready[slot], 0));
            const int r = c->pack(s.xfer_dev[slot], padded_dev + k0 * g.Dp, kc, st); if (r) return r;
            CK(cudaEventRecord(s.xfer_done[slot], st));
        }
        return ROMHC_OK;
    };
    rc = pipeline();
    // the caller's host array may be reused as soon as this returns (also on an error): every host-side read has completed
    const cudaError_t d1 = cudaStreamSynchronize(s.copy), d2 = cudaStreamSynchronize(s.copy2);
    if (rc == ROMHC_OK) { CK(d1); CK(d2); }
    return rc;
}

int romhc_unpack_host(romhc_handle h, const double* padded_dev, double* compact_host, int64_t K, void* stream) {
    CHECK_H(h);
    if (K <= 0) return ROMHC_OK;
    if (!compact_host || !padded_dev) { set_error("null argument"); return ROMHC_ERR_ARG; }
    Context* c = H(h);
    const LevelGeo& g = c->levels[0];
    const int64_t D = int64_t(g.R - 1) * (g.C - 1);
    const int64_t rows = std::max<int64_t>(1, std::min<int64_t>(K, int64_t(XFER_CHUNK_BYTES / (size_t(D) * 8))));
    const bool pinned = host_ptr_is_pinned(compact_host);
    int rc = ensure_xfer_stage(c, size_t(rows) * D * 8, !pinned); if (rc) return rc;
    HostStage& s = c->hstage;
    cudaStream_t st = ST(stream);
    const int copy_threads = host_copy_threads();
    auto pipeline = [&]() -> int {
        int64_t i = 0;
        for (int64_t k0 = 0; k0 < K; k0 += rows, ++i) {
            const int64_t kc = std::min(rows, K - k0);
            const int slot = int(i & 1);
            cudaStream_t cs = slot ? s.copy2 : s.copy;
            const size_t bytes = size_t(kc) * D * 8;
            CK(cudaStreamWaitEvent(st, s.xfer_done[slot], 0));        // the D2H copy that last read this slot has finished
            const int r = c->unpack(padded_dev + k0 * g.Dp, s.xfer_dev[slot], kc, st); if (r) return r;
            CK(cudaEventRecord(s.xfer_ready[slot], st));
            CK(cudaStreamWaitEvent(cs, s.xfer_ready[slot], 0));
            double* dst = compact_host + k0 * D;
            CK(cudaMemcpyAsync(pinned ? dst : s.bounce[slot], s.xfer_dev[slot], bytes, cudaMemcpyDeviceToHost, cs));
            CK(cudaEventRecord(s.xfer_done[slot], cs));
            if (!pinned) {
                HostCopyTask* t = new HostCopyTask{(const char*)s.bounce[slot], (char*)dst, bytes, copy_threads};
                const cudaError_t el = cudaLaunchHostFunc(cs, host_copy_callback, t);
                if (el != cudaSuccess) { delete t; CK(el); }
            }
        }
        return ROMHC_OK;
    };
    rc = pipeline();
    // drain on every path: queued copies / callbacks write into the caller's array
    const cudaError_t d1 = cudaStreamSynchronize(s.copy), d2 = cudaStreamSynchronize(s.copy2);
    if (rc == ROMHC_OK) { CK(d1); CK(d2); }
    return rc;
}

int romhc_reduced_galerkin_host(romhc_handle h, const double* y_host, const double* Ahat_host, const double* bhat_host,
                                int n, int64_t K, double* C_host, int* info_host) {
    CHECK_H(h);
    if (K <= 0) return ROMHC_OK;
    Context* c = H(h);
    const int nb = c->nrb * c->ncb;
    // device staging persists in the context (grow only): repeated online batches pay no allocation
    const size_t need[5] = {size_t(K) * nb * 8, size_t(nb) * n * n * 8, size_t(n) * 8, size_t(K) * n * 8, size_t(K) * 4};
    for (int i = 0; i < 5; ++i)
        if (need[i] > c->rg_cap[i]) {
            if (c->rg_buf[i]) cudaFree(c->rg_buf[i]);
            c->rg_buf[i] = nullptr; c->rg_cap[i] = 0;
            CK(cudaMalloc(&c->rg_buf[i], need[i]));
            c->rg_cap[i] = need[i];
        }
    double* y_d = (double*)c->rg_buf[0]; double* A_d = (double*)c->rg_buf[1]; double* b_d = (double*)c->rg_buf[2];
    double* C_d = (double*)c->rg_buf[3]; int* info_d = (int*)c->rg_buf[4];
    cudaStream_t st = 0;
    CK(cudaMemcpyAsync(A_d, Ahat_host, need[1], cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(b_d, bhat_host, need[2], cudaMemcpyHostToDevice, st));
    // chunks: the H2D copy of chunk i+1 and the D2H copy of chunk i-1 ride on the copy engines while chunk i is solved
    // (they only overlap when the caller's buffers are pinned; pageable buffers serialise, still correct)
    const int64_t chunk = std::max<int64_t>(1 << 16, (K + 7) / 8);
    if (!c->hstage.compute) {
        CK(cudaStreamCreateWithFlags(&c->hstage.compute, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&c->hstage.copy, cudaStreamNonBlocking));
    }
    cudaStream_t s_in = c->hstage.copy, s_run = c->hstage.compute;
    { const int rcp = c->ensure_pinned_stats(K); if (rcp) return rcp; }
    CK(cudaStreamSynchronize(st));
    std::vector<cudaEvent_t> ev;
    auto pipeline = [&]() -> int {
        for (int64_t k0 = 0; k0 < K; k0 += chunk) {
            const int64_t kc = std::min<int64_t>(chunk, K - k0);
            cudaEvent_t e;
            CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ev.push_back(e);
            CK(cudaMemcpyAsync(y_d + k0 * nb, y_host + k0 * nb, size_t(kc) * nb * 8, cudaMemcpyHostToDevice, s_in));
            CK(cudaEventRecord(e, s_in));
            CK(cudaStreamWaitEvent(s_run, e, 0));
            const int r = reduced_solve(y_d + k0 * nb, nb, A_d, b_d, 0, n, kc, C_d + k0 * n, info_d + k0, s_run);
            if (r != ROMHC_OK) return r;
            CK(cudaMemcpyAsync(C_host + k0 * n, C_d + k0 * n, size_t(kc) * n * 8, cudaMemcpyDeviceToHost, s_run));
            if (info_host) CK(cudaMemcpyAsync(c->hstage.it_pin + k0, info_d + k0, size_t(kc) * 4, cudaMemcpyDeviceToHost, s_run));
        }
        return ROMHC_OK;
    };
    const int rc = pipeline();
    cudaStreamSynchronize(s_in);
    cudaStreamSynchronize(s_run);
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
    if (rc == ROMHC_OK && info_host) memcpy(info_host, c->hstage.it_pin, size_t(K) * 4);
    if (rc == ROMHC_OK) CK(cudaGetLastError());
    return rc;
}

}  // extern "C"
