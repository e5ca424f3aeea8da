// Internal C++ declarations shared by the .cu translation units (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <array>
#include <atomic>
#include <map>
#include <vector>

#include "common.cuh"

#define ROMHC_OK 0
#define ROMHC_ERR_ARG 1
#define ROMHC_ERR_CUDA 2
#define ROMHC_ERR_NUMERIC 3   // singular / non-SPD system (reference: numpy.linalg.LinAlgError)
#define ROMHC_ERR_NOTCONVERGED 4
#define ROMHC_ROWV_PAD 64      // rows of padding on both sides of the tile kernels' row-class tables

namespace romhc {

extern std::atomic<long long> g_launches;   // every kernel launch of the library
void set_error(const char* fmt, ...);
const char* last_error();

#define CK(call)                                                                                           \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) {                                                                           \
            ::romhc::set_error("%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));      \
            return ROMHC_ERR_CUDA;                                                                         \
        }                                                                                                  \
    } while (0)

struct SolveStats {
    int launched_iterations;
    int chunks;
    int status;
};

struct SolveWorkspace {
    std::vector<double*> r, za, zb;   // per level
    double* p[2];
    double* r_alt;                    // the fused update kernel writes the new residual here, then the two swap
    double* cfac;
    double* wtab;                     // per-system stencil weight tables of the tile kernels (mgtile.cu)
    double *part_pAp, *part_rz;
    int np;
    double *alpha, *alpha2, *beta, *rz, *rz0, *relres;   // alpha2: the step lengths of the odd iterations (deferred x update)
    int *active, *iters;
};

// multigrid tail (levels handled by one CTA per system in shared memory); POD, passed by value
struct TailParams {
    int nlev;                         // number of levels handled by the tail kernel
    LevelGeo geo[ROMHC_MAX_LEVELS];   // geo[0] = first tail level
    int off_r[ROMHC_MAX_LEVELS];      // smem offsets (doubles) of r_l and z_l
    int off_z[ROMHC_MAX_LEVELS];
    int off_fac;                      // smem offset of the Cholesky factor (direct solve)
    int direct;                       // 1: dense Cholesky on the last level, 0: coarse_sweeps of symmetric GS
    int DL, LD;                       // coarsest DOFs, factor pitch
    int coarse_sweeps;
    int nu;                           // smoothing sweeps per level inside the tail
};

enum ProfKind { PROF_PAPPLY = 0, PROF_UPDATE, PROF_DOWN0, PROF_DOWN1, PROF_TAIL, PROF_UP0, PROF_UP1, PROF_BRIDGE, PROF_NKIND };
struct ProfEvent { int kind; cudaEvent_t a, b; };

// persistent staging of the host-buffer entry point (double buffered: compute stream + copy stream)
struct HostStage {
    double *y[2] = {nullptr, nullptr}, *x[2] = {nullptr, nullptr}, *u[2] = {nullptr, nullptr}, *rel[2] = {nullptr, nullptr};
    int* it[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr};
    cudaStream_t compute = nullptr, copy = nullptr;
    int64_t cap = 0;
    int* it_pin = nullptr;        // pinned staging of the per-system statistics (whole batch)
    double* rel_pin = nullptr;
    int64_t pin_cap = 0;
    double* bounce[2] = {nullptr, nullptr};   // pinned bounce buffers for pageable destinations
    size_t bounce_cap = 0;
    cudaStream_t copy2 = nullptr;
    // romhc_pack_host / romhc_unpack_host: device staging of one compact chunk per slot, reuse events
    double* xfer_dev[2] = {nullptr, nullptr};
    size_t xfer_cap = 0;
    cudaEvent_t xfer_done[2] = {nullptr, nullptr}, xfer_ready[2] = {nullptr, nullptr};
};

struct Context {
    int nrb = 0, ncb = 0, N = 0, device = 0;
    // multigrid hierarchy
    std::vector<LevelGeo> levels;
    int tail_level = 0;        // first level handled by the tail kernel (levels.size() if none)
    int smooth_tail_level = 0; // first level smoothed nu_tail times
    int nu_of(int l) const { return l >= smooth_tail_level ? nu_tail : (l == 0 || nu_mid <= 0 ? nu : nu_mid); }
    // Non-nested transfer ("bridge"): when halving the cells per subdomain gets stuck on an odd count whose grid is too
    // large for the dense coarsest solve, level `bridge_level` hands over to a power-of-two hierarchy through bilinear
    // interpolation inside each subdomain (k_bridge_restrict / k_bridge_prolong); -1: every transfer is nested
    int bridge_level = -1;
    bool use_bridge = true;
    bool bridge_res_emitted = false;   // the last going-down kernel of the bridging level left r - A z in zb[l]
    bool fused_coarse(int l) const { return l + 1 < int(levels.size()) && l != bridge_level; }
    int coarse_D = 0, coarse_LD = 1;
    bool coarse_direct = false;
    int coarse_sweeps = 8;
    // red/black Gauss-Seidel sweeps before and after the coarse correction, V(nu, nu): nu on the finest level, nu_mid on
    // the other levels handled by the batched kernels (0: same as nu), nu_tail on the small levels (Dp <= ROMHC_TAIL_MAX_DP)
    // defaults from a sweep on the 256^2 / 10k-system workload (profiles/README.md): V(2,2) / V(3,3) / V(4,4)
    int nu = 2, nu_mid = 3, nu_tail = 4;
    bool use_tile = true;               // register-tiled multigrid kernels (mgtile.cu) where the level fits
    int tile_ty_cap = 64;               // largest strip height of the tile kernels
    bool tile_prefetch = true;          // L2 prefetch of the next CTA's operand rows (non-persistent tile kernels)
    bool tile_persistent = true;        // persistent, TMA-pipelined tile kernels (one CTA per SM)
    bool use_fused = true;              // PCG update fused into the finest level's going-down kernel
    int papply_pers = 1;                // persistent, double-buffered k_pcg_p_apply (fp32 transport of z and p): 0 off, 1 fp64 stencil form (default),
                                        // 2 fp32 combination + edge form on fp32 differences (-5 % kernel time, +0.06 iterations: no net gain)
    int proj_variant = 0;               // reduced operators: 0 edge-difference kernel (n <= 64), 1 stencil apply + split-K DMMA product (any n)
    bool use_sweep = true;              // greedy error sweep: DMMA kernel (sweep.cu); false: the strip kernel k_energy
    // The preconditioned residual z = M r travels from the finest going-up kernel to k_pcg_p_apply as fp32 (half a stream
    // less in each): z only steers the search direction, so rounding it perturbs the preconditioner by 6e-8 relative and
    // leaves x, r, p and every reduction in fp64 (r.z is formed from the ROUNDED z, consistent with what p_apply reads).
    // z32_want: requested for this V-cycle (solves: yes, the precond() test hook: no); z32_out: what the V-cycle produced
    int use_z32 = 3;                    // 0: off, 1: z_B only (up -> p_apply), 2: also z_A (finest down -> up), 3: also p
    bool p_f32 = false;                 // this solve keeps the search direction p as fp32 (k_pcg_p_apply <-> fused update kernel)
    bool z32_want = false, z32_out = false;
    // deferred update of the iterate (TileArgs::xmode): state of the current launch, set by solve_chunk
    int x_mode = 0;
    const float* x_p_prev = nullptr;
    const double* x_alpha_prev = nullptr;
    bool defer_x = false;               // option "defer_x": x is updated every second iteration with two directions at once (bit-identical; -3 % on the
                                        // fused kernel for 36 % fewer bytes on its odd launches: the kernel is latency bound, so off by default)
    bool za_f32 = false;                // the finest going-down kernel stored z_A as fp32 (only the persistent going-up kernel reads that)
    int tile_nsm = 148;
    std::map<std::array<int, 4>, int*> tile_rinfo_cache;   // (level, TY, halo, NR) -> device row-info table
    bool tile_ready = false;
    int tile_maxt_down = 0, tile_maxt_up = 0;
    std::vector<int*> tile_rowv, tile_colv;   // per level: vertex class of every row / column (device)
    int strip_threads = 512;            // threads per strip CTA (256 or 512)
    size_t strip_budget = 113 * 1024;   // shared memory per strip CTA (>= 2 CTAs per SM so TMA loads overlap compute)
    TailParams tail;
    size_t tail_smem = 0;
    // PCG controls
    double rtol = 1e-12;
    int maxit = 1000;
    int min_check_iter = 8, check_every = 2;
    size_t ws_budget_bytes = size_t(48) << 30;
    int host_chunks = 4;       // pipeline depth of the host-buffer entry point for large batches
    // state
    bool kernels_configured = false;
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    void* ws_base = nullptr;
    int64_t ws_K = 0;
    size_t ws_bytes = 0;
    // option "ws_guard" (doubles, 0 = off): every sub-buffer of the solver workspace is followed by a zone of that many
    // doubles filled with the byte 0xA5; romhc_check_guards counts the bytes that changed (an overwrite detector for the
    // pool's closed compute-sanitizer: any kernel writing past the end of r / z / p / ... of the batch lands there)
    int ws_guard = 0;
    std::vector<std::pair<size_t, size_t>> ws_gaps;      // (offset, length) in doubles
    int check_guards(int64_t* n_bad);
    SolveWorkspace ws;
    int* ws_flags = nullptr;
    int* h_flags = nullptr;     // mapped pinned host memory
    int* d_hflags = nullptr;    // its device address
    // per-kernel timing (option "profile")
    bool prof_on = false, prof_window = false;
    std::vector<ProfEvent> prof_events;
    double prof_ms[PROF_NKIND] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long prof_n[PROF_NKIND] = {0, 0, 0, 0, 0, 0, 0, 0};
    void prof_begin(int kind, cudaStream_t st);
    void prof_end(cudaStream_t st);
    void prof_cancel();
    void prof_collect();

    int build_levels();
    int configure_kernels();
    int ensure_scratch(size_t bytes);
    int ensure_solve_ws(int64_t Kc);
    size_t solve_bytes_per_system() const;
    void release();
    HostStage hstage;
    void* rg_buf[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // staging of romhc_reduced_galerkin_host
    size_t rg_cap[5] = {0, 0, 0, 0, 0};
    int ensure_host_stage(int64_t chunk);
    int ensure_pinned_stats(int64_t K);
    void free_host_stage();

    // solver.cu
    int apply(const double* y, const double* u, double* out, int64_t K, cudaStream_t st);
    int energy(const double* y, const double* u, const double* coef, const double* basis, int nbasis, double* out,
               int64_t K, int mode, int take_sqrt, cudaStream_t st);
    // greedy error sweep on the fp64 tensor cores (sweep.cu); -1: configuration does not fit, use energy()
    int error_sweep(const double* U, const double* coef, const double* basis, int n, double* out, int64_t K, cudaStream_t st);
    int pack(const double* compact, double* padded, int64_t K, cudaStream_t st);
    int unpack(const double* padded, double* compact, int64_t K, cudaStream_t st);
    // fuse_p != nullptr: the PCG update x += alpha p, r -= alpha A p is still pending and may be fused into level 0
    // (*fused tells the caller whether it was)
    int vcycle(const double* y, int Kc, cudaStream_t st, const double** z_result, int* np_rz,
               const double* fuse_p = nullptr, double* fuse_x = nullptr, const double* fuse_alpha = nullptr);
    int solve_chunk(const double* y, int Kc, double* x, int* iters_out, double* relres_out, cudaStream_t st,
                    SolveStats* stats, const double* rhs = nullptr);
    int solve(const double* y, int64_t K, double* x, int* iters_out, double* relres_out, cudaStream_t st,
              SolveStats* stats, const double* rhs = nullptr);
    int precond(const double* y, const double* r, double* z, int64_t K, cudaStream_t st);
    int pcg_update(const double* y, int Kc, const double* p, double* x, const double* alpha, cudaStream_t st);
    int bridge_restrict(int l, const double* y, int Kc, cudaStream_t st);
    int bridge_prolong(int l, const double* e, int Kc, cudaStream_t st);

    // mgtile.cu
    int tile_setup();
    int tile_ntab() const;
    int tile_persistent_grid(const void* func, int threads, size_t smem, int64_t items);
    const int* tile_rinfo(int l, int TY, int halo_top, int NR);
    int tile_pf_dist(const void* func, int threads, size_t smem);
    bool tile_level_ok(int l) const;
    bool tile_up_persistent_ok(int l) const;
    bool tile_fused_ok() const;
    int tile_weight_table(const double* y, int Kc, cudaStream_t st);
    int tile_tail(const double* y, int Kc, double* part_rz, cudaStream_t st);
    int tile_update_down(int l, int Kc, const double* p, double* x, const double* alpha, cudaStream_t st);
    int tile_down(int l, const double* y, int Kc, cudaStream_t st);
    int tile_up(int l, const double* y, int Kc, const double* e, double* part_rz, int* ns_out, cudaStream_t st);

    // reduced.cu
    int project_operators(const double* basis, int n, double* Ahat, double* bhat, cudaStream_t st);
    int project_operators_dmma(const double* basis, int n, double* Ahat, cudaStream_t st);
    int evaluate(const double* points, int m, const double* u, int64_t K, double* out, cudaStream_t st);
    int interp_weights(const double* points, int m, int* idx3, double* w3, cudaStream_t st);
};

// reduced.cu (geometry independent)
int reduced_solve(const double* y, int nb, const double* Ahat, const double* rhs, int rhs_per_system, int n,
                  int64_t K, double* C, int* info, cudaStream_t st);
int row_norms(const double* X, int64_t ld, int64_t K, int64_t D, double* out, cudaStream_t st);
int argmax_first(const double* v, int64_t K, int64_t* idx_out, double* val_out, cudaStream_t st);
int dense_spd_solve(const double* y, int nb, const double* A, const double* rhs, int rhs_per_system, int n, int64_t K,
                    double* C, int* info, cudaStream_t st);
int row_dots(const double* X, int64_t ldx, const double* Y, int64_t ldy, int64_t K, int64_t D, double* out, cudaStream_t st);
int tsqr_r(const double* W, int64_t ld, int b, int64_t Dp, double* R_out, cudaStream_t st);
int poly_features(const double* basis, int64_t ld, int n, int64_t D, const int* terms, int nterms, int degree, double* out,
                  int64_t ldo, cudaStream_t st);
int estimator_contract(const double* c, int64_t K, int n, const double* abasis, int nb, int invert, double* out,
                       cudaStream_t st);

// gram.cu
int gemm_nt(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int64_t M, int64_t Nn,
            int64_t Kd, int symmetric, cudaStream_t st);   // C[M,N] = A[M,Kd] * B[N,Kd]^T  (DMMA)
int gemm_nn(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int64_t M, int64_t Nn,
            int64_t Kd, cudaStream_t st);                  // C[M,N] = A[M,Kd] * B[Kd,N]
int gemm_tn(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int64_t M, int64_t Nn,
            int64_t Kd, cudaStream_t st);                  // C[M,N] = A[Kd,M]^T * B[Kd,N]
int column_mean(const double* X, int64_t ld, int64_t K, int64_t D, double* mean, cudaStream_t st);
int center_rows(double* X, int64_t ld, int64_t K, int64_t D, const double* mean, cudaStream_t st);

}  // namespace romhc
