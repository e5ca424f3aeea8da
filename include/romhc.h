/* romhc.h -- C ABI of the B200-native ROMHighContrast hot path (libromhc.so).
 *
 * The reference (agussomacal/ROMHighContrast) is pure Python with no FFI layer; its hot path is the class API of
 * src/lib.  Every entry point below replaces the numpy/scipy call sites named beside it (paths relative to
 * /root/reference/).  The reference-side binding is a ctypes stub, shown in INTEGRATION.md; the drop-in Python
 * classes in romhighcontrast_b200/lib/ are built on exactly these symbols.
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns 0 (ROMHC_OK) or an error code, the message is
 *     available from romhc_last_error() (thread local).
 *   - "*_dev" pointers are CUDA device pointers on the context's device; `stream` is a cudaStream_t passed as
 *     void* (NULL = default stream).  Calls are asynchronous on `stream` unless stated otherwise.
 *   - all arithmetic is IEEE float64, like the reference.
 *   - "padded grid" layout: one field = (R+1) rows x P doubles, P = roundup(C, 8), element (r, c) at r*P + c,
 *     interior DOFs 1 <= r <= R-1, 1 <= c <= C-1, all other slots zero; Dp = (R+1)*P doubles per field.
 *     R = nrb*N, C = ncb*N.  The reference's "compact" layout is u[(r-1)*(C-1) + (c-1)], D = (R-1)*(C-1)
 *     (src/lib/SolutionsManagers.py:153-163).  romhc_pack / romhc_unpack convert.
 *   - parameters y are (K, nrb*ncb) row-major: y[k][p*ncb + q] = a[k][p][q] (SolutionsManagers.py:190-192).
 *   - threading: a context owns its workspaces, staging buffers and streams; calls on ONE context must not overlap
 *     (the reference is single-threaded Python).  Different contexts (one per thread, or one per GPU) are
 *     independent.  Context-free functions (romhc_gemm_*, romhc_reduced_solve, ...) may be called from several host
 *     threads at once: the ones that need device scratch (romhc_gemm_tn, romhc_gemm_nt with symmetric == 2,
 *     romhc_column_mean) keep one buffer per host thread and device; ONE thread must not let such calls overlap on
 *     different streams.
 */
#ifndef ROMHC_H
#define ROMHC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ROMHC_OK 0
#define ROMHC_ERR_ARG 1
#define ROMHC_ERR_CUDA 2
#define ROMHC_ERR_NUMERIC 3       /* singular / non-SPD system: the Python shim raises numpy.linalg.LinAlgError */
#define ROMHC_ERR_NOTCONVERGED 4

typedef struct romhc_context* romhc_handle;

/* ---- library / context ------------------------------------------------------------------------------------ */
int romhc_version(void);
const char* romhc_last_error(void);
/* SolutionsManagerFEM.__init__(blocks_geometry=(nrb, ncb), N)   src/lib/SolutionsManagers.py:146-219 */
int romhc_create(int nrb, int ncb, int N, int device, romhc_handle* out);
int romhc_destroy(romhc_handle h);
/* options: "rtol" (PCG tolerance on sqrt(r.z / r0.z0), default 1e-12), "maxit", "coarse_sweeps",
 * "workspace_gb", "check_every", "min_check_iter", "nu" / "nu_mid" / "nu_tail" (Gauss-Seidel sweeps of the V(nu,nu) cycle on the
 * finest / intermediate / small levels, defaults 2 / 3 / 4), "strip_kb" (shared memory per strip CTA, default 113 = two
 * CTAs per SM), "threads" (256 / 512 per strip CTA), "profile", "bridge" (1: non-nested transfer to a power-of-two
 * hierarchy when N has an odd factor, default; 0: stop coarsening at the odd level), "z32" (3, default: z = M r, the smoothed
 * iterate z_A and the search direction p of the finest level travel between kernels as fp32; 2: z and z_A; 1: only z; 0: all fp64;
 * solves with a caller-supplied right-hand side always use fp64), "papply_pers" (search-direction kernel p = z + beta p, p^T A p:
 * 1, default: persistent double-buffered kernel, fp64 stencil form; 0: one CTA per strip; 2: fp32 combination + edge form on fp32
 * differences), "defer_x" (0, default; 1: the iterate is updated every second iteration with two directions at once --
 * bit-identical solutions, fewer bytes); process-wide A/B switches of the dense helpers:
 * "gram_variant" (1, default: 128 x 64 DMMA tiles, two CTAs per SM; 0: 128 x 128), "tn_variant" (1, default: gemm_tn on the
 * fp64 tensor cores; 0: the plain-FMA kernel, which also serves operands whose rows are not 16-byte aligned) */
int romhc_set_option(romhc_handle h, const char* name, double value);
/* info[0..15] = D, Dp, P, R, C, nlevels, tail_level, coarse_D, coarse_direct, nrb, ncb, N, workspace bytes per system,
 * tail kernel shared memory, bridge_level (-1: none), cells per subdomain below the bridge */
int romhc_get_info(romhc_handle h, int64_t* info16);
/* with option "profile" = 1: accumulated CUDA-event time (ms) and launch count per solver kernel over the first
 * min_check_iter PCG iterations of every solve (all systems active there).  Index: 0 k_pcg_p_apply, 1 k_pcg_update,
 * 2 k_mg_down level 0, 3 k_mg_down levels >= 1, 4 k_mg_tail, 5 k_mg_up level 0, 6 k_mg_up levels >= 1. */
int romhc_get_profile(romhc_handle h, double* ms8, int64_t* n8);
/* with option "ws_guard" = g > 0: every sub-buffer of the solver workspace (residuals, iterates of every level, search
 * directions, tables, scalars) is followed by g doubles of the byte 0xA5; n_bad = bytes of those zones that changed since
 * the workspace was laid out.  Test instrument (the pool's compute-sanitizer is closed): 0 after any sequence of solves. */
int romhc_check_guards(romhc_handle h, int64_t* n_bad);
/* number of CUDA kernels launched by this library in this process (bench.py "gpu_launches") */
int64_t romhc_launch_count(void);

/* ---- memory helpers (so a non-torch host can use the library) ---------------------------------------------- */
int romhc_malloc(void** dev_ptr, size_t bytes);
int romhc_free(void* dev_ptr);
int romhc_malloc_host(void** host_ptr, size_t bytes);   /* pinned */
int romhc_free_host(void* host_ptr);
int romhc_memcpy_h2d(void* dst_dev, const void* src_host, size_t bytes, void* stream);
int romhc_memcpy_d2h(void* dst_host, const void* src_dev, size_t bytes, void* stream);
int romhc_memset(void* dst_dev, int value, size_t bytes, void* stream);
int romhc_stream_sync(void* stream);

/* ---- layout ---------------------------------------------------------------------------------------------------- */
int romhc_pack(romhc_handle h, const double* compact_dev, double* padded_dev, int64_t K, void* stream);
int romhc_unpack(romhc_handle h, const double* padded_dev, double* compact_dev, int64_t K, void* stream);
/* The same conversions with the compact side in HOST memory (the (K, D) numpy arrays of the reference's API,
 * SolutionsManagers.py:56-139): chunked through pinned bounce buffers, several host threads per chunk, DMA and layout
 * kernel overlapped.  Pinned caller memory is copied directly.  The host array is free for reuse on return; the device
 * side is ordered on `stream` (pack) / complete on return (unpack). */
int romhc_pack_host(romhc_handle h, const double* compact_host, double* padded_dev, int64_t K, void* stream);
int romhc_unpack_host(romhc_handle h, const double* padded_dev, double* compact_host, int64_t K, void* stream);

/* ---- K1a: matrix-free stiffness apply  out_k = A(y_k) u_k  (y_dev == NULL: the H10 operator A_1) ----------------
 * replaces np.einsum("pqij,pq->ij", A_preassembled, a) @ u   SolutionsManagers.py:19-23 */
int romhc_apply(romhc_handle h, const double* y_dev, const double* u_pad_dev, double* out_pad_dev, int64_t K,
                void* stream);

/* ---- K2: norms --------------------------------------------------------------------------------------------------
 * romhc_energy_norm: out[k] = sqrt(u_k^T A(y_k) u_k); y_dev == NULL gives H10norm  SolutionsManagers.py:56-58
 * romhc_l2_norm:     out[k] = sqrt(sum u_k^2)                                       SolutionsManagers.py:60-62
 * romhc_error_norm:  out[k] = || sum_j coef[k][j] basis_j - U_k ||_{A_1}, the greedy sweep
 *                    H10norm(approx - solutions2train)                             ReducedBasis.py:129 */
int romhc_energy_norm(romhc_handle h, const double* y_dev, const double* u_pad_dev, int64_t K, double* out_dev,
                      void* stream);
int romhc_l2_norm(romhc_handle h, const double* u_pad_dev, int64_t K, double* out_dev, void* stream);
int romhc_error_norm(romhc_handle h, const double* U_pad_dev, const double* coef_dev, const double* basis_pad_dev,
                     int n, int64_t K, double* out_dev, void* stream);

/* ---- K1: batched snapshot solves  A(y_k) u_k = b,  b = 1/N^2 ------------------------------------------------------
 * replaces generate_solutions / galerkin   SolutionsManagers.py:17-40, 64-68 (all three `method`s)
 * fp64 CG preconditioned by one geometric multigrid V-cycle (V(2,2) / (3,3) / (4,4) by level; options nu, nu_mid, nu_tail); synchronises `stream` before returning.
 * iters_dev / relres_dev (optional): per-system iteration count and final sqrt(r.z / r0.z0).
 * stats4 (optional, host): {iterations launched (sum over chunks), chunks, status bits, workspace bytes}. */
int romhc_solve(romhc_handle h, const double* y_dev, int64_t K, double* x_pad_dev, int* iters_dev, double* relres_dev,
                void* stream, int64_t* stats4);
/* same solver with caller-supplied right-hand sides rhs_pad_dev (K, Dp) (padded layout, zeros outside the interior);
 * y_dev == NULL solves with a == 1, i.e. A_1 u = f: the H10 Riesz representers the reference leaves unimplemented
 * (generate_riesz(norm="h10"), SolutionsManagers.py:78-84) are m such solves with point-evaluation functionals. */
int romhc_solve_rhs(romhc_handle h, const double* y_dev, const double* rhs_pad_dev, int64_t K, double* x_pad_dev,
                    int* iters_dev, double* relres_dev, void* stream, int64_t* stats4);
/* z = M r: one application of the multigrid preconditioner (test hook) */
int romhc_precond(romhc_handle h, const double* y_dev, const double* r_pad_dev, double* z_pad_dev, int64_t K,
                  void* stream);

/* ---- K4: reduced operators  Ahat[q] = Phi A_q Phi^T (nb, n, n), bhat = Phi b (n)  (bhat_dev may be NULL) ----------
 * replaces the nested einsums of generate_fm_solutions / project_solutions  SolutionsManagers.py:93-103,125-133 */
int romhc_project_operators(romhc_handle h, const double* basis_pad_dev, int n, double* Ahat_dev, double* bhat_dev,
                            void* stream);

/* ---- K5: batched reduced solves  (sum_q y[k][q] Ahat[q]) c_k = rhs ---------------------------------------------------
 * replaces map(galerkin, a) on the reduced system  SolutionsManagers.py:104-105, 135-138
 * rhs_dev: (n) shared by all systems (rhs_per_system = 0) or (K, n); info_dev[k] = 1 where the matrix is not SPD.
 * Any n: n <= 24 quad-per-system kernel (DMMA assembly, Cholesky in registers), n <= 64 warp-per-system, larger n a
 * blocked Cholesky a few systems at a time -- which also makes this the device form of galerkin(a, B_total, A_preassembled)
 * on a caller's dense operators (SolutionsManagers.py:17-40) and of the generic SolutionsManager (:43-68). */
int romhc_reduced_solve(const double* y_dev, int nb, const double* Ahat_dev, const double* rhs_dev,
                        int rhs_per_system, int n, int64_t K, double* C_dev, int* info_dev, void* stream);

/* ---- K3 and the dense helpers (row-major, ld in doubles) ---------------------------------------------------------------
 * gemm_nt: C[M,N] = A[M,Kd] B[N,Kd]^T on the fp64 tensor cores; symmetric == 1 (A == B): Gram matrix, only the lower
 *          tiles are computed and mirrored.                   replaces PCA(...).fit  ReducedBasis.py:196
 *          symmetric == 2: plain product, split-K allowed -- for a small C with a long contraction (Krylov-basis
 *          products of the Gram-free POD) the contraction is spread over ~2 CTAs per SM and the partial sums are
 *          reduced in a fixed order (deterministic; needs 16-byte aligned operands, otherwise the plain kernel runs).
 * gemm_nn: C[M,N] = A[M,Kd] B[Kd,N], small Kd (c Phi)         SolutionsManagers.py:106,139
 * gemm_tn: C[M,N] = A[Kd,M]^T B[Kd,N] (V^T Xc; rows of C in blocks of 32)   POD back-projection */
int romhc_gemm_nt(const double* A_dev, int64_t lda, const double* B_dev, int64_t ldb, double* C_dev, int64_t ldc,
                  int64_t M, int64_t N, int64_t Kd, int symmetric, void* stream);
int romhc_gemm_nn(const double* A_dev, int64_t lda, const double* B_dev, int64_t ldb, double* C_dev, int64_t ldc,
                  int64_t M, int64_t N, int64_t Kd, void* stream);
int romhc_gemm_tn(const double* A_dev, int64_t lda, const double* B_dev, int64_t ldb, double* C_dev, int64_t ldc,
                  int64_t M, int64_t N, int64_t Kd, void* stream);
/* R (b, b), upper triangular, of the QR factorisation of W^T for a row block W (b <= 32 rows of length Dp): W W^T = R^T R.
 * Householder TSQR (fixed reduction tree, backward stable): the rank-revealing orthonormalisation of the block-Lanczos
 * POD, where sklearn's PCA (ReducedBasis.py:196) would call LAPACK on the host. */
int romhc_tsqr_r(const double* W_dev, int64_t ld, int b, int64_t Dp, double* R_dev, void* stream);
int romhc_column_mean(const double* X_dev, int64_t ld, int64_t K, int64_t D, double* mean_dev, void* stream);
int romhc_center_rows(double* X_dev, int64_t ld, int64_t K, int64_t D, const double* mean_dev, void* stream);

/* ---- K6: point evaluation, estimators, argmax -----------------------------------------------------------------------------
 * romhc_evaluate: out[k][j] = P1 interpolant of u_k at points[j]      SolutionsManagers.py:221-244
 * romhc_estimator: out[k][q] = sum_b c[b][k] A[b][q] (invert: 1 / sum_b c[b][k] / A[b][q])   Estimators.py:24-37
 * romhc_argmax: first maximum, NaN counts as maximal (np.argmax)                              ReducedBasis.py:129 */
int romhc_evaluate(romhc_handle h, const double* points_dev, int m, const double* u_pad_dev, int64_t K,
                   double* out_dev, void* stream);
/* (padded index or -1, weight) x 3 per point: the sparse rows of generate_riesz(x, norm="l2")  SolutionsManagers.py:70-77 */
int romhc_interp_weights(romhc_handle h, const double* points_dev, int m, int* idx3_dev, double* w3_dev, void* stream);
/* out[k] = ||X[k, :D]||_2 for a generic row-major matrix (SolutionsManager.l2norm is a staticmethod)  :60-62 */
int romhc_row_norms(const double* X_dev, int64_t ld, int64_t K, int64_t D, double* out_dev, void* stream);
/* out[k] = X[k, :D] . Y[k, :D]: with Y = X A^T this is u^T A u, H10norm of the generic dense-operator manager  :56-58 */
int romhc_row_dots(const double* X_dev, int64_t ldx, const double* Y_dev, int64_t ldy, int64_t K, int64_t D, double* out_dev,
                   void* stream);
int romhc_estimator(const double* c_dev, int64_t K, int n, const double* abasis_dev, int nb, int invert,
                    double* out_dev, void* stream);
int romhc_argmax(const double* v_dev, int64_t K, int64_t* idx_dev, double* val_dev, void* stream);
/* out[f][i] = prod_j basis[terms[f][j]][i] (terms (nterms, degree) int32, -1 = unused factor): the monomial features of
 * the basis values at every DOF, i.e. PolynomialFeatures(degree, include_bias=False) on np.array(reduced_basis).T --
 * the predict step of polynomial_state_estimation_fitting_method_least_squares, src/notebooks/InverseProblemPipeline.ipynb cell 52 */
int romhc_poly_features(const double* basis_dev, int64_t ld, int n, int64_t D, const int* terms_dev, int nterms, int degree,
                        double* out_dev, int64_t ldo, void* stream);

/* ---- host-buffer entry points (what a reference-side ctypes binding calls; copies are inside the call) -------------------
 * romhc_generate_solutions_host == SolutionsManager.generate_solutions(a2try): y_host (K, nb) -> U_host (K, D) compact.
 * romhc_reduced_galerkin_host   == the coefficient part of generate_fm_solutions: y_host (K, nb), Ahat_host (nb,n,n),
 *                                  bhat_host (n) -> C_host (K, n), info_host (K) (optional). */
int romhc_generate_solutions_host(romhc_handle h, const double* y_host, int64_t K, double* U_host, int* iters_host,
                                  double* relres_host);
int romhc_reduced_galerkin_host(romhc_handle h, const double* y_host, const double* Ahat_host, const double* bhat_host,
                                int n, int64_t K, double* C_host, int* info_host);

#ifdef __cplusplus
}
#endif
#endif /* ROMHC_H */
