// K3: fp64 tensor-core (DMMA, mma.sync.m8n8k4.f64) GEMMs for the POD path.
//
// Replaces sklearn's PCA(...).fit(snapshots) (/root/reference/src/lib/ReducedBasis.py:196) by the method of
// snapshots: column mean -> centred Gram matrix G = Xc Xc^T (gemm_nt, symmetric: lower tiles only) -> small
// eigensolve on the host side -> back-projection V^T Xc (gemm_tn).  The same gemm_nt kernel with a 32-wide N tile
// serves the tall-skinny products U W^T (projection right-hand sides Phi A_1 U^T, SolutionsManagers.py:113-124).
// tcgen05.mma has no f64 kind, so this is a warp-level-MMA kernel fed by a cp.async multi-stage pipeline.
#include "common.cuh"
#include "romhc_internal.h"

#include <algorithm>

namespace romhc {
int g_gram_variant = 1;     // 1: 128 x 64 tiles, two CTAs per SM (default); 0: 128 x 128 tiles, one CTA per SM

__device__ __forceinline__ void cp_async16(void* smem, const void* g, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem)), "l"(g), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* g, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(smem)), "l"(g), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// C[M,N] (+)= A[M,Kd] * B[N,Kd]^T ; A, B row-major with the contraction index contiguous.
// CTA tile BM x BN, BK = 16, 8 warps laid out WM x WN, STAGES-deep cp.async ring.
// smem row pitch 20 doubles: the 16 lanes of a half-warp (g = 0..3, t = 0..3) read (g*20 + t) mod 16 = distinct banks.
#define GK_BK 16
#define GK_PITCH 20
// SPLITK (small C, long contraction: the Krylov-basis products of pod.krylov_pca): blockIdx.y selects a range of
// `symmetric` (re-used as the chunk length, a multiple of GK_BK) contraction indices and writes its own M x Nn partial
// (C = scratch, ldc = Nn); k_gemm_tn_reduce sums the partials in a fixed order.  The default instantiations compile
// exactly as before.
template <int BM, int BN, int WM, int WN, int STAGES, int VEC, int MINB = 1, bool SPLITK = false>
__global__ void __launch_bounds__(256, MINB)
k_gemm_nt(const double* __restrict__ A, int64_t lda, const double* __restrict__ B, int64_t ldb, double* __restrict__ C,
          int64_t ldc, int64_t M, int64_t Nn, int64_t Kd, int symmetric, int ntile_n) {
    constexpr int WTM = BM / WM, WTN = BN / WN;       // warp tile
    constexpr int MT = WTM / 8, NT = WTN / 8;         // 8x8 mma tiles per warp
    if constexpr (SPLITK) {
        const int64_t kc = symmetric, k0 = int64_t(blockIdx.y) * kc;
        A += k0; B += k0;
        Kd = (Kd - k0 < kc) ? Kd - k0 : kc;
        C += int64_t(blockIdx.y) * M * ldc;
        symmetric = 0;
    }
    extern __shared__ __align__(16) double smg[];
    double* sA = smg;
    double* sB = smg + size_t(STAGES) * BM * GK_PITCH;
    int tm, tn;
    if (symmetric) {   // linear index over the lower triangle of tiles: tile row i holds (BM / BN) (i + 1) tiles
        constexpr int RT = BM / BN;
        const int t = blockIdx.x;
        int i = int((sqrt(8.0 * t / RT + 1.0) - 1.0) * 0.5);
        while (RT * i * (i + 1) / 2 > t) --i;
        while (RT * (i + 1) * (i + 2) / 2 <= t) ++i;
        tm = i; tn = t - RT * i * (i + 1) / 2;
    } else {
        tm = blockIdx.x / ntile_n; tn = blockIdx.x % ntile_n;
    }
    const int64_t m0 = int64_t(tm) * BM, n0 = int64_t(tn) * BN;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp / WN, wn = warp % WN;
    const int g = lane >> 2, t4 = lane & 3;

    double acc[MT][NT][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int nk = int((Kd + GK_BK - 1) / GK_BK);
    auto load_stage = [&](int stage, int kb) {
        const int64_t k0 = int64_t(kb) * GK_BK;
        double* dA = sA + size_t(stage) * BM * GK_PITCH;
        double* dB = sB + size_t(stage) * BN * GK_PITCH;
        constexpr int EPV = VEC / 8;                  // doubles per cp.async
        constexpr int VPR = GK_BK / EPV;              // vectors per tile row
        for (int v = tid; v < BM * VPR; v += 256) {
            const int r = v / VPR, c = (v % VPR) * EPV;
            const int64_t gr = m0 + r, gc = k0 + c;
            int bytes = 0;
            if (gr < M && gc < Kd) bytes = ((Kd - gc < EPV) ? int(Kd - gc) : EPV) * 8;
            const double* src = A + (gr < M ? gr : 0) * lda + (gc < Kd ? gc : 0);
            if (VEC == 16) cp_async16(dA + r * GK_PITCH + c, src, bytes); else cp_async8(dA + r * GK_PITCH + c, src, bytes);
        }
        for (int v = tid; v < BN * VPR; v += 256) {
            const int r = v / VPR, c = (v % VPR) * EPV;
            const int64_t gr = n0 + r, gc = k0 + c;
            int bytes = 0;
            if (gr < Nn && gc < Kd) bytes = ((Kd - gc < EPV) ? int(Kd - gc) : EPV) * 8;
            const double* src = B + (gr < Nn ? gr : 0) * ldb + (gc < Kd ? gc : 0);
            if (VEC == 16) cp_async16(dB + r * GK_PITCH + c, src, bytes); else cp_async8(dB + r * GK_PITCH + c, src, bytes);
        }
    };
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) load_stage(s, s);
        cp_async_commit();
    }
    for (int kb = 0; kb < nk; ++kb) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const int nxt = kb + STAGES - 1;
        if (nxt < nk) load_stage(nxt % STAGES, nxt);
        cp_async_commit();
        const double* cA = sA + size_t(kb % STAGES) * BM * GK_PITCH + (wm * WTM + g) * GK_PITCH + t4;
        const double* cB = sB + size_t(kb % STAGES) * BN * GK_PITCH + (wn * WTN + g) * GK_PITCH + t4;
#pragma unroll
        for (int kk = 0; kk < GK_BK; kk += 4) {
            double a[MT], b[NT];
#pragma unroll
            for (int i = 0; i < MT; ++i) a[i] = cA[i * 8 * GK_PITCH + kk];
#pragma unroll
            for (int j = 0; j < NT; ++j) b[j] = cB[j * 8 * GK_PITCH + kk];
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();
    // epilogue: c0 = C[g][2 t4], c1 = C[g][2 t4 + 1] of every 8x8 tile
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        const int64_t row = m0 + wm * WTM + i * 8 + g;
        if (row >= M) continue;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int64_t col = n0 + wn * WTN + j * 8 + 2 * t4;
            if (col < Nn) C[row * ldc + col] = acc[i][j][0];
            if (col + 1 < Nn) C[row * ldc + col + 1] = acc[i][j][1];
        }
    }
}

__global__ void k_mirror_lower(double* __restrict__ C, int64_t ldc, int64_t M) {
    const int64_t j = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;   // column
    const int64_t i = blockIdx.y;                                        // row
    if (j < M && j > i) C[i * ldc + j] = C[j * ldc + i];
}

template <int BM, int BN, int WM, int WN, int STAGES, int VEC, int MINB = 1>
static int launch_gemm_nt(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc,
                          int64_t M, int64_t Nn, int64_t Kd, int symmetric, cudaStream_t st) {
    auto kern = k_gemm_nt<BM, BN, WM, WN, STAGES, VEC, MINB>;
    const size_t sm = size_t(STAGES) * (BM + BN) * GK_PITCH * 8;
    // per call: the attribute is per device and a process may drive several
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm)));
    const int64_t tmn = (M + BM - 1) / BM, tnn = (Nn + BN - 1) / BN;
    const int64_t ntiles = symmetric ? (BM / BN) * tmn * (tmn + 1) / 2 : tmn * tnn;
    ++g_launches; kern<<<(unsigned)ntiles, 256, sm, st>>>(A, lda, B, ldb, C, ldc, M, Nn, Kd, symmetric, int(tnn));
    CK(cudaGetLastError());
    return ROMHC_OK;
}

__global__ void k_gemm_tn_reduce(const double* __restrict__ part, int nchunks, int Mm, int64_t Nn, double* __restrict__ C,
                                 int64_t ldc);
// Device scratch of the context-free dense helpers (split-K partials, gemm_tn partials, column sums): one buffer per
// HOST THREAD and device, grown on demand (cudaFree synchronises, so a buffer still in use is never released early).
// Two threads therefore never share partials; one thread must not overlap helper calls on different streams.
struct HelperScratch { void* p = nullptr; size_t bytes = 0; int dev = -1; };
static thread_local HelperScratch g_scr;
#define g_tn_scratch (g_scr.p)
static int scratch_reserve(size_t need) {
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (dev != g_scr.dev || need > g_scr.bytes) {
        if (g_scr.p) cudaFree(g_scr.p);
        g_scr = HelperScratch();
        CK(cudaMalloc(&g_scr.p, need));
        g_scr.bytes = need;
        g_scr.dev = dev;
    }
    return ROMHC_OK;
}

// Split-K form for few output tiles and a long contraction (M, Nn <= a few thousand, Kd ~ 10^5): enough K ranges to
// put ~2 CTAs on every SM, partials in scratch, deterministic reduction.
template <int BM, int BN, int WM, int WN, int STAGES, int MINB>
static int launch_gemm_nt_splitk(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc,
                                 int64_t M, int64_t Nn, int64_t Kd, int nsplit, cudaStream_t st) {
    auto kern = k_gemm_nt<BM, BN, WM, WN, STAGES, 16, MINB, true>;
    const size_t sm = size_t(STAGES) * (BM + BN) * GK_PITCH * 8;
    // per call: the attribute is per device and a process may drive several
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm)));
    const int64_t kc = ((Kd + nsplit - 1) / nsplit + GK_BK - 1) / GK_BK * GK_BK;
    nsplit = int((Kd + kc - 1) / kc);
    if (int rc = scratch_reserve(size_t(nsplit) * M * Nn * 8)) return rc;
    const int64_t tmn = (M + BM - 1) / BM, tnn = (Nn + BN - 1) / BN;
    ++g_launches; kern<<<dim3((unsigned)(tmn * tnn), (unsigned)nsplit), 256, sm, st>>>(A, lda, B, ldb, (double*)g_tn_scratch, Nn, M, Nn,
                                                                                   Kd, int(kc), int(tnn));
    CK(cudaGetLastError());
    ++g_launches; k_gemm_tn_reduce<<<dim3((unsigned)((Nn + 255) / 256), (unsigned)M), 256, 0, st>>>((double*)g_tn_scratch, nsplit, int(M),
                                                                                       Nn, C, ldc);
    CK(cudaGetLastError());
    return ROMHC_OK;
}

int gemm_nt(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int64_t M, int64_t Nn,
            int64_t Kd, int symmetric, cudaStream_t st) {
    if (M <= 0 || Nn <= 0) return ROMHC_OK;
    const bool al16 = (lda % 2 == 0) && (ldb % 2 == 0) && ((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0);
    if (symmetric == 2) {          // caller allows split-K (never symmetric); fall through when it would not help
        symmetric = 0;
        const bool skinny = Nn <= 32;
        const int64_t tiles = ((M + 127) / 128) * (skinny ? 1 : (Nn + 63) / 64);
        // number of contraction ranges: fill 1..4 waves of 2 CTAs per SM as completely as possible, >= 2048 indices each
        // (>= 256 for short contractions: the Krylov-basis products of the block Lanczos on a K x K Gram matrix have
        // Kd = K ~ 10^4 and a handful of output tiles -- four ranges left them at 0.57 ms per product)
        const int64_t cap = std::min<int64_t>(Kd / (Kd >= 32768 ? 2048 : 256), 4096), slots = 2 * 148;
        int64_t nsplit = 1;
        double best = 0.0;
        for (int64_t w = 1; w <= 4; ++w) {
            const int64_t ns = std::max<int64_t>(1, std::min<int64_t>(slots * w / tiles, cap));
            const int64_t waves = (tiles * ns + slots - 1) / slots;
            const double eff = double(tiles * ns) / double(slots * waves);
            if (eff > best + 0.02) { best = eff; nsplit = ns; }
        }
        if (al16 && nsplit >= 2 && M <= 65535)
            return skinny ? launch_gemm_nt_splitk<128, 32, 8, 1, 4, 1>(A, lda, B, ldb, C, ldc, M, Nn, Kd, int(nsplit), st)
                          : launch_gemm_nt_splitk<128, 64, 4, 2, 3, 2>(A, lda, B, ldb, C, ldc, M, Nn, Kd, int(nsplit), st);
    }
    if (symmetric && (M != Nn)) { set_error("gemm_nt: symmetric needs M == N"); return ROMHC_ERR_ARG; }
    int rc;
    if (Nn <= 32 && !symmetric) {
        rc = al16 ? launch_gemm_nt<128, 32, 8, 1, 4, 16>(A, lda, B, ldb, C, ldc, M, Nn, Kd, 0, st)
                  : launch_gemm_nt<128, 32, 8, 1, 4, 8>(A, lda, B, ldb, C, ldc, M, Nn, Kd, 0, st);
    } else if (g_gram_variant == 1 && al16) {
        // 128 x 64 tiles, two CTAs per SM: 4 warps per scheduler hide the DMMA issue latency that capped the 128 x 128
        // single-CTA version at 78 % tensor-pipe activity (K = 10 000: 235 -> 211 ms, 27.6 -> 30.8 TFLOP/s, bit identical)
        rc = launch_gemm_nt<128, 64, 4, 2, 3, 16, 2>(A, lda, B, ldb, C, ldc, M, Nn, Kd, symmetric, st);
    } else {
        rc = al16 ? launch_gemm_nt<128, 128, 2, 4, 4, 16>(A, lda, B, ldb, C, ldc, M, Nn, Kd, symmetric, st)
                  : launch_gemm_nt<128, 128, 2, 4, 4, 8>(A, lda, B, ldb, C, ldc, M, Nn, Kd, symmetric, st);
    }
    if (rc) return rc;
    if (symmetric) {
        ++g_launches; k_mirror_lower<<<dim3((unsigned)((M + 255) / 256), (unsigned)M), 256, 0, st>>>(C, ldc, M);
        CK(cudaGetLastError());
    }
    return ROMHC_OK;
}

// ---- C[M,N] = A[M,Kd] * B[Kd,N], small Kd (reconstruction c Phi, SolutionsManagers.py:106,139) ------------------------
// thread = one column pair of C for 8 rows; A rows broadcast from smem; B streamed through registers.
#define NN_ROWS 8
__global__ void __launch_bounds__(256)
k_gemm_nn(const double* __restrict__ A, int64_t lda, const double* __restrict__ B, int64_t ldb, double* __restrict__ C,
          int64_t ldc, int64_t M, int64_t Nn, int Kd, int accumulate) {
    extern __shared__ __align__(16) double sAr[];   // NN_ROWS x Kd
    const int64_t m0 = int64_t(blockIdx.y) * NN_ROWS;
    for (int i = threadIdx.x; i < NN_ROWS * Kd; i += blockDim.x) {
        const int r = i / Kd, c = i % Kd;
        sAr[i] = (m0 + r < M) ? A[(m0 + r) * lda + c] : 0.0;
    }
    __syncthreads();
    const int64_t col = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (col >= Nn) return;
    double acc[NN_ROWS];
#pragma unroll
    for (int r = 0; r < NN_ROWS; ++r) acc[r] = (accumulate && m0 + r < M) ? C[(m0 + r) * ldc + col] : 0.0;
    for (int j = 0; j < Kd; ++j) {
        const double b = B[int64_t(j) * ldb + col];
#pragma unroll
        for (int r = 0; r < NN_ROWS; ++r) acc[r] = fma(sAr[r * Kd + j], b, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < NN_ROWS; ++r)
        if (m0 + r < M) C[(m0 + r) * ldc + col] = acc[r];
}
int gemm_nn(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int64_t M, int64_t Nn,
            int64_t Kd, cudaStream_t st) {
    if (M <= 0 || Nn <= 0) return ROMHC_OK;
    const int64_t rows_blocks = (M + NN_ROWS - 1) / NN_ROWS;
    const int KC = 2048;                            // contraction chunk held in shared memory; longer ones accumulate into C
    CK(cudaFuncSetAttribute(k_gemm_nn, cudaFuncAttributeMaxDynamicSharedMemorySize, NN_ROWS * KC * 8));
    if (Kd <= 0) {
        for (int64_t r = 0; r < M; ++r) CK(cudaMemsetAsync(C + r * ldc, 0, size_t(Nn) * 8, st));
        return ROMHC_OK;
    }
    for (int64_t k0 = 0; k0 < Kd; k0 += KC) {
        const int kc = int(std::min<int64_t>(KC, Kd - k0));
        for (int64_t b0 = 0; b0 < rows_blocks; b0 += 65535) {
            const int nb = int(std::min<int64_t>(65535, rows_blocks - b0));
            ++g_launches; k_gemm_nn<<<dim3((unsigned)((Nn + 255) / 256), nb), 256, size_t(NN_ROWS) * kc * 8, st>>>(
                A + b0 * NN_ROWS * lda + k0, lda, B + k0 * ldb, ldb, C + b0 * NN_ROWS * ldc, ldc, M - b0 * NN_ROWS, Nn, kc,
                k0 > 0 ? 1 : 0);
        }
    }
    CK(cudaGetLastError());
    return ROMHC_OK;
}

// ---- C[M,N] = A[Kd,M]^T * B[Kd,N], small M (POD back-projection V^T Xc); split over Kd, two passes ---------------------
#define TN_MAXM 32
#define TN_CHUNK 128
__global__ void __launch_bounds__(256)
k_gemm_tn_partial(const double* __restrict__ A, int64_t lda, const double* __restrict__ B, int64_t ldb,
                  double* __restrict__ part, int Mm, int64_t Nn, int64_t Kd) {
    __shared__ double sV[TN_CHUNK * TN_MAXM];
    const int64_t k0 = int64_t(blockIdx.y) * TN_CHUNK;
    const int kn = (Kd - k0 < TN_CHUNK) ? int(Kd - k0) : TN_CHUNK;
    for (int i = threadIdx.x; i < kn * Mm; i += blockDim.x) sV[i] = A[(k0 + i / Mm) * lda + i % Mm];
    __syncthreads();
    const int64_t col = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (col >= Nn) return;
    double acc[TN_MAXM];
#pragma unroll
    for (int m = 0; m < TN_MAXM; ++m) acc[m] = 0.0;
    for (int k = 0; k < kn; ++k) {
        const double b = B[(k0 + k) * ldb + col];
#pragma unroll
        for (int m = 0; m < TN_MAXM; ++m)
            if (m < Mm) acc[m] = fma(sV[k * Mm + m], b, acc[m]);
    }
    double* dst = part + (int64_t(blockIdx.y) * Mm) * Nn + col;
#pragma unroll
    for (int m = 0; m < TN_MAXM; ++m)
        if (m < Mm) dst[int64_t(m) * Nn] = acc[m];
}
__global__ void k_gemm_tn_reduce(const double* __restrict__ part, int nchunks, int Mm, int64_t Nn, double* __restrict__ C,
                                 int64_t ldc) {
    const int64_t col = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    const int m = blockIdx.y;
    if (col >= Nn) return;
    double s = 0.0;
    for (int c = 0; c < nchunks; ++c) s += part[(int64_t(c) * Mm + m) * Nn + col];
    C[int64_t(m) * ldc + col] = s;
}
// ---- the same product on the fp64 tensor cores -------------------------------------------------------------------------
// C[m][n] = sum_k A[k][m] B[k][n]: the contraction runs over ROWS of both operands (snapshots), so a stage is TNM_BK
// rows of B (each a contiguous 2 KB run of 256 columns: long coalesced reads, unlike the 128-byte row pieces of the NT
// form) plus the matching TNM_BK x 32 block of A, staged by cp.async in a 3-deep ring.  8 warps, each owns all MT m-tiles
// of 32 columns (4 n-tiles): 4 MT DMMA per (MT + 4) shared loads per k-step of 4.  Shared row pitches = 4 mod 16 doubles:
// the 16 lanes of a half-warp (k = t4, index = g) hit 16 distinct 8-byte banks.  blockIdx.y = range of the contraction;
// partials in scratch, summed in a fixed order by k_gemm_tn_reduce.  Two CTAs per SM (2 x 111 KB of shared memory).
#define TNM_BK 16
#define TNM_BN 256
#define TNM_PA 36
#define TNM_PB 260
#define TNM_STAGES 3
template <int MT>
__global__ void __launch_bounds__(256, 2)
k_gemm_tn_mma(const double* __restrict__ A, int64_t lda, const double* __restrict__ B, int64_t ldb, double* __restrict__ C,
              int64_t ldc, int Mm, int64_t Nn, int64_t Kd, int64_t kc) {
    extern __shared__ __align__(16) double smt[];
    double* sA = smt;                                              // STAGES x BK x PA
    double* sB = smt + size_t(TNM_STAGES) * TNM_BK * TNM_PA;       // STAGES x BK x PB
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int64_t n0 = int64_t(blockIdx.x) * TNM_BN;
    const int64_t kbeg = int64_t(blockIdx.y) * kc;
    const int64_t kend = (kbeg + kc < Kd) ? kbeg + kc : Kd;
    C += int64_t(blockIdx.y) * Mm * ldc;
    double acc[MT][4][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const int nk = int((kend - kbeg + TNM_BK - 1) / TNM_BK);
    auto load_stage = [&](int stage, int kb) {
        const int64_t k0 = kbeg + int64_t(kb) * TNM_BK;
        double* dA = sA + size_t(stage) * TNM_BK * TNM_PA;
        double* dB = sB + size_t(stage) * TNM_BK * TNM_PB;
        for (int v = tid; v < TNM_BK * (TNM_BN / 2); v += 256) {   // 16-byte vectors of the B rows
            const int r = v / (TNM_BN / 2), c = (v % (TNM_BN / 2)) * 2;
            const int64_t gr = k0 + r, gc = n0 + c;
            int bytes = 0;
            if (gr < kend && gc < Nn) bytes = (Nn - gc < 2) ? 8 : 16;
            const double* src = B + (gr < kend ? gr : kbeg) * ldb + (gc < Nn ? gc : 0);
            cp_async16(dB + r * TNM_PB + c, src, bytes);
        }
        for (int v = tid; v < TNM_BK * MT * 8; v += 256) {          // A block, 8 bytes at a time (any lda)
            const int r = v / (MT * 8), c = v % (MT * 8);
            const int64_t gr = k0 + r;
            const bool ok = gr < kend && c < Mm;
            const double* src = A + (ok ? gr * lda + c : kbeg * lda);
            cp_async8(dA + r * TNM_PA + c, src, ok ? 8 : 0);
        }
    };
#pragma unroll
    for (int s = 0; s < TNM_STAGES - 1; ++s) {
        if (s < nk) load_stage(s, s);
        cp_async_commit();
    }
    for (int kb = 0; kb < nk; ++kb) {
        cp_async_wait<TNM_STAGES - 2>();
        __syncthreads();
        const int nxt = kb + TNM_STAGES - 1;
        if (nxt < nk) load_stage(nxt % TNM_STAGES, nxt);
        cp_async_commit();
        const double* cA = sA + size_t(kb % TNM_STAGES) * TNM_BK * TNM_PA + t4 * TNM_PA + g;
        const double* cB = sB + size_t(kb % TNM_STAGES) * TNM_BK * TNM_PB + t4 * TNM_PB + warp * 32 + g;
#pragma unroll
        for (int kk = 0; kk < TNM_BK; kk += 4) {
            double a[MT], b[4];
#pragma unroll
            for (int i = 0; i < MT; ++i) a[i] = cA[kk * TNM_PA + i * 8];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = cB[kk * TNM_PB + j * 8];
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        const int row = i * 8 + g;
        if (row >= Mm) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t col = n0 + warp * 32 + j * 8 + 2 * t4;
            if (col < Nn) C[int64_t(row) * ldc + col] = acc[i][j][0];
            if (col + 1 < Nn) C[int64_t(row) * ldc + col + 1] = acc[i][j][1];
        }
    }
}

template <int MT>
static int launch_gemm_tn_mma(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int64_t M,
                              int64_t Nn, int64_t Kd, cudaStream_t st) {
    auto kern = k_gemm_tn_mma<MT>;
    const size_t sm = size_t(TNM_STAGES) * TNM_BK * (TNM_PA + TNM_PB) * 8;
    // per call: the attribute is per device and a process may drive several
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm)));
    const int64_t ntiles = (Nn + TNM_BN - 1) / TNM_BN;
    int64_t nsplit = std::max<int64_t>(1, std::min<int64_t>((8 * 2 * 148 + ntiles - 1) / ntiles, Kd / 512));
    nsplit = std::min<int64_t>(nsplit, 65535);
    const int64_t kc = ((Kd + nsplit - 1) / nsplit + TNM_BK - 1) / TNM_BK * TNM_BK;
    nsplit = (Kd + kc - 1) / kc;
    if (nsplit == 1) {
        ++g_launches; kern<<<dim3((unsigned)ntiles, 1), 256, sm, st>>>(A, lda, B, ldb, C, ldc, int(M), Nn, Kd, kc);
        CK(cudaGetLastError());
        return ROMHC_OK;
    }
    if (int rc = scratch_reserve(size_t(nsplit) * M * Nn * 8)) return rc;
    ++g_launches; kern<<<dim3((unsigned)ntiles, (unsigned)nsplit), 256, sm, st>>>(A, lda, B, ldb, (double*)g_tn_scratch, Nn, int(M), Nn,
                                                                               Kd, kc);
    CK(cudaGetLastError());
    ++g_launches; k_gemm_tn_reduce<<<dim3((unsigned)((Nn + 255) / 256), (unsigned)M), 256, 0, st>>>((double*)g_tn_scratch, int(nsplit),
                                                                                       int(M), Nn, C, ldc);
    CK(cudaGetLastError());
    return ROMHC_OK;
}

int g_tn_variant = 1;       // 1: DMMA kernel (default, needs 16-byte aligned B rows); 0: the plain-FMA kernel above
int gemm_tn(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int64_t M, int64_t Nn,
            int64_t Kd, cudaStream_t st) {
    if (M <= 0 || Nn <= 0) return ROMHC_OK;
    if (M > TN_MAXM) {                       // row blocks of 32: the kernels keep all their output rows in registers
        for (int64_t m0 = 0; m0 < M; m0 += TN_MAXM) {
            const int rc = gemm_tn(A + m0, lda, B, ldb, C + m0 * ldc, ldc, std::min<int64_t>(TN_MAXM, M - m0), Nn, Kd, st);
            if (rc) return rc;
        }
        return ROMHC_OK;
    }
    if (g_tn_variant == 1 && Kd > 0 && (ldb % 2 == 0) && ((uintptr_t)B % 16 == 0)) {
        switch ((M + 7) / 8) {
            case 1: return launch_gemm_tn_mma<1>(A, lda, B, ldb, C, ldc, M, Nn, Kd, st);
            case 2: return launch_gemm_tn_mma<2>(A, lda, B, ldb, C, ldc, M, Nn, Kd, st);
            case 3: return launch_gemm_tn_mma<3>(A, lda, B, ldb, C, ldc, M, Nn, Kd, st);
            default: return launch_gemm_tn_mma<4>(A, lda, B, ldb, C, ldc, M, Nn, Kd, st);
        }
    }
    const int nch = int((Kd + TN_CHUNK - 1) / TN_CHUNK);
    if (int rc = scratch_reserve(size_t(nch) * M * Nn * 8)) return rc;
    ++g_launches; k_gemm_tn_partial<<<dim3((unsigned)((Nn + 255) / 256), nch), 256, 0, st>>>(A, lda, B, ldb, (double*)g_tn_scratch,
                                                                                 int(M), Nn, Kd);
    ++g_launches; k_gemm_tn_reduce<<<dim3((unsigned)((Nn + 255) / 256), (unsigned)M), 256, 0, st>>>((double*)g_tn_scratch, nch, int(M),
                                                                                       Nn, C, ldc);
    CK(cudaGetLastError());
    return ROMHC_OK;
}

// ---- column mean over the K rows and in-place centring (PCA is mean-centred, ReducedBasis.py:196) ------------------------
__global__ void __launch_bounds__(256)
k_colsum_partial(const double* __restrict__ X, int64_t ld, int64_t K, int64_t D, double* __restrict__ part, int rows_per) {
    const int64_t col = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (col >= D) return;
    const int64_t r0 = int64_t(blockIdx.y) * rows_per, r1 = (r0 + rows_per < K) ? r0 + rows_per : K;
    double s = 0.0;
    for (int64_t r = r0; r < r1; ++r) s += X[r * ld + col];
    part[int64_t(blockIdx.y) * D + col] = s;
}
__global__ void k_colsum_final(const double* __restrict__ part, int nparts, int64_t D, double inv, double* __restrict__ mean) {
    const int64_t col = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (col >= D) return;
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += part[int64_t(p) * D + col];
    mean[col] = s * inv;
}
__global__ void k_center(double* __restrict__ X, int64_t ld, int64_t K, int64_t D, const double* __restrict__ mean) {
    const int64_t col = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    const int64_t r = blockIdx.y;
    if (col < D) X[r * ld + col] -= mean[col];
}
int column_mean(const double* X, int64_t ld, int64_t K, int64_t D, double* mean, cudaStream_t st) {
    const int rows_per = 128;
    const int np = int((K + rows_per - 1) / rows_per);
    if (int rc = scratch_reserve(size_t(np) * D * 8)) return rc;
    ++g_launches; k_colsum_partial<<<dim3((unsigned)((D + 255) / 256), np), 256, 0, st>>>(X, ld, K, D, (double*)g_tn_scratch, rows_per);
    ++g_launches; k_colsum_final<<<(unsigned)((D + 255) / 256), 256, 0, st>>>((double*)g_tn_scratch, np, D, 1.0 / double(K), mean);
    CK(cudaGetLastError());
    return ROMHC_OK;
}
int center_rows(double* X, int64_t ld, int64_t K, int64_t D, const double* mean, cudaStream_t st) {
    for (int64_t r0 = 0; r0 < K; r0 += 65535) {
        const int nr = int(std::min<int64_t>(65535, K - r0));
        ++g_launches; k_center<<<dim3((unsigned)((D + 255) / 256), nr), 256, 0, st>>>(X + r0 * ld, ld, nr, D, mean);
    }
    CK(cudaGetLastError());
    return ROMHC_OK;
}

}  // namespace romhc
