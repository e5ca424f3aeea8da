// K4/K5/K6: reduced operators, batched small SPD solves, point evaluation, estimators, argmax.
//
// Replaces (batched, on device):
//   generate_fm_solutions()  /root/reference/src/lib/SolutionsManagers.py:88-106  (Phi A_pq Phi^T, Phi b, K n x n solves)
//   project_solutions()      :108-139
//   evaluate_solutions()     :221-244
//   EstimatorLinear/Inv      /root/reference/src/lib/Estimators.py:24-37
//   np.argmax                /root/reference/src/lib/ReducedBasis.py:129
#include "common.cuh"
#include "romhc_internal.h"

#include <algorithm>
#include <type_traits>
#include <vector>

namespace romhc {

// ======================================================================================================
// K4: Ahat_q = Phi A_q Phi^T summed over the mesh edges of block q:
//   phi_i^T A_q phi_j = 1/2 * sum_{cells in q} sum_{4 edges e of the cell} (dphi_i)(e) (dphi_j)(e)
// (every cell gives half of each of its edges' weight -- the closed form of SolutionsManagers.py:187-215).
// grid = (N cell rows of the block, nb blocks); partial n x n matrices are reduced by k_reduce_ahat.
// ======================================================================================================
#define PROJ_EB 64   // edges per batch (16 cells)
#define PROJ_MAXN 64  // n * n <= 16 entries per thread
__global__ void __launch_bounds__(256)
k_project_partial(LevelGeo g, const double* __restrict__ basis, int n, double* __restrict__ part) {
    extern __shared__ __align__(16) double G[];   // PROJ_EB x (n | 1)
    const int ldg = n | 1;
    const int q = blockIdx.y, crl = blockIdx.x;
    const int bp = q / g.ncb, bq = q % g.ncb;
    const int cr = bp * g.N + crl;                 // cell row
    const int tid = threadIdx.x, nt = blockDim.x;
    const int nn = n * n;
    const int nslot = (nn + 255) >> 8;            // entries tid, tid+256, ... (warp-uniform count, <= 16)
    double acc[PROJ_MAXN * PROJ_MAXN / 256];
#pragma unroll
    for (int s = 0; s < PROJ_MAXN * PROJ_MAXN / 256; ++s) acc[s] = 0.0;
    for (int c0 = 0; c0 < g.N; c0 += PROJ_EB / 4) {
        const int ncell = min(PROJ_EB / 4, g.N - c0);
        __syncthreads();
        // G[e][i] = phi_i(v1) - phi_i(v2) for the 4 edges of each cell
        for (int idx = tid; idx < ncell * 4 * n; idx += nt) {
            const int i = idx % n, e = idx / n;
            const int cell = e >> 2, ed = e & 3;
            const int cc = bq * g.N + c0 + cell;
            // vertices: top (cr,cc)-(cr,cc+1), bottom (cr+1,cc)-(cr+1,cc+1), left (cr,cc)-(cr+1,cc), right (cr,cc+1)-(cr+1,cc+1)
            const int r1 = cr + (ed == 1), c1 = cc + (ed == 3);
            const int r2 = cr + (ed != 0), c2 = cc + (ed != 2);
            const double* ph = basis + size_t(i) * g.Dp;
            // column C of the last row may alias the next row's zero column only when P == C; read safely
            const double v1 = (c1 < g.C) ? ph[size_t(r1) * g.P + c1] : 0.0;
            const double v2 = (c2 < g.C) ? ph[size_t(r2) * g.P + c2] : 0.0;
            G[e * ldg + i] = v1 - v2;
        }
        __syncthreads();
        const int ne = ncell * 4;
#pragma unroll
        for (int s = 0; s < PROJ_MAXN * PROJ_MAXN / 256; ++s) {
            if (s < nslot) {
                const int ent = tid + s * 256;
                if (ent < nn) {
                    const int i = ent / n, j = ent % n;
                    double a = acc[s];
                    for (int e = 0; e < ne; ++e) a = fma(G[e * ldg + i], G[e * ldg + j], a);
                    acc[s] = a;
                }
            }
        }
    }
    double* dst = part + (size_t(q) * g.N + crl) * nn;
#pragma unroll
    for (int s = 0; s < PROJ_MAXN * PROJ_MAXN / 256; ++s) {
        const int ent = tid + s * 256;
        if (s < nslot && ent < nn) dst[ent] = 0.5 * acc[s];
    }
}

__global__ void k_reduce_ahat(const double* __restrict__ part, int nparts, int nn, double* __restrict__ Ahat) {
    const int q = blockIdx.y;
    const int ent = blockIdx.x * blockDim.x + threadIdx.x;
    if (ent >= nn) return;
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += part[(size_t(q) * nparts + p) * nn + ent];
    Ahat[size_t(q) * nn + ent] = s;
}

// bhat_i = Phi_i . b,  b == 1/N^2 on interior DOFs (padding slots hold zeros)
__global__ void __launch_bounds__(256) k_project_rhs(LevelGeo g, const double* __restrict__ basis, double scale,
                                                     double* __restrict__ bhat) {
    __shared__ double red[32];
    const double* ph = basis + size_t(blockIdx.x) * g.Dp;
    double acc = 0.0;
    for (int i = threadIdx.x; i < g.Dp; i += blockDim.x) acc += ph[i];
    const double tot = block_sum(acc, red, threadIdx.x, blockDim.x);
    if (threadIdx.x == 0) bhat[blockIdx.x] = tot * scale;
}

// ---- the same reduced operators as dense contractions on the fp64 tensor cores: W_q = A_q Phi^T by the matrix-free stencil
// with the unit coefficient vector e_q (nb applications over the n basis rows), then ONE split-K DMMA product
// Phi (n x Dp) . [W_0; ...; W_{nb-1}]^T (nb n x Dp) -> (n, nb n), permuted to (nb, n, n).  Any n.
__global__ void k_unit_rows(double* __restrict__ Y, int nb, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;               // row (q, j) of the (nb n, nb) coefficient matrix
    if (i >= nb * n * nb) return;
    const int row = i / nb, col = i - row * nb;
    Y[i] = (row / n == col) ? 1.0 : 0.0;
}
__global__ void k_permute_ahat(const double* __restrict__ Cm, int n, int nb, double* __restrict__ Ahat) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nb * n * n) return;
    const int q = e / (n * n), i = (e / n) % n, j = e % n;
    Ahat[e] = Cm[size_t(i) * nb * n + size_t(q) * n + j];
}

int Context::project_operators_dmma(const double* basis, int n, double* Ahat, cudaStream_t st) {
    const LevelGeo& g = levels[0];
    const int nb = nrb * ncb;
    const size_t wd = size_t(nb) * n * g.Dp, yd = size_t(nb) * n * nb, cd = size_t(n) * nb * n;
    int rc = ensure_scratch((wd + yd + cd) * 8); if (rc) return rc;
    double* W = (double*)scratch;
    double* Y = W + wd;
    double* Cm = Y + yd;
    CK(cudaMemsetAsync(W, 0, wd * 8, st));
    ++g_launches; k_unit_rows<<<(unsigned)((yd + 255) / 256), 256, 0, st>>>(Y, nb, n);
    for (int q = 0; q < nb; ++q) {
        rc = apply(Y + size_t(q) * n * nb, basis, W + size_t(q) * n * g.Dp, n, st);
        if (rc) return rc;
    }
    rc = gemm_nt(basis, g.Dp, W, g.Dp, Cm, int64_t(nb) * n, n, int64_t(nb) * n, g.Dp, 2, st);
    if (rc) return rc;
    ++g_launches; k_permute_ahat<<<(unsigned)((cd + 255) / 256), 256, 0, st>>>(Cm, n, nb, Ahat);
    CK(cudaGetLastError());
    return ROMHC_OK;
}

int Context::project_operators(const double* basis, int n, double* Ahat, double* bhat, cudaStream_t st) {
    if (n < 1) { set_error("project_operators: n must be positive, got %d", n); return ROMHC_ERR_ARG; }
    const LevelGeo& g = levels[0];
    const int nb = nrb * ncb, nn = n * n;
    if (n > PROJ_MAXN || proj_variant == 1) {
        int rc = project_operators_dmma(basis, n, Ahat, st); if (rc) return rc;
    } else {
        int rc = ensure_scratch(size_t(nb) * g.N * nn * 8); if (rc) return rc;
        const size_t sm = size_t(PROJ_EB) * (n | 1) * 8;
        ++g_launches; k_project_partial<<<dim3(g.N, nb), 256, sm, st>>>(g, basis, n, (double*)scratch);
        ++g_launches; k_reduce_ahat<<<dim3((nn + 127) / 128, nb), 128, 0, st>>>((double*)scratch, g.N, nn, Ahat);
    }
    if (bhat) { ++g_launches; k_project_rhs<<<n, 256, 0, st>>>(g, basis, 1.0 / (double(N) * double(N)), bhat); }
    CK(cudaGetLastError());
    return ROMHC_OK;
}

// ======================================================================================================
// K5: batched reduced solves  (sum_q y_kq Ahat_q) c_k = rhs_k.  n <= 24: a quad of lanes per system, assembly on the
// fp64 tensor cores, Cholesky in registers (k_reduced_solve_quad); larger n: one warp per system, Cholesky in smem.
// Ahat is used through its lower triangle (packed) -- the reference symmetrises implicitly by calling
// scipy.linalg.solve(assume_a='pos') (SolutionsManagers.py:29).
// ======================================================================================================
#define RS_WARPS 8
__global__ void __launch_bounds__(RS_WARPS * 32)
k_reduced_solve(const double* __restrict__ y, int nb, const double* __restrict__ Ahat, const double* __restrict__ rhs,
                int rhs_per_system, int n, int64_t K, double* __restrict__ C, int* __restrict__ info,
                int ahat_in_smem) {
    extern __shared__ __align__(16) double sm[];
    const int npk = n * (n + 1) / 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* Apk = sm;                                             // nb * npk packed lower triangles (optional)
    double* wbase = sm + (ahat_in_smem ? size_t(nb) * npk : 0);
    double* M = wbase + size_t(warp) * (npk + nb + 2);
    double* yb = M + npk;
    if (ahat_in_smem) {
        for (int idx = threadIdx.x; idx < nb * npk; idx += blockDim.x) {
            const int q = idx / npk, e = idx % npk;
            int i = int((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
            while (i * (i + 1) / 2 > e) --i;
            while ((i + 1) * (i + 2) / 2 <= e) ++i;
            const int j = e - i * (i + 1) / 2;
            Apk[idx] = Ahat[(size_t(q) * n + i) * n + j];
        }
    }
    __syncthreads();
    const int64_t stride = int64_t(gridDim.x) * RS_WARPS;
    for (int64_t k = int64_t(blockIdx.x) * RS_WARPS + warp; k < K; k += stride) {
        for (int q = lane; q < nb; q += 32) yb[q] = y[k * nb + q];
        __syncwarp();
        // assemble the lower triangle
        for (int e = lane; e < npk; e += 32) {
            double a = 0.0;
            if (ahat_in_smem) {
                for (int q = 0; q < nb; ++q) a = fma(yb[q], Apk[q * npk + e], a);
            } else {
                int i = int((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
                while (i * (i + 1) / 2 > e) --i;
                while ((i + 1) * (i + 2) / 2 <= e) ++i;
                const int j = e - i * (i + 1) / 2;
                for (int q = 0; q < nb; ++q) a = fma(yb[q], Ahat[(size_t(q) * n + i) * n + j], a);
            }
            M[e] = a;
        }
        __syncwarp();
        // Cholesky, right looking; lane owns rows lane, lane+32
        int bad = 0;
        for (int kc = 0; kc < n; ++kc) {
            const double d = M[kc * (kc + 1) / 2 + kc];
            if (!(d > 0.0)) bad = 1;
            const double ld = sqrt(d), ild = 1.0 / ld;
            __syncwarp();
            for (int i = kc + 1 + lane; i < n; i += 32) M[i * (i + 1) / 2 + kc] *= ild;
            if (lane == 0) M[kc * (kc + 1) / 2 + kc] = ild;      // store 1 / L_kk
            __syncwarp();
            for (int i = kc + 1 + lane; i < n; i += 32) {
                const double lik = M[i * (i + 1) / 2 + kc];
                double* row = M + i * (i + 1) / 2;
                for (int j = kc + 1; j <= i; ++j) row[j] = fma(-lik, M[j * (j + 1) / 2 + kc], row[j]);
            }
            __syncwarp();
        }
        // L L^T c = rhs (column oriented, rows in registers)
        const double* b = rhs + (rhs_per_system ? k * n : 0);
        const int j0 = lane, j1 = lane + 32;
        double b0 = j0 < n ? b[j0] : 0.0, b1 = j1 < n ? b[j1] : 0.0;
        for (int i = 0; i < n; ++i) {
            const double bi = __shfl_sync(0xffffffffu, i < 32 ? b0 : b1, i & 31);
            const double yi = bi * M[i * (i + 1) / 2 + i];
            if (j0 == i) b0 = yi;
            if (j1 == i) b1 = yi;
            if (j0 > i && j0 < n) b0 = fma(-M[j0 * (j0 + 1) / 2 + i], yi, b0);
            if (j1 > i && j1 < n) b1 = fma(-M[j1 * (j1 + 1) / 2 + i], yi, b1);
        }
        for (int i = n - 1; i >= 0; --i) {
            const double bi = __shfl_sync(0xffffffffu, i < 32 ? b0 : b1, i & 31);
            const double xi = bi * M[i * (i + 1) / 2 + i];
            if (j0 == i) b0 = xi;
            if (j1 == i) b1 = xi;
            if (j0 < i) b0 = fma(-M[i * (i + 1) / 2 + j0], xi, b0);
            if (j1 < i) b1 = fma(-M[i * (i + 1) / 2 + j1], xi, b1);
        }
        if (j0 < n) C[k * n + j0] = b0;
        if (j1 < n) C[k * n + j1] = b1;
        if (lane == 0 && info) info[k] = bad;
        __syncwarp();
    }
}

// ---- quad-per-system variant (n <= 24): assembly on the fp64 tensor cores, factorisation in registers ------------------
// A warp owns 8 systems, 4 lanes each.  The assembly A_k = sum_q y_kq Ahat_q is the GEMM (8 systems x nb) . (nb x entries)
// and runs as DMMA m8n8k4 tiles: the A fragment is y (lane L: system L/4, block 4 qg + L%4), the B fragment comes from a
// shared-memory table built once per CTA, and the C fragment IS the register layout of the factorisation: lane l of a quad
// holds rows 4m + l (m = 0..NR-1) of its system, row block m as 4m + 4 column slots (slots right of the diagonal mirror the
// symmetric entry and are never read).  A tile is (row block m, column pair p): quad lane l receives (row 4m+l, cols 2p, 2p+1).
// Right-looking Cholesky with the forward substitution interleaved: per column one quad-wide broadcast of the pivot and of
// every L_jk (shuffles of width 4), all trailing updates are independent register FMAs; the diagonal keeps 1 / L_kk.
// Back substitution: per row one partial dot product per lane and a quad reduction.  No shared memory after the assembly,
// no synchronisation besides the shuffles; occupancy is set by registers (12 warps per SM at n = 20).
__device__ __forceinline__ void dmma884_rs(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
#define RQ_THREADS 128
#define RQ_MAXN 24
// compile-time loop: the body sees its index as a constant expression, so every register-array subscript below is static
// (nvcc does not fully unroll loops of this size from a pragma alone and would put the matrix into local memory)
template <int I, int E, class F> __device__ __forceinline__ void static_for(F&& f) {
    if constexpr (I < E) { f(std::integral_constant<int, I>{}); static_for<I + 1, E>(f); }
}
template <int N> struct RqCfg {
    static constexpr int NR = (N + 3) / 4;                       // row blocks
    static constexpr int NE = (N + 1) & ~1;                      // columns, rounded up to a DMMA column pair
    __host__ __device__ static constexpr int width(int m) { return 4 * m + 4 < NE ? 4 * m + 4 : NE; }
    __host__ __device__ static constexpr int tile0(int m) { return m == 0 ? 0 : tile0(m - 1) + width(m - 1) / 2; }
    static constexpr int NT = tile0(NR);                         // tiles (row block, column pair)
    static constexpr int OCC = N <= 12 ? 4 : (N <= 16 ? 3 : 2);  // CTAs per SM without register spills (n = 20 on B200, 1M systems: 0.885 ms at 2, 0.920 at 3 with spills, 1.52 at 1)
};

template <int N, int OCC>
__global__ void __launch_bounds__(RQ_THREADS, OCC)
k_reduced_solve_quad(const double* __restrict__ y, int nb, const double* __restrict__ Ahat, const double* __restrict__ rhs,
                     int rhs_per_system, int64_t K, double* __restrict__ C, int* __restrict__ info, int nqg) {
    using Cfg = RqCfg<N>;
    constexpr int NR = Cfg::NR, NT = Cfg::NT, NE = Cfg::NE;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) double sm[];       // Bt[nqg][NT][32]: B fragments, lane-contiguous
    for (int idx = threadIdx.x; idx < nqg * NT * 32; idx += blockDim.x) {
        const int L = idx & 31, tile = (idx >> 5) % NT, qg = (idx >> 5) / NT;
        const int kq = L & 3, slot = L >> 2;           // B fragment of m8n8k4: lane L holds B[k = L % 4][n = L / 4]
        int m = 0;
        while (m + 1 < NR && Cfg::tile0(m + 1) <= tile) ++m;
        const int p = tile - Cfg::tile0(m);
        const int row = 4 * m + (slot >> 1), col = 2 * p + (slot & 1), q = 4 * qg + kq;
        double v = 0.0;
        if (q < nb && row < N && col < N)
            v = (col <= row) ? Ahat[(size_t(q) * N + row) * N + col] : Ahat[(size_t(q) * N + col) * N + row];
        sm[idx] = v;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, l = lane & 3, g = lane >> 2;
    const int64_t nwarp = int64_t(gridDim.x) * (RQ_THREADS / 32);
    for (int64_t s0 = (int64_t(blockIdx.x) * (RQ_THREADS / 32) + (threadIdx.x >> 5)) * 8; s0 < K; s0 += nwarp * 8) {
        const int64_t k = s0 + g;
        const bool live = k < K;
        double R[NR][NE];                              // row block m uses the first width(m) slots
        static_for<0, NR>([&](auto m_) {
            constexpr int m = decltype(m_)::value;
            static_for<0, Cfg::width(m)>([&](auto c_) { R[m][decltype(c_)::value] = 0.0; });
        });
        // ---- assembly: NT tiles x nqg DMMAs ----
        for (int qg = 0; qg < nqg; ++qg) {
            const int q = 4 * qg + l;
            const double a = (live && q < nb) ? y[k * nb + q] : 0.0;
            const double* bt = sm + size_t(qg) * NT * 32 + lane;
            static_for<0, NR>([&](auto m_) {
                constexpr int m = decltype(m_)::value;
                static_for<0, Cfg::width(m) / 2>([&](auto p_) {
                    constexpr int p = decltype(p_)::value;
                    dmma884_rs(R[m][2 * p], R[m][2 * p + 1], a, bt[(Cfg::tile0(m) + p) * 32]);
                });
            });
        }
        // rows >= N of the last block stay zero (the table holds zeros there) and never act as pivots
        double B[NR];
        const double* rb = rhs + (rhs_per_system ? k * N : 0);
#pragma unroll
        for (int m = 0; m < NR; ++m) B[m] = (live && 4 * m + l < N) ? rb[4 * m + l] : 0.0;
        // ---- Cholesky + forward substitution ----
        int bad = 0;
        static_for<0, N>([&](auto kc_) {
            constexpr int kc = decltype(kc_)::value, mk = kc >> 2, lk = kc & 3;
            const double d = __shfl_sync(FULL, R[mk][kc], lk, 4);
            if (!(d > 0.0)) bad = 1;
            const double ild = rsqrt(d);
            static_for<mk, NR>([&](auto m_) { constexpr int m = decltype(m_)::value; R[m][kc] *= ild; });
            if (l == lk) R[mk][kc] = ild;
            const double ck = __shfl_sync(FULL, B[mk] * ild, lk, 4);
            if (l == lk) B[mk] = ck;
            static_for<mk, NR>([&](auto m_) {
                constexpr int m = decltype(m_)::value;
                if (m > mk || l > lk) B[m] = fma(-R[m][kc], ck, B[m]);
            });
            static_for<kc + 1, N>([&](auto j_) {
                constexpr int j = decltype(j_)::value;
                const double Lj = __shfl_sync(FULL, R[j >> 2][kc], j & 3, 4);
                static_for<(j >> 2), NR>([&](auto m_) {
                    constexpr int m = decltype(m_)::value;
                    R[m][j] = fma(-R[m][kc], Lj, R[m][j]);
                });
            });
        });
        // ---- back substitution: L^T x = c ----
        static_for<0, N>([&](auto ii_) {
            constexpr int i = N - 1 - decltype(ii_)::value, mi = i >> 2, li = i & 3;
            double sacc = 0.0;
            static_for<mi, NR>([&](auto m_) {
                constexpr int m = decltype(m_)::value;
                if (m > mi || l > li) sacc = fma(R[m][i], B[m], sacc);
            });
            sacc += __shfl_xor_sync(FULL, sacc, 1, 4);
            sacc += __shfl_xor_sync(FULL, sacc, 2, 4);
            if (l == li) B[mi] = (B[mi] - sacc) * R[mi][i];
        });
        if (live) {
#pragma unroll
            for (int m = 0; m < NR; ++m)
                if (4 * m + l < N) C[k * N + 4 * m + l] = B[m];
            if (info && l == 0) info[k] = bad;
        }
    }
}

template <int N, int OCC>
static int launch_reduced_quad(const double* y, int nb, const double* Ahat, const double* rhs, int rps, int64_t K,
                               double* C, int* info, int nsm, cudaStream_t st) {
    const int nqg = (nb + 3) / 4;
    const size_t smb = size_t(nqg) * RqCfg<N>::NT * 32 * 8;
    if (smb > 227 * 1024) return -1;
    CK(cudaFuncSetAttribute(k_reduced_solve_quad<N, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb));
    const int per_sm = std::max<int>(1, std::min<int>(OCC, int((227 * 1024) / (smb + 1024))));
    const int64_t want = (K + 8 * (RQ_THREADS / 32) - 1) / (8 * (RQ_THREADS / 32));
    const int grid = int(std::min<int64_t>(want, int64_t(nsm) * per_sm));
    ++g_launches;
    k_reduced_solve_quad<N, OCC><<<grid, RQ_THREADS, smb, st>>>(y, nb, Ahat, rhs, rps, K, C, info, nqg);
    CK(cudaGetLastError());
    return ROMHC_OK;
}

template <int N>
static int dispatch_reduced_quad(int n, const double* y, int nb, const double* Ahat, const double* rhs, int rps, int64_t K,
                                 double* C, int* info, int nsm, cudaStream_t st) {
    if (n == N) {
        return launch_reduced_quad<N, RqCfg<N>::OCC>(y, nb, Ahat, rhs, rps, K, C, info, nsm, st);
    }
    if constexpr (N > 1) return dispatch_reduced_quad<N - 1>(n, y, nb, Ahat, rhs, rps, K, C, info, nsm, st);
    return -1;
}

int reduced_solve(const double* y, int nb, const double* Ahat, const double* rhs, int rhs_per_system, int n,
                  int64_t K, double* C, int* info, cudaStream_t st) {
    if (n < 1) { set_error("reduced_solve: n must be positive, got %d", n); return ROMHC_ERR_ARG; }
    if (K <= 0) return ROMHC_OK;
    if (n > 64) return dense_spd_solve(y, nb, Ahat, rhs, rhs_per_system, n, K, C, info, st);   // blocked Cholesky (dense.cu)
    const int npk = n * (n + 1) / 2;
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    if (n <= RQ_MAXN) {
        const int rc = dispatch_reduced_quad<RQ_MAXN>(n, y, nb, Ahat, rhs, rhs_per_system, K, C, info, nsm, st);
        if (rc >= 0) return rc;          // -1: the fragment table does not fit shared memory -> warp-per-system kernel
    }
    const size_t per_warp = size_t(npk + nb + 2) * 8;
    const size_t tab = size_t(nb) * npk * 8;
    const int in_smem = (tab + RS_WARPS * per_warp) <= 96 * 1024;
    const size_t smb = (in_smem ? tab : 0) + RS_WARPS * per_warp;
    // per call: the attribute is per device and a process may drive several
    CK(cudaFuncSetAttribute(k_reduced_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const int64_t want = (K + RS_WARPS - 1) / RS_WARPS;
    const int grid = int(std::min<int64_t>(want, int64_t(nsm) * 8));
    ++g_launches; k_reduced_solve<<<grid, RS_WARPS * 32, smb, st>>>(y, nb, Ahat, rhs, rhs_per_system, n, K, C, info, in_smem);
    CK(cudaGetLastError());
    return ROMHC_OK;
}

// ======================================================================================================
// K6: P1 point evaluation (SolutionsManagers.py:221-244).  A tiny setup kernel turns every point into three
// (padded index, weight) pairs with numpy's linspace / searchsorted arithmetic, then a gather kernel.
// ======================================================================================================
__device__ __forceinline__ double linspace_at(double start, double stop, int num, int i) {
    // numpy.linspace: y = arange(num) * step + start with the last sample forced to `stop`
    if (i == num - 1) return stop;
    const double step = (stop - start) / double(num - 1);
    return __dadd_rn(__dmul_rn(double(i), step), start);
}
__device__ __forceinline__ int searchsorted_left(double start, double stop, int num, double x) {
    int lo = 0, hi = num;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (linspace_at(start, stop, num, mid) < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void k_eval_setup(LevelGeo g, const double* __restrict__ pts, int m, int* __restrict__ idx,
                             double* __restrict__ wts) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const double x = pts[2 * j], yv = pts[2 * j + 1];
    const double x0 = -g.ncb / 2.0, x1 = g.ncb / 2.0, y0 = -g.nrb / 2.0, y1 = g.nrb / 2.0;
    const int ncv = g.C + 1, nrv = g.R + 1;
    int px = searchsorted_left(x0, x1, ncv, x) - 1;
    int py = searchsorted_left(y0, y1, nrv, yv) - 1;
    int* id = idx + 3 * j;
    double* w = wts + 3 * j;
    if (px < 0 || py < 0 || px > g.C - 1 || py > g.R - 1) {
        // on the left/top boundary line the reference's wrap-around indexing only touches boundary vertices (value 0);
        // points outside the domain raise IndexError there and evaluate to 0 here.
        id[0] = id[1] = id[2] = -1; w[0] = w[1] = w[2] = 0.0;
        return;
    }
    const double cx0 = linspace_at(x0, x1, ncv, px), cx1 = linspace_at(x0, x1, ncv, px + 1);
    const double cy0 = linspace_at(y0, y1, nrv, py), cy1 = linspace_at(y0, y1, nrv, py + 1);
    const double qx = (x - cx0) / (cx1 - cx0), qy = (yv - cy0) / (cy1 - cy0);
    int vr[3], vc[3];
    if (qx + qy < 1) {
        w[0] = 1 - qx - qy; vr[0] = py;     vc[0] = px;
        w[1] = qx;          vr[1] = py;     vc[1] = px + 1;
        w[2] = qy;          vr[2] = py + 1; vc[2] = px;
    } else {
        w[0] = qx + qy - 1; vr[0] = py + 1; vc[0] = px + 1;
        w[1] = 1 - qx;      vr[1] = py + 1; vc[1] = px;
        w[2] = 1 - qy;      vr[2] = py;     vc[2] = px + 1;
    }
    for (int t = 0; t < 3; ++t) {
        const bool interior = vr[t] >= 1 && vr[t] <= g.R - 1 && vc[t] >= 1 && vc[t] <= g.C - 1;
        id[t] = interior ? vr[t] * g.P + vc[t] : -1;
    }
}

__global__ void k_eval_gather(const double* __restrict__ u, int64_t Dp, int64_t K, int m, const int* __restrict__ idx,
                              const double* __restrict__ wts, double* __restrict__ out) {
    const int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (t >= K * m) return;
    const int64_t k = t / m;
    const int j = int(t - k * m);
    const double* us = u + k * Dp;
    const int* id = idx + 3 * j;
    const double* w = wts + 3 * j;
    const double v0 = id[0] >= 0 ? us[id[0]] : 0.0;
    const double v1 = id[1] >= 0 ? us[id[1]] : 0.0;
    const double v2 = id[2] >= 0 ? us[id[2]] : 0.0;
    out[t] = __dadd_rn(__dadd_rn(__dmul_rn(w[0], v0), __dmul_rn(w[1], v1)), __dmul_rn(w[2], v2));
}

int Context::evaluate(const double* points, int m, const double* u, int64_t K, double* out, cudaStream_t st) {
    if (m <= 0 || K <= 0) return ROMHC_OK;
    const LevelGeo& g = levels[0];
    int rc = ensure_scratch(size_t(m) * 3 * (8 + 8)); if (rc) return rc;
    double* wts = (double*)scratch;
    int* idx = (int*)(wts + size_t(m) * 3);
    ++g_launches; k_eval_setup<<<(m + 127) / 128, 128, 0, st>>>(g, points, m, idx, wts);
    const int64_t tot = K * m;
    ++g_launches; k_eval_gather<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(u, g.Dp, K, m, idx, wts, out);
    CK(cudaGetLastError());
    return ROMHC_OK;
}

// (padded index, weight) triples of every point -- the rows of the l2 "Riesz" matrix, generate_riesz(norm="l2")
// (SolutionsManagers.py:70-77) without evaluating the D unit vectors.
int Context::interp_weights(const double* points, int m, int* idx3, double* w3, cudaStream_t st) {
    if (m <= 0) return ROMHC_OK;
    ++g_launches; k_eval_setup<<<(m + 127) / 128, 128, 0, st>>>(levels[0], points, m, idx3, w3);
    CK(cudaGetLastError());
    return ROMHC_OK;
}

// euclidean row norms of a generic row-major matrix (SolutionsManager.l2norm is a staticmethod: no geometry)
__global__ void __launch_bounds__(256) k_row_norms(const double* __restrict__ X, int64_t ld, int64_t D,
                                                   double* __restrict__ out) {
    __shared__ double red[32];
    const double* x = X + int64_t(blockIdx.x) * ld;
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < D; i += blockDim.x) acc = fma(x[i], x[i], acc);
    const double tot = block_sum(acc, red, threadIdx.x, blockDim.x);
    if (threadIdx.x == 0) out[blockIdx.x] = sqrt(tot);
}
int row_norms(const double* X, int64_t ld, int64_t K, int64_t D, double* out, cudaStream_t st) {
    if (K <= 0) return ROMHC_OK;
    ++g_launches; k_row_norms<<<(unsigned)K, 256, 0, st>>>(X, ld, D, out);
    CK(cudaGetLastError());
    return ROMHC_OK;
}

// ======================================================================================================
// argmax with numpy semantics: first maximum wins, NaN beats everything (np.argmax returns the first NaN)
// ======================================================================================================
__device__ __forceinline__ bool better(double v, int64_t i, double bv, int64_t bi) {
    const bool vn = v != v, bn = bv != bv;
    if (bi < 0) return true;
    if (vn != bn) return vn;
    if (vn && bn) return i < bi;
    return v > bv || (v == bv && i < bi);
}
__global__ void __launch_bounds__(256) k_argmax(const double* __restrict__ v, int64_t K, double* __restrict__ pv,
                                                int64_t* __restrict__ pi) {
    __shared__ double sv[256];
    __shared__ int64_t si[256];
    double bv = 0.0; int64_t bi = -1;
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < K; i += int64_t(gridDim.x) * blockDim.x)
        if (better(v[i], i, bv, bi)) { bv = v[i]; bi = i; }
    sv[threadIdx.x] = bv; si[threadIdx.x] = bi;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            const double ov = sv[threadIdx.x + s]; const int64_t oi = si[threadIdx.x + s];
            if (oi >= 0 && better(ov, oi, sv[threadIdx.x], si[threadIdx.x])) { sv[threadIdx.x] = ov; si[threadIdx.x] = oi; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { pv[blockIdx.x] = sv[0]; pi[blockIdx.x] = si[0]; }
}
__global__ void k_argmax_final(const double* __restrict__ pv, const int64_t* __restrict__ pi, int n,
                               int64_t* __restrict__ idx_out, double* __restrict__ val_out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double bv = 0.0; int64_t bi = -1;
    for (int i = 0; i < n; ++i)
        if (pi[i] >= 0 && better(pv[i], pi[i], bv, bi)) { bv = pv[i]; bi = pi[i]; }
    *idx_out = bi; *val_out = bv;
}

static void* g_argmax_scratch = nullptr;
int argmax_first(const double* v, int64_t K, int64_t* idx_out, double* val_out, cudaStream_t st) {
    if (K <= 0) { set_error("argmax of an empty sequence"); return ROMHC_ERR_ARG; }
    const int nblk = int(std::min<int64_t>((K + 255) / 256, 1024));
    if (!g_argmax_scratch) CK(cudaMalloc(&g_argmax_scratch, 1024 * 16));
    double* pv = (double*)g_argmax_scratch;
    int64_t* pi = (int64_t*)(pv + 1024);
    ++g_launches; k_argmax<<<nblk, 256, 0, st>>>(v, K, pv, pi);
    ++g_launches; k_argmax_final<<<1, 32, 0, st>>>(pv, pi, nblk, idx_out, val_out);
    CK(cudaGetLastError());
    return ROMHC_OK;
}

// ======================================================================================================
// Estimators (Estimators.py:24-37): out[k, q] = sum_b c[b, k] * A[b, q]   (invert: with 1/A and 1/result)
// c is (n, K) exactly as np.linalg.lstsq returns it in BaseReducedBasis.state_estimation.
// ======================================================================================================
__global__ void k_estimator(const double* __restrict__ c, int64_t K, int n, const double* __restrict__ ab, int nb,
                            int invert, double* __restrict__ out) {
    const int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (t >= K * nb) return;
    const int64_t k = t / nb;
    const int q = int(t - k * nb);
    double acc = 0.0;
    for (int b = 0; b < n; ++b) {
        const double a = ab[b * nb + q];
        acc += c[int64_t(b) * K + k] * (invert ? 1.0 / a : a);
    }
    out[t] = invert ? 1.0 / acc : acc;
}
int estimator_contract(const double* c, int64_t K, int n, const double* abasis, int nb, int invert, double* out,
                       cudaStream_t st) {
    if (K <= 0) return ROMHC_OK;
    const int64_t tot = K * nb;
    ++g_launches; k_estimator<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(c, K, n, abasis, nb, invert, out);
    CK(cudaGetLastError());
    return ROMHC_OK;
}

// ======================================================================================================
// Monomial features of the basis functions at every DOF: out[f][i] = prod_{j < degree, terms[f][j] >= 0} basis[terms[f][j]][i]
// (sklearn PolynomialFeatures(include_bias=False) evaluated on the rows of the basis; the predict step of the notebook's
// polynomial least squares, InverseProblemPipeline.ipynb cell 52).  terms: (nterms, degree) int32, -1 = unused factor.
// ======================================================================================================
__global__ void __launch_bounds__(256)
k_poly_features(const double* __restrict__ basis, int64_t ld, int64_t D, const int* __restrict__ terms, int degree,
                double* __restrict__ out, int64_t ldo) {
    const int f = blockIdx.y;
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= D) return;
    double v = 1.0;
    for (int j = 0; j < degree; ++j) {
        const int t = terms[f * degree + j];
        if (t >= 0) v *= basis[int64_t(t) * ld + i];
    }
    out[int64_t(f) * ldo + i] = v;
}
int poly_features(const double* basis, int64_t ld, int n, int64_t D, const int* terms, int nterms, int degree, double* out,
                  int64_t ldo, cudaStream_t st) {
    (void)n;
    if (nterms <= 0 || D <= 0) return ROMHC_OK;
    if (degree < 1 || degree > 8) { set_error("poly_features: degree must be in [1, 8]"); return ROMHC_ERR_ARG; }
    for (int f0 = 0; f0 < nterms; f0 += 65535) {
        const int nf = std::min(65535, nterms - f0);
        ++g_launches;
        k_poly_features<<<dim3((unsigned)((D + 255) / 256), nf), 256, 0, st>>>(basis, ld, D, terms + size_t(f0) * degree, degree,
                                                                          out + int64_t(f0) * ldo, ldo);
    }
    CK(cudaGetLastError());
    return ROMHC_OK;
}

}  // namespace romhc
