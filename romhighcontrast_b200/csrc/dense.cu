// Dense SPD systems of any size:  (sum_q y[k][q] A[q]) c_k = rhs_k  for n > 64 (the batched register / shared-memory
// kernels of reduced.cu serve n <= 64).
//
// Replaces galerkin() on a caller's dense operators (/root/reference/src/lib/SolutionsManagers.py:17-40: einsum assembly +
// scipy.linalg.solve(assume_a='pos')) and with it the generic SolutionsManager(A_preassembled, B_total) of :43-68.
// Blocked right-looking Cholesky, block size 32, a few systems at a time (grid.y = system): assemble -> per block column
// {factor the diagonal block in shared memory, triangular solve of the panel below it, symmetric rank-32 update of the
// trailing lower triangle} -> blocked forward / backward substitution.  Plain FP64 FMA kernels: this is an API-completeness
// path, not a headline one.
#include "common.cuh"
#include "romhc_internal.h"

#include <algorithm>

namespace romhc {

#define DN_B 32

// M[s] = sum_q y[s][q] A[q]   (full n x n, row-major)
__global__ void __launch_bounds__(256)
k_dn_assemble(const double* __restrict__ y, int nb, const double* __restrict__ A, int64_t nn, double* __restrict__ M) {
    const int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (e >= nn) return;
    const double* ys = y + int64_t(blockIdx.y) * nb;
    double acc = 0.0;
    for (int q = 0; q < nb; ++q) acc = fma(ys[q], A[int64_t(q) * nn + e], acc);
    M[int64_t(blockIdx.y) * nn + e] = acc;
}

// Cholesky of the diagonal block j0 (bs x bs) in shared memory; the factor overwrites the lower triangle
__global__ void __launch_bounds__(DN_B * DN_B)
k_dn_potrf(double* __restrict__ M, int n, int j0, int bs, int* __restrict__ info) {
    __shared__ double L[DN_B][DN_B + 1];
    double* Ms = M + int64_t(blockIdx.y) * n * n;
    const int r = threadIdx.y, c = threadIdx.x;
    if (r < bs && c < bs) L[r][c] = Ms[int64_t(j0 + r) * n + j0 + c];
    __syncthreads();
    for (int k = 0; k < bs; ++k) {
        const double d = L[k][k];
        __syncthreads();
        if (r == k && c == k) {
            if (!(d > 0.0)) info[blockIdx.y] = 1;
            L[k][k] = sqrt(d);
        }
        __syncthreads();
        if (c == k && r > k && r < bs) L[r][k] /= L[k][k];
        __syncthreads();
        if (r > k && c > k && c <= r && r < bs) L[r][c] = fma(-L[r][k], L[c][k], L[r][c]);
        __syncthreads();
    }
    if (r < bs && c <= r) Ms[int64_t(j0 + r) * n + j0 + c] = L[r][c];
}

// panel below the diagonal block: X L^T = A_panel, one thread per row
__global__ void __launch_bounds__(128)
k_dn_trsm(double* __restrict__ M, int n, int j0, int bs) {
    __shared__ double L[DN_B][DN_B + 1];
    double* Ms = M + int64_t(blockIdx.y) * n * n;
    for (int i = threadIdx.x; i < bs * bs; i += blockDim.x) L[i / bs][i % bs] = Ms[int64_t(j0 + i / bs) * n + j0 + i % bs];
    __syncthreads();
    const int row = j0 + bs + blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    double x[DN_B];
    double* a = Ms + int64_t(row) * n + j0;
#pragma unroll
    for (int k = 0; k < DN_B; ++k) x[k] = k < bs ? a[k] : 0.0;
#pragma unroll
    for (int k = 0; k < DN_B; ++k) {
        if (k < bs) {
            double v = x[k];
#pragma unroll
            for (int c = 0; c < k; ++c) v = fma(-x[c], L[k][c], v);
            x[k] = v / L[k][k];
        }
    }
#pragma unroll
    for (int k = 0; k < DN_B; ++k)
        if (k < bs) a[k] = x[k];
}

// trailing update: A[ti][tj] -= P[ti] P[tj]^T for the lower tiles ti >= tj of the trailing matrix (P = panel columns j0..j0+31)
__global__ void __launch_bounds__(256)
k_dn_syrk(double* __restrict__ M, int n, int j0, int t0, int nt) {
    __shared__ double Pi[DN_B][DN_B + 1], Pj[DN_B][DN_B + 1];
    double* Ms = M + int64_t(blockIdx.y) * n * n;
    // blockIdx.x -> (ti, tj), ti >= tj, both in [0, nt)
    int ti = int((sqrt(8.0 * blockIdx.x + 1.0) - 1.0) * 0.5);
    while (ti * (ti + 1) / 2 > int(blockIdx.x)) --ti;
    while ((ti + 1) * (ti + 2) / 2 <= int(blockIdx.x)) ++ti;
    const int tj = int(blockIdx.x) - ti * (ti + 1) / 2;
    const int ri = t0 + ti * DN_B, rj = t0 + tj * DN_B;
    for (int i = threadIdx.x; i < DN_B * DN_B; i += blockDim.x) {
        const int r = i / DN_B, c = i % DN_B;
        Pi[r][c] = (ri + r < n) ? Ms[int64_t(ri + r) * n + j0 + c] : 0.0;
        Pj[r][c] = (rj + r < n) ? Ms[int64_t(rj + r) * n + j0 + c] : 0.0;
    }
    __syncthreads();
    const int c = threadIdx.x % DN_B, rb = threadIdx.x / DN_B;      // 8 row groups of 4 rows
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int r = rb * 4 + q;
        if (ri + r < n && rj + c < n && rj + c <= ri + r) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < DN_B; ++k) acc = fma(Pi[r][k], Pj[c][k], acc);
            Ms[int64_t(ri + r) * n + rj + c] -= acc;
        }
    }
}

// L L^T c = b, one CTA per system, blocked by DN_B; b is overwritten by the solution
__global__ void __launch_bounds__(256)
k_dn_solve(const double* __restrict__ M, int n, const double* __restrict__ rhs, int rps, double* __restrict__ Cout,
           int64_t sys0) {
    __shared__ double xb[DN_B];
    const double* Ms = M + int64_t(blockIdx.x) * n * n;
    double* b = Cout + (sys0 + blockIdx.x) * n;
    const double* r = rhs + (rps ? (sys0 + blockIdx.x) * n : 0);
    for (int i = threadIdx.x; i < n; i += blockDim.x) b[i] = r[i];
    __syncthreads();
    for (int j0 = 0; j0 < n; j0 += DN_B) {                          // forward
        const int bs = min(DN_B, n - j0);
        if (threadIdx.x == 0) {
            for (int k = 0; k < bs; ++k) {
                double v = b[j0 + k];
                for (int c = 0; c < k; ++c) v = fma(-Ms[int64_t(j0 + k) * n + j0 + c], xb[c], v);
                xb[k] = v / Ms[int64_t(j0 + k) * n + j0 + k];
            }
            for (int k = 0; k < bs; ++k) b[j0 + k] = xb[k];
        }
        __syncthreads();
        for (int i = j0 + bs + threadIdx.x; i < n; i += blockDim.x) {
            double v = b[i];
            for (int c = 0; c < bs; ++c) v = fma(-Ms[int64_t(i) * n + j0 + c], xb[c], v);
            b[i] = v;
        }
        __syncthreads();
    }
    for (int j0 = ((n - 1) / DN_B) * DN_B; j0 >= 0; j0 -= DN_B) {   // backward with L^T
        const int bs = min(DN_B, n - j0);
        if (threadIdx.x == 0) {
            for (int k = bs - 1; k >= 0; --k) {
                double v = b[j0 + k];
                for (int c = k + 1; c < bs; ++c) v = fma(-Ms[int64_t(j0 + c) * n + j0 + k], xb[c], v);
                xb[k] = v / Ms[int64_t(j0 + k) * n + j0 + k];
            }
            for (int k = 0; k < bs; ++k) b[j0 + k] = xb[k];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < j0; i += blockDim.x) {
            double v = b[i];
            for (int c = 0; c < bs; ++c) v = fma(-Ms[int64_t(j0 + c) * n + i], xb[c], v);
            b[i] = v;
        }
        __syncthreads();
    }
}

static void* g_dn_ws = nullptr;
static size_t g_dn_bytes = 0;
static int g_dn_dev = -1;

int dense_spd_solve(const double* y, int nb, const double* A, const double* rhs, int rhs_per_system, int n, int64_t K,
                    double* C, int* info, cudaStream_t st) {
    if (K <= 0) return ROMHC_OK;
    const int64_t nn = int64_t(n) * n;
    int dev = 0;
    CK(cudaGetDevice(&dev));
    int64_t chunk = std::max<int64_t>(1, std::min<int64_t>({K, int64_t((size_t(1) << 30) / (size_t(nn) * 8)), int64_t(1024)}));
    const size_t need = size_t(chunk) * nn * 8 + size_t(chunk) * 4;
    if (dev != g_dn_dev || need > g_dn_bytes) {
        if (g_dn_ws) cudaFree(g_dn_ws);
        g_dn_ws = nullptr; g_dn_bytes = 0;
        CK(cudaMalloc(&g_dn_ws, need));
        g_dn_bytes = need; g_dn_dev = dev;
    }
    double* M = (double*)g_dn_ws;
    int* flag = (int*)((char*)g_dn_ws + size_t(chunk) * nn * 8);
    for (int64_t k0 = 0; k0 < K; k0 += chunk) {
        const int kc = int(std::min<int64_t>(chunk, K - k0));
        CK(cudaMemsetAsync(flag, 0, size_t(kc) * 4, st));
        ++g_launches;
        k_dn_assemble<<<dim3((unsigned)((nn + 255) / 256), kc), 256, 0, st>>>(y + k0 * nb, nb, A, nn, M);
        for (int j0 = 0; j0 < n; j0 += DN_B) {
            const int bs = std::min(DN_B, n - j0);
            ++g_launches; k_dn_potrf<<<dim3(1, kc), dim3(DN_B, DN_B), 0, st>>>(M, n, j0, bs, flag);
            const int rows = n - j0 - bs;
            if (rows > 0) {
                ++g_launches; k_dn_trsm<<<dim3((rows + 127) / 128, kc), 128, 0, st>>>(M, n, j0, bs);
                const int nt = (rows + DN_B - 1) / DN_B;
                ++g_launches; k_dn_syrk<<<dim3(nt * (nt + 1) / 2, kc), 256, 0, st>>>(M, n, j0, j0 + bs, nt);
            }
        }
        ++g_launches; k_dn_solve<<<kc, 256, 0, st>>>(M, n, rhs, rhs_per_system, C, k0);
        if (info) CK(cudaMemcpyAsync(info + k0, flag, size_t(kc) * 4, cudaMemcpyDeviceToDevice, st));
    }
    CK(cudaGetLastError());
    return ROMHC_OK;
}

// out[k] = sum_i X[k][i] * Y[k][i]  (row-wise dot products: u^T (A u) of the generic manager's H10norm)
__global__ void __launch_bounds__(256) k_row_dots(const double* __restrict__ X, int64_t ldx, const double* __restrict__ Y,
                                                  int64_t ldy, int64_t D, double* __restrict__ out) {
    __shared__ double red[32];
    const double* x = X + int64_t(blockIdx.x) * ldx;
    const double* yv = Y + int64_t(blockIdx.x) * ldy;
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < D; i += blockDim.x) acc = fma(x[i], yv[i], acc);
    const double tot = block_sum(acc, red, threadIdx.x, blockDim.x);
    if (threadIdx.x == 0) out[blockIdx.x] = tot;
}
int row_dots(const double* X, int64_t ldx, const double* Y, int64_t ldy, int64_t K, int64_t D, double* out, cudaStream_t st) {
    if (K <= 0) return ROMHC_OK;
    ++g_launches; k_row_dots<<<(unsigned)K, 256, 0, st>>>(X, ldx, Y, ldy, D, out);
    CK(cudaGetLastError());
    return ROMHC_OK;
}

}  // namespace romhc

namespace romhc {

// ======================================================================================================
// R factor of a tall-skinny matrix by Householder TSQR (b <= 32 columns): the rank-revealing orthonormalisation step of
// the block Lanczos POD (pod._orthonormal_rows) needs R of the (Dp x b) block to full fp64 accuracy -- a b x b Gram
// matrix would resolve only sqrt(eps) of its dynamic range.  Level 1: every CTA folds its chunk of rows, 384 at a time,
// into a running b x b triangle ([R; tile] -> Householder QR in shared memory -> R); upper levels reduce groups of 12
// triangles the same way until one is left.  Deterministic (fixed tree), backward stable, no library call.
// element (i, j) of the input: in[j * ld + i] (transposed == 1: the rows of a (b, Dp) block are its columns) or in[i * ld + j].
// ======================================================================================================
#define TS_MAXB 32
#define TS_TL 384
#define TS_PITCH (TS_MAXB + 1)
__global__ void __launch_bounds__(256)
k_tsqr(const double* __restrict__ in, int64_t ld, int transposed, int64_t rows_total, int b, int64_t rows_per_cta,
       double* __restrict__ Rout) {
    extern __shared__ __align__(16) double smq[];
    double* M = smq;                                    // (b + TS_TL) x TS_PITCH
    double* vv = M + size_t(TS_MAXB + TS_TL) * TS_PITCH; // Householder vector
    __shared__ double red[32];
    __shared__ double s_tau, s_v0, s_beta;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t i0 = int64_t(blockIdx.x) * rows_per_cta;
    const int64_t i1 = (i0 + rows_per_cta < rows_total) ? i0 + rows_per_cta : rows_total;
    for (int i = tid; i < b * TS_PITCH; i += 256) M[i] = 0.0;
    __syncthreads();
    for (int64_t t0 = i0; t0 < i1 || t0 == i0; t0 += TS_TL) {
        const int nt = int((i1 - t0 < TS_TL) ? (i1 - t0 > 0 ? i1 - t0 : 0) : TS_TL);
        const int L = b + nt;
        // strictly lower part of the running triangle is zero by construction; load the tile below it
        if (transposed) {
            for (int idx = tid; idx < nt * b; idx += 256) {
                const int j = idx / nt, i = idx - j * nt;              // consecutive threads -> consecutive rows: coalesced
                M[(b + i) * TS_PITCH + j] = in[int64_t(j) * ld + t0 + i];
            }
        } else {
            for (int idx = tid; idx < nt * b; idx += 256) {
                const int i = idx / b, j = idx - i * b;
                M[(b + i) * TS_PITCH + j] = in[(t0 + i) * ld + j];
            }
        }
        __syncthreads();
        for (int j = 0; j < b; ++j) {
            // ||M[j:, j]||^2
            double acc = 0.0;
            for (int i = j + tid; i < L; i += 256) { const double v = M[i * TS_PITCH + j]; acc = fma(v, v, acc); }
            const double nrm2 = block_sum(acc, red, tid, 256);
            if (tid == 0) {
                const double x0 = M[j * TS_PITCH + j];
                const double nx = sqrt(nrm2);
                if (nx == 0.0) { s_tau = 0.0; s_v0 = 1.0; s_beta = 0.0; }
                else {
                    const double beta = x0 >= 0.0 ? -nx : nx;
                    s_beta = beta; s_v0 = x0 - beta; s_tau = (beta - x0) / beta;
                }
            }
            __syncthreads();
            const double tau = s_tau, iv0 = 1.0 / s_v0;
            for (int i = j + tid; i < L; i += 256) vv[i] = (i == j) ? 1.0 : M[i * TS_PITCH + j] * iv0;
            __syncthreads();
            if (tau != 0.0) {
                for (int k = j + 1 + warp; k < b; k += 8) {            // a warp owns whole columns: dot, then update
                    double w = 0.0;
                    for (int i = j + lane; i < L; i += 32) w = fma(vv[i], M[i * TS_PITCH + k], w);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
                    w *= tau;
                    for (int i = j + lane; i < L; i += 32) M[i * TS_PITCH + k] = fma(-w, vv[i], M[i * TS_PITCH + k]);
                }
            }
            __syncthreads();
            for (int i = j + tid; i < L; i += 256) M[i * TS_PITCH + j] = (i == j) ? (tau != 0.0 ? s_beta : M[i * TS_PITCH + j]) : 0.0;
            __syncthreads();
        }
        if (t0 + TS_TL >= i1) break;
    }
    double* Ro = Rout + int64_t(blockIdx.x) * b * b;
    for (int idx = tid; idx < b * b; idx += 256) {
        const int i = idx / b, j = idx - i * b;
        Ro[idx] = (j >= i) ? M[i * TS_PITCH + j] : 0.0;
    }
}

static void* g_ts_ws = nullptr;
static size_t g_ts_bytes = 0;
static int g_ts_dev = -1;

// W: (b, Dp) row-major block, ld = row pitch; R_out: (b, b) upper triangular with W W^T = R^T R
int tsqr_r(const double* W, int64_t ld, int b, int64_t Dp, double* R_out, cudaStream_t st) {
    if (b < 1 || b > TS_MAXB) { set_error("tsqr_r: block size must be in [1, %d], got %d", TS_MAXB, b); return ROMHC_ERR_ARG; }
    if (Dp < 1) { set_error("tsqr_r: empty block"); return ROMHC_ERR_ARG; }
    int dev = 0, nsm = 148;
    CK(cudaGetDevice(&dev));
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const size_t smb = (size_t(TS_MAXB + TS_TL) * TS_PITCH + (TS_MAXB + TS_TL)) * 8;
    CK(cudaFuncSetAttribute(k_tsqr, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smb)));
    int64_t nblk = std::max<int64_t>(1, std::min<int64_t>(nsm, (Dp + 63) / 64));
    const int64_t per = (Dp + nblk - 1) / nblk;
    nblk = (Dp + per - 1) / per;
    const size_t need = size_t(nblk + (nblk + 11) / 12 + 2) * b * b * 8 * 2;
    if (dev != g_ts_dev || need > g_ts_bytes) {
        if (g_ts_ws) cudaFree(g_ts_ws);
        g_ts_ws = nullptr; g_ts_bytes = 0;
        CK(cudaMalloc(&g_ts_ws, need));
        g_ts_bytes = need; g_ts_dev = dev;
    }
    double* bufA = (double*)g_ts_ws;
    double* bufB = bufA + size_t(nblk + 1) * b * b;
    ++g_launches;
    k_tsqr<<<(unsigned)nblk, 256, smb, st>>>(W, ld, 1, Dp, b, per, nblk == 1 ? R_out : bufA);
    int64_t cur = nblk;
    double* src = bufA; double* dst = bufB;
    while (cur > 1) {
        const int64_t groups = (cur + 11) / 12;
        ++g_launches;
        k_tsqr<<<(unsigned)groups, 256, smb, st>>>(src, b, 0, cur * b, b, int64_t(12) * b, groups == 1 ? R_out : dst);
        cur = groups;
        std::swap(src, dst);
    }
    CK(cudaGetLastError());
    return ROMHC_OK;
}

}  // namespace romhc
