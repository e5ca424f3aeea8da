// Greedy error sweep on the fp64 tensor cores:  out[k] = || sum_j coef[k][j] basis_j - U_k ||_{A_1}
//
// Replaces sm.H10norm(approx_solutions_coefs - solutions2train) of the reference's greedy loop
// (/root/reference/src/lib/ReducedBasis.py:129; H10norm: src/lib/SolutionsManagers.py:56-58) for all K snapshots at once.
//
// With a == 1 every mesh edge of the P1 stiffness has weight 1, and because all boundary / padding slots of the padded
// layout hold zeros the energy is a flat 1-D formula over the slots i of rows 0 .. R-1:
//     v^T A_1 v = sum_i (v_i - v_{i+1})^2 + (v_i - v_{i+P})^2          (slot i+1 of a row end is the next row's boundary 0)
// so the kernel never looks at the geometry.  v = C Phi - U is a GEMM (K x n) . (n x Dp) minus the streamed snapshots:
// a CTA owns MS = 8 MT systems and a segment of RS grid rows and walks down the rows; per row the basis slab (n x P,
// L2 resident: every CTA re-reads it, the snapshots come from HBM exactly once) is staged by cp.async into a double
// buffer, each warp initialises its DMMA accumulators with -u (fragment-shaped 16-byte loads), adds C Phi with
// m8n8k4 DMMAs, and the differences are taken on the accumulator fragments: the south difference against the previous
// row's fragments kept in registers, the east difference in-lane / by quad shuffles, and at warp boundaries through a
// small shared-memory exchange that is consumed one row later (one barrier per row in total).  Partial sums per
// (system, segment) are reduced in a fixed order by k_reduce_partials.
#include "common.cuh"
#include "romhc_internal.h"

#include <algorithm>

namespace romhc {

__device__ __forceinline__ void sw_cp_async16(void* smem, const void* g, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem)), "l"(g), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void sw_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void sw_cp_async_wait0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void sw_dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void sw_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

#define SW_THREADS 256
#define SW_RS 32            // grid rows per segment (one extra row is recomputed per segment: 3 %)

static inline int sw_pitch_a(int nk) { return ((4 * nk + 15) & ~15) + 4; }
template <int MT, int NTW>
static size_t sw_smem_bytes(int nk) {
    const int MS = 8 * MT, PP = 64 * NTW + 4;
    return (size_t(MS) * sw_pitch_a(nk) + 2 * 8 * MS + 8 * MS + size_t(2) * 4 * nk * PP) * 8;
}

template <int MT, int NTW>
__global__ void __launch_bounds__(SW_THREADS, (NTW <= 4) ? 2 : 1)
k_error_sweep(const double* __restrict__ U, const double* __restrict__ coef, const double* __restrict__ Phi, int n, int nk,
              int64_t K, int64_t Dp, int P, int R, int nseg, double* __restrict__ part) {
    constexpr int WCOLS = 64 * NTW, PP = WCOLS + 4, MS = 8 * MT;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) double sm[];
    const int PA = ((4 * nk + 15) & ~15) + 4;          // pitches = 4 mod 16 doubles: conflict-free fragment loads
    double* coefA = sm;                                 // MS x PA
    double* xch = coefA + MS * PA;                      // 2 x 8 warps x MS: first column of every warp's range
    double* red = xch + 2 * 8 * MS;                     // 8 warps x MS
    double* phis = red + 8 * MS;                        // 2 x (4 nk) x PP
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int seg = blockIdx.x;
    const int64_t sys0 = int64_t(blockIdx.y) * MS;
    const int r0 = seg * SW_RS, r1 = min(r0 + SW_RS, R);      // owned rows [r0, r1); row r1 is computed for its north edge only
    const int kk = 4 * nk;
    for (int i = tid; i < MS * kk; i += SW_THREADS) {
        const int s = i / kk, k = i - s * kk;
        coefA[s * PA + k] = (sys0 + s < K && k < n) ? coef[(sys0 + s) * n + k] : 0.0;
    }
    auto load_phi = [&](int buf, int r) {
        double* dst = phis + size_t(buf) * kk * PP;
        for (int v = tid; v < kk * (WCOLS / 2); v += SW_THREADS) {
            const int k = v / (WCOLS / 2), c = (v - k * (WCOLS / 2)) * 2;
            const bool ok = k < n && c < P && r < R;
            sw_cp_async16(dst + k * PP + c, ok ? Phi + int64_t(k) * Dp + int64_t(r) * P + c : Phi, ok ? 16 : 0);
        }
    };
    load_phi(0, r0);
    sw_cp_async_commit();
    double acc[MT][NTW][2], prev[MT][NTW][2], esum[MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) {
        esum[m] = 0.0;
#pragma unroll
        for (int j = 0; j < NTW; ++j) prev[m][j][0] = prev[m][j][1] = 0.0;
    }
    const int wcol0 = warp * 8 * NTW;
    for (int r = r0; r <= r1; ++r) {
        const int buf = (r - r0) & 1;
        // -u into the accumulators (issued before the wait so that the latency overlaps it)
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            const int64_t sys = sys0 + 8 * m + g;
#pragma unroll
            for (int j = 0; j < NTW; ++j) {
                const int col = wcol0 + 8 * j + 2 * t;
                double2 u = make_double2(0.0, 0.0);
                if (r < R && sys < K && col < P) u = *reinterpret_cast<const double2*>(U + sys * Dp + int64_t(r) * P + col);
                acc[m][j][0] = -u.x; acc[m][j][1] = -u.y;
            }
            if (t == 0 && r + 2 <= r1 && r + 2 < R && sys < K) {
#pragma unroll
                for (int j = 0; j < NTW; j += 2)
                    if (wcol0 + 8 * j < P) sw_prefetch_l2(U + sys * Dp + int64_t(r + 2) * P + wcol0 + 8 * j);
            }
        }
        sw_cp_async_wait0();
        __syncthreads();
        if (r < r1) load_phi(buf ^ 1, r + 1);
        sw_cp_async_commit();
        // east edge of the previous row at the warp boundary (its neighbour was published one row ago)
        if (r > r0 && t == 3) {
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                const double nb = (warp < 7) ? xch[(buf ^ 1) * 8 * MS + (warp + 1) * MS + 8 * m + g] : 0.0;
                const double d = prev[m][NTW - 1][1] - nb;
                esum[m] = fma(d, d, esum[m]);
            }
        }
        if (r < R) {
            const double* pa = coefA + g * PA + t;
            const double* pb = phis + size_t(buf) * kk * PP + t * PP + wcol0 + g;
            for (int ks = 0; ks < nk; ++ks) {
                double a[MT], b[NTW];
#pragma unroll
                for (int m = 0; m < MT; ++m) a[m] = pa[8 * m * PA + 4 * ks];
#pragma unroll
                for (int j = 0; j < NTW; ++j) b[j] = pb[4 * ks * PP + 8 * j];
#pragma unroll
                for (int m = 0; m < MT; ++m)
#pragma unroll
                    for (int j = 0; j < NTW; ++j) sw_dmma884(acc[m][j][0], acc[m][j][1], a[m], b[j]);
            }
        }
        if (r > r0) {                                   // south edges of row r - 1
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int j = 0; j < NTW; ++j) {
                    const double d0 = prev[m][j][0] - acc[m][j][0], d1 = prev[m][j][1] - acc[m][j][1];
                    esum[m] = fma(d0, d0, esum[m]);
                    esum[m] = fma(d1, d1, esum[m]);
                }
        }
        if (r < r1) {                                   // east edges of row r inside the warp's range
#pragma unroll
            for (int m = 0; m < MT; ++m) {
#pragma unroll
                for (int j = 0; j < NTW; ++j) {
                    const double d = acc[m][j][0] - acc[m][j][1];
                    esum[m] = fma(d, d, esum[m]);
                    const double s1 = __shfl_sync(FULL, acc[m][j][0], (lane + 1) & 31);
                    double nxt = s1;
                    if (j + 1 < NTW) {
                        const double s2 = __shfl_sync(FULL, acc[m][j + 1 < NTW ? j + 1 : j][0], (lane - 3) & 31);
                        if (t == 3) nxt = s2;
                    }
                    if (t < 3 || j + 1 < NTW) {
                        const double e = acc[m][j][1] - nxt;
                        esum[m] = fma(e, e, esum[m]);
                    }
                }
                if (t == 0) xch[buf * 8 * MS + warp * MS + 8 * m + g] = acc[m][0][0];
            }
        }
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int j = 0; j < NTW; ++j) { prev[m][j][0] = acc[m][j][0]; prev[m][j][1] = acc[m][j][1]; }
    }
#pragma unroll
    for (int m = 0; m < MT; ++m) {
        double s = esum[m];
        s += __shfl_xor_sync(FULL, s, 1);
        s += __shfl_xor_sync(FULL, s, 2);
        if (t == 0) red[warp * MS + 8 * m + g] = s;
    }
    __syncthreads();
    if (tid < MS && sys0 + tid < K) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w * MS + tid];
        part[(sys0 + tid) * nseg + seg] = s;
    }
}

// ---- variant 2: independent warps -------------------------------------------------------------------------------------------
// ncu on the kernel above: the per-row barrier phases the eight warps (all load, then all issue DMMAs): DMMA sub-pipe 39 %
// busy, stalls `wait` / `math_pipe_throttle` / `barrier`.  Here a warp owns its column band for the whole row segment: its
// own two-stage cp.async ring for the band's slab of the basis (band + one column), the east neighbour of the band's last
// column from that extra column by 5 FMAs per lane and a quad reduction instead of the shared exchange.  No block barrier
// inside the row loop; the warps drift apart and overlap each other's loads and DMMAs.
// Measured (tests/probe_sweep.py): 1.94 ms against 1.92 ms at K = 10 000, n = 20, 256^2 (1.31 / 1.39 ms at n = 8), 16.6 against
// 12.0 ms at 512^2 -- the barrier was not the limit: a warp has its snapshot loads in flight only while it is not computing,
// so about half of the HBM latency stays exposed in both variants.  Kept as an option, not the default.
template <int MT, int NTW>
static size_t sw2_smem_bytes(int nk) {
    const int MS = 8 * MT, PPw = 8 * NTW + 4;
    return (size_t(MS) * sw_pitch_a(nk) + 8 * MS + size_t(8) * 2 * 4 * nk * PPw) * 8;
}
__device__ __forceinline__ void sw_cp_async_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

template <int MT, int NTW>
__global__ void __launch_bounds__(SW_THREADS, (NTW <= 4) ? 2 : 1)
k_error_sweep2(const double* __restrict__ U, const double* __restrict__ coef, const double* __restrict__ Phi, int n, int nk,
               int64_t K, int64_t Dp, int P, int R, int nseg, double* __restrict__ part) {
    constexpr int WB = 8 * NTW, PPw = WB + 4, MS = 8 * MT;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) double sm[];
    const int PA = ((4 * nk + 15) & ~15) + 4;
    double* coefA = sm;                                 // MS x PA
    double* red = coefA + MS * PA;                      // 8 warps x MS
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int kk = 4 * nk;
    double* ring = red + 8 * MS + size_t(warp) * 2 * kk * PPw;   // this warp's two stages of kk x PPw
    const int seg = blockIdx.x;
    const int64_t sys0 = int64_t(blockIdx.y) * MS;
    const int r0 = seg * SW_RS, r1 = min(r0 + SW_RS, R);
    for (int i = tid; i < MS * kk; i += SW_THREADS) {
        const int s = i / kk, k = i - s * kk;
        coefA[s * PA + k] = (sys0 + s < K && k < n) ? coef[(sys0 + s) * n + k] : 0.0;
    }
    const int wcol0 = warp * WB;
    constexpr int CH = (WB + 2) / 2;                    // 16-byte chunks per basis row: the band and one more column (+ 1 unused)
    auto load_phi = [&](int buf, int r) {
        double* dst = ring + size_t(buf) * kk * PPw;
        for (int v = lane; v < kk * CH; v += 32) {
            const int k = v / CH, c = (v - k * CH) * 2;
            const bool ok = k < n && wcol0 + c < P && r < R;
            sw_cp_async16(dst + k * PPw + c, ok ? Phi + int64_t(k) * Dp + int64_t(r) * P + wcol0 + c : Phi, ok ? 16 : 0);
        }
    };
    load_phi(0, r0);
    sw_cp_async_commit();
    __syncthreads();                                    // coefA complete
    double acc[MT][NTW][2], prev[MT][NTW][2], esum[MT], vedge[MT], pedge[MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) {
        esum[m] = 0.0; pedge[m] = 0.0;
#pragma unroll
        for (int j = 0; j < NTW; ++j) prev[m][j][0] = prev[m][j][1] = 0.0;
    }
    const bool band_live = wcol0 < P;
    for (int r = r0; r <= r1 && band_live; ++r) {
        const int buf = (r - r0) & 1;
        if (r < r1) load_phi(buf ^ 1, r + 1);
        sw_cp_async_commit();
        double uedge[MT];
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            const int64_t sys = sys0 + 8 * m + g;
#pragma unroll
            for (int j = 0; j < NTW; ++j) {
                const int col = wcol0 + 8 * j + 2 * t;
                double2 u = make_double2(0.0, 0.0);
                if (r < R && sys < K && col < P) u = *reinterpret_cast<const double2*>(U + sys * Dp + int64_t(r) * P + col);
                acc[m][j][0] = -u.x; acc[m][j][1] = -u.y;
            }
            uedge[m] = (r < R && sys < K && wcol0 + WB < P) ? U[sys * Dp + int64_t(r) * P + wcol0 + WB] : 0.0;
            if (t == 0 && r + 2 <= r1 && r + 2 < R && sys < K) {
#pragma unroll
                for (int j = 0; j < NTW; j += 2)
                    if (wcol0 + 8 * j < P) sw_prefetch_l2(U + sys * Dp + int64_t(r + 2) * P + wcol0 + 8 * j);
            }
        }
        sw_cp_async_wait1();                            // this row's slab has landed (the next one may still be in flight)
        __syncwarp();
#pragma unroll
        for (int m = 0; m < MT; ++m) vedge[m] = 0.0;
        if (r < R) {
            const double* pa = coefA + g * PA + t;
            const double* pb = ring + size_t(buf) * kk * PPw + t * PPw + g;
            const double* pe = ring + size_t(buf) * kk * PPw + t * PPw + WB;
            for (int ks = 0; ks < nk; ++ks) {
                double a[MT], b[NTW];
#pragma unroll
                for (int m = 0; m < MT; ++m) a[m] = pa[8 * m * PA + 4 * ks];
#pragma unroll
                for (int j = 0; j < NTW; ++j) b[j] = pb[4 * ks * PPw + 8 * j];
                const double be = pe[4 * ks * PPw];
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    vedge[m] = fma(a[m], be, vedge[m]);
#pragma unroll
                    for (int j = 0; j < NTW; ++j) sw_dmma884(acc[m][j][0], acc[m][j][1], a[m], b[j]);
                }
            }
        }
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            double v = vedge[m];
            v += __shfl_xor_sync(FULL, v, 1);
            v += __shfl_xor_sync(FULL, v, 2);
            vedge[m] = v - uedge[m];                    // v at the first column of the next band (0 beyond the grid row)
        }
        if (r > r0) {                                   // south edges of row r - 1
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int j = 0; j < NTW; ++j) {
                    const double d0 = prev[m][j][0] - acc[m][j][0], d1 = prev[m][j][1] - acc[m][j][1];
                    esum[m] = fma(d0, d0, esum[m]);
                    esum[m] = fma(d1, d1, esum[m]);
                }
        }
        if (r < r1) {                                   // east edges of row r
#pragma unroll
            for (int m = 0; m < MT; ++m) {
#pragma unroll
                for (int j = 0; j < NTW; ++j) {
                    const double d = acc[m][j][0] - acc[m][j][1];
                    esum[m] = fma(d, d, esum[m]);
                    const double s1 = __shfl_sync(FULL, acc[m][j][0], (lane + 1) & 31);
                    double nxt = s1;
                    if (j + 1 < NTW) {
                        const double s2 = __shfl_sync(FULL, acc[m][j + 1 < NTW ? j + 1 : j][0], (lane - 3) & 31);
                        if (t == 3) nxt = s2;
                    } else if (t == 3) {
                        nxt = vedge[m];
                    }
                    const double e = acc[m][j][1] - nxt;
                    esum[m] = fma(e, e, esum[m]);
                }
            }
        }
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int j = 0; j < NTW; ++j) { prev[m][j][0] = acc[m][j][0]; prev[m][j][1] = acc[m][j][1]; }
        __syncwarp();                                   // every lane is done with stage buf before it is refilled
    }
    (void)pedge;
#pragma unroll
    for (int m = 0; m < MT; ++m) {
        double s = esum[m];
        s += __shfl_xor_sync(FULL, s, 1);
        s += __shfl_xor_sync(FULL, s, 2);
        if (t == 0) red[warp * MS + 8 * m + g] = s;
    }
    __syncthreads();
    if (tid < MS && sys0 + tid < K) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w * MS + tid];
        part[(sys0 + tid) * nseg + seg] = s;
    }
}

__global__ void k_sweep_reduce(const double* __restrict__ part, int np, double* __restrict__ out, int64_t K) {
    const int64_t k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (k >= K) return;
    double s = 0.0;
    for (int i = 0; i < np; ++i) s += part[k * np + i];
    out[k] = sqrt(s);
}

int g_sweep_variant = 1;     // 1: one barrier per row with the shared exchange (default), 2: independent warps (option "sweep" = 2)
template <int MT, int NTW>
static int launch_error_sweep(const LevelGeo& g, const double* U, const double* coef, const double* basis, int n, int64_t K,
                              double* part, int nseg, cudaStream_t st) {
    const int nk = (n + 3) / 4;
    if (g_sweep_variant == 2 && sw2_smem_bytes<MT, NTW>(nk) <= 227 * 1024) {
        const size_t smb2 = sw2_smem_bytes<MT, NTW>(nk);
        CK(cudaFuncSetAttribute(k_error_sweep2<MT, NTW>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smb2)));
        const int64_t ntiles2 = (K + 8 * MT - 1) / (8 * MT);
        for (int64_t t0 = 0; t0 < ntiles2; t0 += 65535) {
            const int nt = int(std::min<int64_t>(65535, ntiles2 - t0));
            const int64_t k0 = t0 * 8 * MT;
            ++g_launches;
            k_error_sweep2<MT, NTW><<<dim3(nseg, nt), SW_THREADS, smb2, st>>>(U + k0 * g.Dp, coef + k0 * n, basis, n, nk, K - k0, g.Dp,
                                                                             g.P, g.R, nseg, part + k0 * nseg);
        }
        CK(cudaGetLastError());
        return ROMHC_OK;
    }
    const size_t smb = sw_smem_bytes<MT, NTW>(nk);
    CK(cudaFuncSetAttribute(k_error_sweep<MT, NTW>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smb)));
    const int64_t ntiles = (K + 8 * MT - 1) / (8 * MT);
    for (int64_t t0 = 0; t0 < ntiles; t0 += 65535) {
        const int nt = int(std::min<int64_t>(65535, ntiles - t0));
        const int64_t k0 = t0 * 8 * MT;
        ++g_launches;
        k_error_sweep<MT, NTW><<<dim3(nseg, nt), SW_THREADS, smb, st>>>(U + k0 * g.Dp, coef + k0 * n, basis, n, nk, K - k0, g.Dp,
                                                                       g.P, g.R, nseg, part + k0 * nseg);
    }
    CK(cudaGetLastError());
    return ROMHC_OK;
}

// returns -1 when the configuration does not fit this kernel (the caller falls back to k_energy)
int Context::error_sweep(const double* U, const double* coef, const double* basis, int n, double* out, int64_t K, cudaStream_t st) {
    const LevelGeo& g = levels[0];
    if (n < 1 || g.P > 512 || (g.P & 1)) return -1;
    const int nk = (n + 3) / 4;
    const int nseg = (g.R + SW_RS - 1) / SW_RS;
    size_t smb;
    int variant;
    if (g.P <= 64) { variant = 0; smb = sw_smem_bytes<8, 1>(nk); }
    else if (g.P <= 128) { variant = 1; smb = sw_smem_bytes<4, 2>(nk); }
    else if (g.P <= 256) { variant = 2; smb = sw_smem_bytes<2, 4>(nk); }
    else { variant = 3; smb = sw_smem_bytes<1, 8>(nk); }
    if (smb > 227 * 1024) return -1;
    int rc = ensure_scratch(size_t(K) * nseg * 8); if (rc) return rc;
    double* part = (double*)scratch;
    switch (variant) {
        case 0: rc = launch_error_sweep<8, 1>(g, U, coef, basis, n, K, part, nseg, st); break;
        case 1: rc = launch_error_sweep<4, 2>(g, U, coef, basis, n, K, part, nseg, st); break;
        case 2: rc = launch_error_sweep<2, 4>(g, U, coef, basis, n, K, part, nseg, st); break;
        default: rc = launch_error_sweep<1, 8>(g, U, coef, basis, n, K, part, nseg, st); break;
    }
    if (rc) return rc;
    ++g_launches; k_sweep_reduce<<<(unsigned)((K + 127) / 128), 128, 0, st>>>(part, nseg, out, K);
    CK(cudaGetLastError());
    return ROMHC_OK;
}

}  // namespace romhc
