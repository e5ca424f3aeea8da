// Shared device/host helpers for the ROMHighContrast B200 hot path (sm_100a).
//
// Grid layout ("padded grid"): one system's P1 coefficient vector lives on the full vertex
// grid of its level, rows 0..R (R+1 rows) with row pitch P = roundup(C, 8) doubles, element
// (r, c) at r*P + c.  Interior DOFs are 1 <= r <= R-1, 1 <= c <= C-1 (reference numbering
// u[(r-1)*(C-1) + (c-1)], SolutionsManagers.py:153-163); every other slot is a stored zero, so
// the Dirichlet boundary needs no branches: (r, c-1) at c == 1 reads the zero column 0 and
// (r, c+1) at c == C-1 reads either a pad zero (P > C) or (r+1, 0) == 0 (P == C, wrap-around).
// Rows are 64-byte aligned, which is what the 1-D TMA bulk copies (cp.async.bulk) need.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define ROMHC_MAX_BLOCKS 256      // nrb * ncb
#define ROMHC_MAX_LEVELS 12
#define ROMHC_DIRECT_MAX 64       // coarsest level solved by dense Cholesky up to this many DOFs
#define ROMHC_TAIL_MAX_DP 4352    // levels with Dp <= this are smoothed nu_tail times (and may run inside the tail kernel)
#define ROMHC_DEEP_TAIL_MAX_DP 1100   // the tail kernel proper starts at the first level this small (if the coarsest is)

struct LevelGeo {
    int nrb, ncb;   // subdomain blocks (rows, cols)
    int N;          // cells per block per dimension at this level
    int R, C;       // cells per dimension (R = nrb*N rows, C = ncb*N columns)
    int P;          // row pitch in doubles
    int Dp;         // padded doubles per system = (R+1)*P
};

static inline LevelGeo make_level(int nrb, int ncb, int N) {
    LevelGeo g;
    g.nrb = nrb; g.ncb = ncb; g.N = N; g.R = nrb * N; g.C = ncb * N;
    g.P = (g.C + 7) / 8 * 8;
    g.Dp = (g.R + 1) * g.P;
    return g;
}

#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier + 1-D TMA bulk copy (SASS: UBLKCP / SYNCS) -------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// smem -> global bulk store (async proxy), bulk-group completion
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// make generic-proxy smem writes visible to the async proxy (before a bulk store reads them)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- operand strips: rows [row0, row0 + nrow) of one system's padded grid, staged in smem with row pitch P ----------
__device__ __forceinline__ uint32_t strip_tx_bytes(const LevelGeo& g, int row0, int nrow) {
    const int lo = max(row0, 0), hi = min(row0 + nrow, g.R + 1);
    return hi > lo ? uint32_t(hi - lo) * uint32_t(g.P) * 8u : 0u;
}
// one thread: issue the bulk load of the in-range rows (the caller has armed `bar` with the byte count)
__device__ __forceinline__ void strip_issue(double* sm, const double* gsys, const LevelGeo& g, int row0, int nrow,
                                            uint64_t* bar) {
    const int lo = max(row0, 0), hi = min(row0 + nrow, g.R + 1);
    if (hi > lo) bulk_g2s(sm + size_t(lo - row0) * g.P, gsys + size_t(lo) * g.P, uint32_t(hi - lo) * uint32_t(g.P) * 8u, bar);
}
// all threads: rows outside [0, R] behave as zeros
__device__ __forceinline__ void strip_zero_oob(double* sm, const LevelGeo& g, int row0, int nrow, int tid, int nt) {
    const int lo = min(max(row0, 0), row0 + nrow), hi = max(min(row0 + nrow, g.R + 1), lo);
    const int ntop = (lo - row0) * g.P;
    for (int i = tid; i < ntop; i += nt) sm[i] = 0.0;
    for (int i = (hi - row0) * g.P + tid; i < nrow * g.P; i += nt) sm[i] = 0.0;
}
// one thread: bulk store of rows [r_lo, r_hi) (clamped to the grid) from a strip whose first row is row0
__device__ __forceinline__ void strip_store(double* gsys, const double* sm, const LevelGeo& g, int row0, int r_lo,
                                            int r_hi) {
    const int lo = max(r_lo, 0), hi = min(r_hi, g.R + 1);
    if (hi > lo) bulk_s2g(gsys + size_t(lo) * g.P, sm + size_t(lo - row0) * g.P, uint32_t(hi - lo) * uint32_t(g.P) * 8u);
}

// Load rows [row_lo, row_hi) of one system's padded grid into smem (row pitch P) with ONE bulk copy
// for the in-range part [max(row_lo,0), min(row_hi,R+1)) and zero fill for rows outside the grid.
// Must be called by all threads; `bar` must have been initialised (count 1) and made visible by a
// __syncthreads().  Caller waits with mbar_wait(bar, parity).
__device__ __forceinline__ void strip_load_issue(double* sm, const double* gsys, const LevelGeo& g, int row_lo,
                                                 int row_hi, uint64_t* bar, int tid, int nthreads,
                                                 uint32_t extra_tx_bytes = 0) {
    const int lo = max(row_lo, 0), hi = min(row_hi, g.R + 1);
    if (tid == 0) {
        const uint32_t bytes = hi > lo ? uint32_t(hi - lo) * uint32_t(g.P) * 8u : 0u;
        mbar_expect_tx(bar, bytes + extra_tx_bytes);
        if (bytes) bulk_g2s(sm + size_t(lo - row_lo) * g.P, gsys + size_t(lo) * g.P, bytes, bar);
    }
    // rows outside [0, R] behave as zeros (they never overlap the bulk-copied rows)
    const int nz_top = (lo - row_lo) * g.P;
    for (int i = tid; i < nz_top; i += nthreads) sm[i] = 0.0;
    const int beg_bot = (max(hi, row_lo) - row_lo) * g.P, end_bot = (row_hi - row_lo) * g.P;
    for (int i = beg_bot + tid; i < end_bot; i += nthreads) sm[i] = 0.0;
}

// ---- reductions ------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Deterministic block sum; result valid in thread 0.  `red` = smem scratch of >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* red, int tid, int nthreads) {
    v = warp_sum(v);
    const int w = tid >> 5, l = tid & 31, nw = (nthreads + 31) >> 5;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    double s = 0.0;
    if (w == 0) {
        s = l < nw ? red[l] : 0.0;
        s = warp_sum(s);
    }
    return s;
}

// ---- stencil weights for one column while the row index advances monotonically ------------------------
// Stiffness row of vertex (r, c) (probe-verified against SolutionsManagers.py:187-215):
// edge to (r, c+1) has weight (k[r-1][c] + k[r][c]) / 2, edge to (r+1, c) weight (k[r][c-1] + k[r][c]) / 2,
// k = coefficient of cell (row, col) = a[row / N][col / N]; diagonal = sum of the four adjacent cells.
// The operator is always applied in DIFFERENCE form, sum_nb w_nb (u - u_nb): it stays accurate when the
// solution is nearly constant inside a 1e10 inclusion, where diag*u - sum w*u_nb cancels catastrophically.
struct ColW {
    const double* a;
    int ncb, N, bl, br, rb, rm;
    double wW, wE, wN, wS, dg, idg;
    __device__ __forceinline__ void compute() {
        const int bu = rb - (rm == 0), bd = rb;
        const double aul = a[bu * ncb + bl], aur = a[bu * ncb + br];
        const double adl = a[bd * ncb + bl], adr = a[bd * ncb + br];
        wW = 0.5 * (aul + adl);
        wE = 0.5 * (aur + adr);
        wN = 0.5 * (aul + aur);
        wS = 0.5 * (adl + adr);
        dg = (aul + aur) + (adl + adr);
        idg = 1.0 / dg;
    }
    // c in [1, C-1], r in [1, R-1]
    __device__ __forceinline__ void init(const double* a_, const LevelGeo& g, int c, int r) {
        a = a_; ncb = g.ncb; N = g.N;
        bl = (c - 1) / N; br = c / N;
        rb = r / N; rm = r - rb * N;
        compute();
    }
    __device__ __forceinline__ void advance(int step) {
        const bool was0 = (rm == 0);
        rm += step;
        bool wrapped = false;
        while (rm >= N) { rm -= N; ++rb; wrapped = true; }
        if (wrapped || was0) compute();
    }
};

// weights of an arbitrary interior vertex (slow path: integer divisions)
__device__ __forceinline__ void vertex_weights(const double* a, const LevelGeo& g, int r, int c, double& wW,
                                               double& wE, double& wN, double& wS) {
    const int bu = (r - 1) / g.N, bd = r / g.N, bl = (c - 1) / g.N, br = c / g.N;
    const double aul = a[bu * g.ncb + bl], aur = a[bu * g.ncb + br];
    const double adl = a[bd * g.ncb + bl], adr = a[bd * g.ncb + br];
    wW = 0.5 * (aul + adl);
    wE = 0.5 * (aur + adr);
    wN = 0.5 * (aul + aur);
    wS = 0.5 * (adl + adr);
}

// load one system's nrb*ncb coefficients into smem; y == nullptr means a == 1 (the H10 operator A_1)
__device__ __forceinline__ void load_coef(double* sa, const double* y, int64_t k, int nb, int tid, int nthreads) {
    for (int i = tid; i < nb; i += nthreads) sa[i] = y ? y[k * nb + i] : 1.0;
}

#endif  // __CUDACC__
