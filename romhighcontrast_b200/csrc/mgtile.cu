// Register-tiled multigrid strip kernels (the V-cycle's level-0/1 smoothers, restriction and prolongation).
//
// Same algorithm as k_mg_down / k_mg_up in solver.cu (red/black Gauss-Seidel V(nu,nu), P1 anti-diagonal transfers;
// preconditioner of the batched PCG that replaces galerkin(), /root/reference/src/lib/SolutionsManagers.py:17-40),
// different execution model: the shared-memory strip kernels spend ~120 instructions and ~15 shared-memory accesses
// per grid point (ncu, profiles/), which caps them at ~45 % of the HBM roofline.  Here every thread owns a 4 x 4
// tile of the strip in REGISTERS for the whole kernel (z and r: 32 doubles), loads it straight from global memory
// with 128-bit coalesced loads, and only the tile's perimeter travels through shared memory between the red and the
// black half sweeps.  The exchange buffers are laid out edge-by-edge (left/right columns: [row][column group],
// top/bottom rows: [row group][j][column group]) so that every warp access is dense -- no bank conflicts and
// ~1.7 shared-memory accesses per point update instead of ~5.5.
//
// Stencil weights come from a per-system table (k_weight_table): the P1 stiffness weights do not depend on the mesh
// width, so one table of (2 nrb - 1) x (2 ncb - 1) vertex classes ("inside block b" / "on the interface b-1 | b" per
// direction) serves every level; each CTA stages its system's table in shared memory with one TMA bulk copy.
// Rows that are not interface rows use w_N = w_S = diag / 4 exactly, so a thread keeps only {w_W, w_E, 1/diag} / diag
// of its four columns in registers; interface rows (one in N) take a general path that reads all weights from the table.
#include "common.cuh"
#include "romhc_internal.h"

#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <array>
#include <map>
#include <vector>

namespace romhc {

#define TWD 8              // doubles per weight-table entry: {wW, wE, wN, wS} / diag, 1 / diag, diag, 0, 0
#define TILE_MAXT 512      // largest CTA of the tile kernels: 4 warps per SM sub-partition -> 128 registers per thread

// ---- per-system weight table ----------------------------------------------------------------------------------------
// entry (rv, cv): rv even -> vertex row inside block row rv / 2, rv odd -> on the interface between block rows
// (rv - 1) / 2 and (rv + 1) / 2; same for cv.  Entry nrv * ncv is all zeros (boundary / padding columns).
__global__ void k_weight_table(const double* __restrict__ y, double* __restrict__ tab, int nrb, int ncb, int64_t K) {
    const int64_t k = blockIdx.x;
    if (k >= K) return;
    const int nrv = 2 * nrb - 1, ncv = 2 * ncb - 1, ne = nrv * ncv;
    const double* a = y ? y + k * int64_t(nrb) * ncb : nullptr;      // y == nullptr: a == 1 (the H10 operator A_1)
    for (int e = threadIdx.x; e <= ne; e += blockDim.x) {
        double* o = tab + (k * int64_t(ne + 1) + e) * TWD;
        if (e == ne) {
            for (int i = 0; i < TWD; ++i) o[i] = 0.0;
            continue;
        }
        const int rv = e / ncv, cv = e - rv * ncv;
        const int bd = (rv + 1) >> 1, bu = bd - (rv & 1);
        const int br = (cv + 1) >> 1, bl = br - (cv & 1);
        const double aul = a ? a[bu * ncb + bl] : 1.0, aur = a ? a[bu * ncb + br] : 1.0;
        const double adl = a ? a[bd * ncb + bl] : 1.0, adr = a ? a[bd * ncb + br] : 1.0;
        const double dg = (aul + aur) + (adl + adr);
        const double idg = 1.0 / dg;
        o[0] = 0.5 * (aul + adl) * idg;
        o[1] = 0.5 * (aur + adr) * idg;
        o[2] = 0.5 * (aul + aur) * idg;
        o[3] = 0.5 * (adl + adr) * idg;
        o[4] = idg;
        o[5] = dg;
        o[6] = 0.0;
        o[7] = 0.0;
    }
}

struct TileArgs {
    LevelGeo g, gc;
    const int* rowv;      // vertex class of row rho (valid for rho in [-ROWV_PAD, R + ROWV_PAD]); -1 = not an interior row
    const int* colv;      // vertex class of column c in [0, P); -1 = boundary / padding
    const double* tab;    // weight tables, ntab entries per system
    int ntab, ncv;
    int TY, ns, nu, has_coarse;
    int p_f32;            // k_mgp_update_down: the search direction p is stored as fp32 (k_pcg_p_apply wrote it that way)
    int in_f32;           // k_mgp_up: z_in (the going-down kernel's z_A) is fp32; going-down kernels: store z_A as fp32
    int out_f32;          // k_mgp_up: store z as fp32 (row pitch P floats, system pitch Dp floats) and form r.z from the rounded values
    int emit_res;         // persistent going-down kernels, has_coarse == 0: also store the residual r - A z of the owned rows
                          // on the red points ((row + col) even), packed with row pitch P / 2, for k_bridge_gather
    int NR, halo_top;     // rows of the CTA's region, rows above the owned strip
    int pf_dist;          // L2 prefetch distance in CTAs (= CTAs resident on the GPU at once); 0: off
    const int* rinfo;     // persistent kernels: per (strip, row group) row types (2 bits per row) | primary class << 8
    // k_mgp_update_down, deferred update of the iterate (fp32 search directions only): x is touched every SECOND
    // iteration.  0: x += alpha p now; 1: leave x alone (the direction stays in its buffer, alpha in its array);
    // 2: x += alpha_prev p_prev + alpha p -- the same two fused multiply-adds in the same order as two single updates, so x
    // is bit-identical; the round trip of x through HBM in between is what goes away (16 of 42 bytes per point and launch).
    // The mode itself is a template argument of the kernel (XMODE): the plain update compiles to the kernel as it was.
    const float* p_prev;
    const double* alpha_prev;
};

// shared memory (doubles): [T: ntab * TWD][red: 32][mbarrier: 2][EL: NR * (CG + 2)][ER: same][ET: (NRG + 2) * 4 * CG][EB: same]
static size_t tile_smem_bytes(int ntab, int NR, int CG) {
    const int NRG = NR / 4;
    return (size_t(ntab) * TWD + 34 + size_t(2) * NR * (CG + 2) + size_t(2) * (NRG + 2) * 4 * CG) * 8;
}

// 32-bit shared-window accesses: every address is one register plus a compile-time displacement
__device__ __forceinline__ double lds_f64(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ double2 lds_f64x2(uint32_t a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ int lds_s32(uint32_t a) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}

// Per-thread state.  CGT > 0: column groups per CTA known at compile time (P = 4 CGT), all displacements are immediates.
template <int CGT, bool CLS_SMEM = false>
struct TileThread {
    uint32_t xl, xr;      // own slots of row 0 of the tile in EL / ER   ([4 ty][tx + 1])
    uint32_t et, eb;      // own slots of column 0 of the tile in ET / EB ([(ty + 1) * 4][tx])
    uint32_t T;           // weight table
    int rt;               // 2 bits per tile row: 0 not an interior row, 1 fast (class == the tile's primary class), 2 general
    int cg;               // column groups (runtime copy)
    int rho0;
    int tx;               // column group of this thread
    // vertex class + 1 of the tile's rows / columns, 16 bits each (0: not interior): in registers, or (CLS_SMEM, the
    // persistent kernels) read from the row-info / column-class tables in shared memory when a general row needs them
    uint32_t rvp[CLS_SMEM ? 1 : 2], cep[CLS_SMEM ? 1 : 2];
    uint32_t ri_s, cv_s;  // CLS_SMEM: shared addresses of {rinfo, rvp0, rvp1} of this (strip, row group) and of the int4 column classes
    __device__ __forceinline__ int CG() const { return CGT > 0 ? CGT : cg; }
    __device__ __forceinline__ int CGp() const { return CG() + 2; }
};

// weights of the thread's four columns on the rows of the tile's primary class (w_N = w_S = 1/4 there)
template <bool NEED_DG>
struct ColWeights {
    double hW[4], hE[4], idg[4], dg[NEED_DG ? 4 : 1];
};

template <int CGT, bool CS>
__device__ __forceinline__ uint32_t tile_entry(const TileThread<CGT, CS>& t, const TileArgs& a, int I, int j) {
    int rv, cv;
    if (CS) {
        rv = int((uint32_t(lds_s32(t.ri_s + 4 + 4 * (I >> 1))) >> (16 * (I & 1))) & 0xffffu) - 1;
        cv = lds_s32(t.cv_s + 4 * j);
    } else {
        rv = int((t.rvp[CS ? 0 : (I >> 1)] >> (16 * (I & 1))) & 0xffffu) - 1;
        cv = int((t.cep[CS ? 0 : (j >> 1)] >> (16 * (j & 1))) & 0xffffu) - 1;
    }
    const int e = cv >= 0 ? rv * a.ncv + cv : a.ntab - 1;
    return t.T + uint32_t(e) * (TWD * 8);
}
template <int CGT>
__device__ __forceinline__ void tile_pack_classes(TileThread<CGT>& t, const int (&rv)[4], const int (&cv)[4]) {
    t.rvp[0] = uint32_t(rv[0] + 1) | (uint32_t(rv[1] + 1) << 16);
    t.rvp[1] = uint32_t(rv[2] + 1) | (uint32_t(rv[3] + 1) << 16);
    t.cep[0] = uint32_t(cv[0] + 1) | (uint32_t(cv[1] + 1) << 16);
    t.cep[1] = uint32_t(cv[2] + 1) | (uint32_t(cv[3] + 1) << 16);
}

// mbarrier + TMA bulk copy of the system's weight table, shared-memory carve-up, zeroed rim of the exchange buffers,
// row classification, primary-class weights
template <int CGT, bool NEED_DG>
__device__ __forceinline__ void tile_setup_thread(TileThread<CGT>& t, ColWeights<NEED_DG>& w, const TileArgs& a,
                                                  unsigned char* smem_raw, int64_t k, int rho0, uint32_t& red_addr) {
    const int CG = CGT > 0 ? CGT : int(blockDim.x), NRG = blockDim.y;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const uint32_t base = smem_u32(smem_raw);
    const uint32_t red = base + uint32_t(a.ntab) * (TWD * 8);
    const uint32_t bar = red + 32 * 8;
    const uint32_t EL = red + 34 * 8;
    const uint32_t ER = EL + uint32_t(a.NR) * (CG + 2) * 8;
    const uint32_t ET = ER + uint32_t(a.NR) * (CG + 2) * 8;
    const uint32_t EB = ET + uint32_t(NRG + 2) * 4 * CG * 8;
    red_addr = red;
    t.cg = CG; t.T = base; t.rho0 = rho0; t.tx = tx;
    if (tx == 0 && ty == 0) {
        uint64_t* b = reinterpret_cast<uint64_t*>(smem_raw + (bar - base));
        mbar_init(b, 1);
        mbar_fence_init();
        const uint32_t bytes = uint32_t(a.ntab) * TWD * 8u;
        mbar_expect_tx(b, bytes);
        bulk_g2s(smem_raw, a.tab + k * int64_t(a.ntab) * TWD, bytes, b);
    }
    t.xl = EL + uint32_t((4 * ty) * (CG + 2) + tx + 1) * 8;
    t.xr = ER + uint32_t((4 * ty) * (CG + 2) + tx + 1) * 8;
    t.et = ET + uint32_t((ty + 1) * 4 * CG + tx) * 8;
    t.eb = EB + uint32_t((ty + 1) * 4 * CG + tx) * 8;
    const int CGp = CG + 2;
    if (tx == 0) {
#pragma unroll
        for (int I = 0; I < 4; ++I) sts_f64(t.xr + (I * CGp - 1) * 8, 0.0);      // left of column group 0
    }
    if (tx == CG - 1) {
#pragma unroll
        for (int I = 0; I < 4; ++I) sts_f64(t.xl + (I * CGp + 1) * 8, 0.0);      // column P aliases (row + 1, 0) == 0
    }
    if (ty == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) sts_f64(t.eb + (j - 4) * CG * 8, 0.0);       // above the region
    }
    if (ty == NRG - 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) sts_f64(t.et + (j + 4) * CG * 8, 0.0);       // below the region
    }
    // row classes: the first non-interface interior row defines the primary class
    int rv[4], prim = -1;
#pragma unroll
    for (int I = 0; I < 4; ++I) {
        rv[I] = __ldg(a.rowv + rho0 + I);
        if (prim < 0 && rv[I] >= 0 && !(rv[I] & 1)) prim = rv[I];
    }
    t.rt = 0;
#pragma unroll
    for (int I = 0; I < 4; ++I) t.rt |= (rv[I] < 0 ? 0 : (rv[I] == prim ? 1 : 2)) << (2 * I);
    // the mbarrier must be initialised before anybody polls it; then wait for the table and fetch the
    // primary-class weights of the four columns
    __syncthreads();
    {
        uint64_t* b = reinterpret_cast<uint64_t*>(smem_raw + (bar - base));
        mbar_wait(b, 0);
    }
    const int4 cv = __ldg(reinterpret_cast<const int4*>(a.colv) + tx);
    const int cvs[4] = {cv.x, cv.y, cv.z, cv.w};
    tile_pack_classes(t, rv, cvs);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int e = (cvs[j] >= 0 && prim >= 0) ? prim * a.ncv + cvs[j] : a.ntab - 1;
        const uint32_t p = t.T + uint32_t(e) * (TWD * 8);
        const double2 u = lds_f64x2(p), v = lds_f64x2(p + 32);
        w.hW[j] = u.x; w.hE[j] = u.y; w.idg[j] = v.x;
        if (NEED_DG) w.dg[j] = v.y;
    }
}

// publish the perimeter values of colour X (0 red: (row + col) even, 1 black) of the tile
template <int X, int CGT, bool CS>
__device__ __forceinline__ void tile_publish(const double (&z)[4][4], const TileThread<CGT, CS>& t) {
    const int CG = t.CG(), CGp = t.CGp();
#pragma unroll
    for (int I = 0; I < 4; ++I) {
        if (((I + X) & 1) == 0) sts_f64(t.xl + I * CGp * 8, z[I][0]);
        else                    sts_f64(t.xr + I * CGp * 8, z[I][3]);
    }
    sts_f64(t.et + (X) * CG * 8, z[0][X]);
    sts_f64(t.et + (X + 2) * CG * 8, z[0][X + 2]);
    sts_f64(t.eb + (1 - X) * CG * 8, z[3][1 - X]);
    sts_f64(t.eb + (3 - X) * CG * 8, z[3][3 - X]);
}

// One half sweep over the tile's points of colour X.  Rows in the halo are relaxed like all others: what they
// produce beyond the validity cone (one row per half sweep) is never used.
// MODE 0: Gauss-Seidel update z = (r + sum_nb w z_nb) / diag
// MODE 1: residual d = r - diag z + sum_nb w z_nb, stored in place of z
// MODE 2: first half sweep from a zero initial guess, z = r / diag (no neighbours)
template <int X, int MODE, bool NEED_DG, int CGT, bool CS>
__device__ __forceinline__ void tile_phase(double (&z)[4][4], const double (&r)[4][4], const ColWeights<NEED_DG>& w,
                                           const TileThread<CGT, CS>& t, const TileArgs& a) {
    const int CG = t.CG(), CGp = t.CGp();
#pragma unroll
    for (int I = 0; I < 4; ++I) {
        const int jX = (I + X) & 1;
        const int rtype = (t.rt >> (2 * I)) & 3;
        if (rtype == 0) continue;
        double zWh = 0.0, zEh = 0.0, zN[2] = {0.0, 0.0}, zS[2] = {0.0, 0.0};
        if (MODE != 2) {
            if (jX == 0) zWh = lds_f64(t.xr + (I * CGp - 1) * 8);
            else         zEh = lds_f64(t.xl + (I * CGp + 1) * 8);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int j = jX + 2 * q;
                zN[q] = (I > 0) ? z[I > 0 ? I - 1 : 0][j] : lds_f64(t.eb + (j - 4) * CG * 8);
                zS[q] = (I < 3) ? z[I < 3 ? I + 1 : 3][j] : lds_f64(t.et + (j + 4) * CG * 8);
            }
        }
        if (rtype == 1) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int j = jX + 2 * q;
                if (MODE == 2) { z[I][j] = r[I][j] * w.idg[j]; continue; }
                const double zW = (j > 0) ? z[I][j > 0 ? j - 1 : 0] : zWh;
                const double zE = (j < 3) ? z[I][j < 3 ? j + 1 : 3] : zEh;
                const double upd = fma(r[I][j], w.idg[j], fma(0.25, zN[q] + zS[q], fma(w.hW[j], zW, w.hE[j] * zE)));
                z[I][j] = (MODE == 1) ? w.dg[NEED_DG ? j : 0] * (upd - z[I][j]) : upd;
            }
        } else {
            // interface row (or a second block row inside the tile): all weights of every point from the table
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int j = jX + 2 * q;
                const uint32_t p = tile_entry(t, a, I, j);
                const double2 wv = lds_f64x2(p + 32);      // 1 / diag, diag
                if (MODE == 2) { z[I][j] = r[I][j] * wv.x; continue; }
                const double2 we = lds_f64x2(p), ns = lds_f64x2(p + 16);
                const double zW = (j > 0) ? z[I][j > 0 ? j - 1 : 0] : zWh;
                const double zE = (j < 3) ? z[I][j < 3 ? j + 1 : 3] : zEh;
                const double upd = fma(r[I][j], wv.x, (we.x * zW + we.y * zE) + (ns.x * zN[q] + ns.y * zS[q]));
                z[I][j] = (MODE == 1) ? wv.y * (upd - z[I][j]) : upd;
            }
        }
    }
}

// One CTA per SM cannot overlap its own loads with its own arithmetic, so every CTA asks the L2 (cp.async.bulk.prefetch.L2,
// fire and forget) for the operand rows of the CTA that will run on this SM slot next: CTA ids are scheduled in
// linear order, the one `pf_dist` ahead starts when this one retires.
__device__ __forceinline__ void bulk_prefetch_l2(const void* gptr, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tile_prefetch_rows(const double* base, const LevelGeo& g, int64_t k, int row_lo, int nrow) {
    const int lo = max(row_lo, 0), hi = min(row_lo + nrow, g.R + 1);
    if (hi > lo) bulk_prefetch_l2(base + k * g.Dp + size_t(lo) * g.P, uint32_t(hi - lo) * uint32_t(g.P) * 8u);
}

// a tile row is one 32-byte sector: 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256)
__device__ __forceinline__ void tile_load_row(double (&v)[4], const double* row) {
    asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(row));
}
__device__ __forceinline__ void tile_store_row(double* row, const double (&v)[4]) {
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(row), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}

// ---- going down: z = nu RB-GS sweeps from 0; r_coarse = P^T (r - A z) ----------------------------------------------------
template <int CGT>
__global__ void __launch_bounds__(TILE_MAXT, 1)
k_mgt_down(TileArgs a, const double* __restrict__ r_in, double* __restrict__ z_out, double* __restrict__ rc_out,
           const int* __restrict__ active) {
    const int64_t k = blockIdx.y;
    if (!active[k]) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int P = CGT > 0 ? 4 * CGT : a.g.P;
    const int R = a.g.R;
    const int y0 = blockIdx.x * a.TY;
    if (threadIdx.x == 0 && threadIdx.y == 1 && a.pf_dist > 0) {
        const int64_t n = k * gridDim.x + blockIdx.x + a.pf_dist;
        const int64_t kn = n / gridDim.x;
        if (kn < gridDim.y && active[kn])
            tile_prefetch_rows(r_in, a.g, kn, int(n - kn * gridDim.x) * a.TY - a.halo_top, a.NR);
    }
    const int rho0 = y0 - a.halo_top + 4 * int(threadIdx.y), c0 = 4 * int(threadIdx.x);
    double z[4][4], r[4][4];
    const double* rp = r_in + k * a.g.Dp + int64_t(rho0) * P + c0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { z[i][j] = 0.0; r[i][j] = 0.0; }
        if (unsigned(rho0 + i) <= unsigned(R)) tile_load_row(r[i], rp + i * P);
    }
    TileThread<CGT> t;
    ColWeights<true> w;
    uint32_t red;
    tile_setup_thread(t, w, a, smem_raw, k, rho0, red);
    const int nu = a.nu;
    tile_phase<0, 2, true>(z, r, w, t, a);
    tile_publish<0>(z, t);
    __syncthreads();
    tile_phase<1, 0, true>(z, r, w, t, a);
    tile_publish<1>(z, t);
    __syncthreads();
    for (int sw = 1; sw < nu; ++sw) {
        tile_phase<0, 0, true>(z, r, w, t, a);
        tile_publish<0>(z, t);
        __syncthreads();
        tile_phase<1, 0, true>(z, r, w, t, a);
        tile_publish<1>(z, t);
        __syncthreads();
    }
    double* zo = z_out + k * a.g.Dp + int64_t(rho0) * P + c0;
    const int own_lo = y0 - rho0, own_hi = min(y0 + a.TY, R + 1) - rho0;     // owned tile rows: own_lo <= i < own_hi
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (i >= own_lo && i < own_hi) tile_store_row(zo + i * P, z[i]);
    if (!a.has_coarse) return;
    // residual on the red points (it vanishes on the just-relaxed black points), in place of z
    tile_phase<0, 1, true>(z, r, w, t, a);
    tile_publish<0>(z, t);
    __syncthreads();
    // r_c(I, J) = d(2I, 2J) + (d(2I-1, 2J+1) + d(2I+1, 2J-1)) / 2: tile-local (0,0), (0,2), (2,0), (2,2)
    const LevelGeo& gc = a.gc;
    const int CG = t.CG(), CGp = t.CGp();
    const double dN31 = lds_f64(t.eb - 3 * CG * 8), dN33 = lds_f64(t.eb - 1 * CG * 8);
    const double dW13 = lds_f64(t.xr + (1 * CGp - 1) * 8), dW33 = lds_f64(t.xr + (3 * CGp - 1) * 8);
    double rc[2][2];
    rc[0][0] = z[0][0] + 0.5 * (dN31 + dW13);
    rc[0][1] = z[0][2] + 0.5 * (dN33 + z[1][1]);
    rc[1][0] = z[2][0] + 0.5 * (z[1][1] + dW33);
    rc[1][1] = z[2][2] + 0.5 * (z[1][3] + z[3][1]);
    const int J0 = 2 * int(threadIdx.x);
    const bool okJ0 = J0 >= 1 && J0 <= gc.C - 1, okJ1 = J0 + 1 <= gc.C - 1;
    double* co = rc_out + k * gc.Dp + J0;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int rho = rho0 + 2 * q;
        const int I = rho >> 1;
        if (rho >= y0 && rho < y0 + a.TY && I <= gc.R) {
            const bool okI = I >= 1 && I <= gc.R - 1;
            *reinterpret_cast<double2*>(co + size_t(I) * gc.P) =
                make_double2(okI && okJ0 ? rc[q][0] : 0.0, okI && okJ1 ? rc[q][1] : 0.0);
        }
    }
}

// ---- going up: z += P e, nu BR-GS sweeps; optional r.z partials ----------------------------------------------------------
template <int CGT>
__global__ void __launch_bounds__(TILE_MAXT, 1)
k_mgt_up(TileArgs a, const double* __restrict__ e_c, const double* __restrict__ z_in, const double* __restrict__ r_in,
         double* __restrict__ z_out, const int* __restrict__ active, double* __restrict__ part_rz) {
    const int64_t k = blockIdx.y;
    if (!active[k]) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int P = CGT > 0 ? 4 * CGT : a.g.P;
    const int R = a.g.R;
    const int y0 = blockIdx.x * a.TY;
    if (threadIdx.x == 0 && threadIdx.y == 1 && a.pf_dist > 0) {
        const int64_t n = k * gridDim.x + blockIdx.x + a.pf_dist;
        const int64_t kn = n / gridDim.x;
        if (kn < gridDim.y && active[kn]) {
            const int yn = int(n - kn * gridDim.x) * a.TY - a.halo_top;
            tile_prefetch_rows(z_in, a.g, kn, yn, a.NR);
            tile_prefetch_rows(r_in, a.g, kn, yn + 1, a.NR - 2);
            if (a.has_coarse) tile_prefetch_rows(e_c, a.gc, kn, yn >> 1, a.NR / 2 + 1);
        }
    }
    const int rho0 = y0 - a.halo_top + 4 * int(threadIdx.y), c0 = 4 * int(threadIdx.x);
    double z[4][4], r[4][4];
    const double* zp = z_in + k * a.g.Dp + int64_t(rho0) * P + c0;
    const double* rp = r_in + k * a.g.Dp + int64_t(rho0) * P + c0;
    const int NR = a.NR;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int lr = 4 * int(threadIdx.y) + i;
#pragma unroll
        for (int j = 0; j < 4; ++j) { z[i][j] = 0.0; r[i][j] = 0.0; }
        if (unsigned(rho0 + i) <= unsigned(R)) {
            tile_load_row(z[i], zp + i * P);
            if (lr >= 1 && lr <= NR - 2) tile_load_row(r[i], rp + i * P);     // the outermost rows are never valid
        }
    }
    double e[3][3];
    if (a.has_coarse) {
        const LevelGeo& gc = a.gc;
        const int I0 = rho0 >> 1, J0 = 2 * int(threadIdx.x);
        const double* es = e_c + k * gc.Dp + int64_t(I0) * gc.P + J0;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            e[q][0] = e[q][1] = e[q][2] = 0.0;
            if (unsigned(I0 + q) <= unsigned(gc.R)) {
                const double2 v = *reinterpret_cast<const double2*>(es + q * gc.P);
                e[q][0] = v.x; e[q][1] = v.y;
                if (J0 + 2 < gc.P) e[q][2] = es[q * gc.P + 2];
            }
        }
    }
    TileThread<CGT> t;
    ColWeights<false> w;
    uint32_t red;
    tile_setup_thread(t, w, a, smem_raw, k, rho0, red);
    if (a.has_coarse) {
        // prolongation on the red points: (even, even) copies the coarse vertex, (odd, odd) is the midpoint of the
        // coarse cell's anti-diagonal (I, J+1)-(I+1, J).  Black values are overwritten by the first half sweep.
        z[0][0] += e[0][0]; z[0][2] += e[0][1];
        z[2][0] += e[1][0]; z[2][2] += e[1][1];
        z[1][1] += 0.5 * (e[0][1] + e[1][0]); z[1][3] += 0.5 * (e[0][2] + e[1][1]);
        z[3][1] += 0.5 * (e[1][1] + e[2][0]); z[3][3] += 0.5 * (e[1][2] + e[2][1]);
    }
    tile_publish<0>(z, t);
    __syncthreads();
    const int nu = a.nu;
    for (int sw = 0; sw < nu; ++sw) {
        tile_phase<1, 0, false>(z, r, w, t, a);
        tile_publish<1>(z, t);
        __syncthreads();
        tile_phase<0, 0, false>(z, r, w, t, a);
        if (sw + 1 < nu) {
            tile_publish<0>(z, t);
            __syncthreads();
        }
    }
    double* zo = z_out + k * a.g.Dp + int64_t(rho0) * P + c0;
    const int own_lo = y0 - rho0, own_hi = min(y0 + a.TY, R + 1) - rho0;
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (i >= own_lo && i < own_hi) {
            tile_store_row(zo + i * P, z[i]);
#pragma unroll
            for (int j = 0; j < 4; ++j) acc = fma(r[i][j], z[i][j], acc);
        }
    }
    if (part_rz) {
        const int tid = threadIdx.y * blockDim.x + threadIdx.x, nt = blockDim.x * blockDim.y;
        double* redp = reinterpret_cast<double*>(smem_raw + (red - smem_u32(smem_raw)));
        const double tot = block_sum(acc, redp, tid, nt);
        if (tid == 0) part_rz[k * a.ns + blockIdx.x] = tot;
    }
}

// =====================================================================================================
// Persistent, TMA-pipelined variants.  One CTA per SM walks over the work items (system, strip) round robin.  While
// item i is relaxed in registers, the operand rows of item i+1 (z, r, the coarse correction, the weight table) are
// already in flight: 1-D TMA bulk copies into a shared-memory staging area, signalled by one mbarrier.  At the top of
// an item every thread moves its tile from the staging area into registers, a __syncthreads() frees the area and the
// next item's copies are issued.  The register tile is what makes a single staging buffer enough.
// Tiles whose four rows are interior rows of one vertex class (all but the boundary / interface tiles; the flag is
// warp uniform) run straight-line code: eight independent updates per half sweep, no branches.
// Staging layout (doubles): [T0][T1][red 32][mbarrier 2][EL][ER][ET][EB][Rs: NR rows][Zs: NR rows][Es: NR/2+1 coarse rows]
// =====================================================================================================
struct TileStage {
    uint32_t T0, tb, red, bar, colv, rinfo, ex, Zs, Rs, Es;
};
// doubles reserved in front of the exchange buffers: two weight tables, 2 x 32 reduction slots, mbarrier, the
// column-class table (P ints) and the row-info table (ns * NRG ints)
static __host__ __device__ __forceinline__ uint32_t tile_stage_head(int ntab, int P, int nrinfo) {
    return uint32_t(2 * ntab * TWD + 64 + 2 + (P + 1) / 2 + (nrinfo + 1) / 2 + 1) & ~1u;
}
__device__ __forceinline__ TileStage tile_stage_carve(uint32_t base, const TileArgs& a, int CG, int NRG, bool with_z) {
    TileStage s;
    s.tb = uint32_t(a.ntab) * (TWD * 8);
    s.T0 = base;
    s.red = base + 2 * s.tb;
    s.bar = s.red + 64 * 8;
    s.colv = s.bar + 16;
    s.rinfo = s.colv + uint32_t((4 * CG + 1) / 2) * 8;
    s.ex = base + tile_stage_head(a.ntab, 4 * CG, 3 * a.ns * NRG) * 8;
    const uint32_t exb = (uint32_t(2) * a.NR * (CG + 2) + uint32_t(2) * (NRG + 2) * 4 * CG) * 8;
    const uint32_t strip = uint32_t(a.NR) * (4 * CG) * 8;
    s.Rs = s.ex + exb;
    s.Zs = s.Rs + strip;
    s.Es = s.Zs + (with_z ? strip : 0);
    return s;
}
static size_t tile_stage_bytes(int ntab, int NR, int CG, int ns, bool with_z, int e_rows, int Pc) {
    const int NRG = NR / 4;
    // + 32 bytes: neighbour reads at the very end of the last strip (east neighbour of the region's last point, coarse
    // vertex right of the last column) may touch one element past it; the values are never used, but the address has to
    // stay inside the CTA's shared-memory window
    return (size_t(tile_stage_head(ntab, 4 * CG, 3 * ns * NRG)) + size_t(2) * NR * (CG + 2) + size_t(2) * (NRG + 2) * 4 * CG +
            size_t(with_z ? 2 : 1) * NR * 4 * CG + size_t(e_rows) * Pc) * 8 + 32;
}
__device__ __forceinline__ int4 lds_s32x4(uint32_t a) {
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
// all threads: copy the column-class and row-info tables into shared memory (once per CTA)
__device__ __forceinline__ void tile_stage_tables(const TileStage& s, const TileArgs& a, int P, int nrinfo) {
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nt = blockDim.x * blockDim.y;
    for (int i = tid; i < P; i += nt)
        asm volatile("st.shared.s32 [%0], %1;" ::"r"(s.colv + i * 4), "r"(__ldg(a.colv + i)) : "memory");
    for (int i = tid; i < nrinfo; i += nt)
        asm volatile("st.shared.s32 [%0], %1;" ::"r"(s.rinfo + i * 4), "r"(__ldg(a.rinfo + i)) : "memory");
}

// one-time per-thread setup of the exchange-buffer addresses and the zero rim
template <int CGT>
__device__ __forceinline__ void tile_exchange_init(TileThread<CGT, true>& t, const TileArgs& a, uint32_t ex) {
    const int CG = CGT > 0 ? CGT : int(blockDim.x), NRG = blockDim.y;
    const int tx = threadIdx.x, ty = threadIdx.y, CGp = CG + 2;
    const uint32_t EL = ex;
    const uint32_t ER = EL + uint32_t(a.NR) * CGp * 8;
    const uint32_t ET = ER + uint32_t(a.NR) * CGp * 8;
    const uint32_t EB = ET + uint32_t(NRG + 2) * 4 * CG * 8;
    t.cg = CG; t.tx = tx;
    t.xl = EL + uint32_t((4 * ty) * CGp + tx + 1) * 8;
    t.xr = ER + uint32_t((4 * ty) * CGp + tx + 1) * 8;
    t.et = ET + uint32_t((ty + 1) * 4 * CG + tx) * 8;
    t.eb = EB + uint32_t((ty + 1) * 4 * CG + tx) * 8;
    if (tx == 0) {
#pragma unroll
        for (int I = 0; I < 4; ++I) sts_f64(t.xr + (I * CGp - 1) * 8, 0.0);
    }
    if (tx == CG - 1) {
#pragma unroll
        for (int I = 0; I < 4; ++I) sts_f64(t.xl + (I * CGp + 1) * 8, 0.0);
    }
    if (ty == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) sts_f64(t.eb + (j - 4) * CG * 8, 0.0);
    }
    if (ty == NRG - 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) sts_f64(t.et + (j + 4) * CG * 8, 0.0);
    }
}

// per-item: primary-class weights of the thread's four columns (table already in shared memory).
// rinfo = (row types, 2 bits per row) | primary class << 8, precomputed on the host per (strip, row group).
template <int CGT, bool NEED_DG>
__device__ __forceinline__ void tile_item_init(TileThread<CGT, true>& t, ColWeights<NEED_DG>& w, const TileArgs& a, uint32_t T,
                                               int rho0, int rinfo, uint32_t ri_s, uint32_t colv_s) {
    t.T = T; t.rho0 = rho0;
    t.rt = rinfo & 0xff;
    t.ri_s = ri_s; t.cv_s = colv_s + threadIdx.x * 16;
    const int prim = rinfo >> 8;          // -1: no fast row in this tile
    const int4 cv = lds_s32x4(t.cv_s);
    const int cvs[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int e = (cvs[j] >= 0 && prim >= 0) ? prim * a.ncv + cvs[j] : a.ntab - 1;
        const uint32_t p = T + uint32_t(e) * (TWD * 8);
        const double2 u = lds_f64x2(p), v = lds_f64x2(p + 32);
        w.hW[j] = u.x; w.hE[j] = u.y; w.idg[j] = v.x;
        if (NEED_DG) w.dg[j] = v.y;
    }
}

// Thread tx reads the 32 bytes at base + 32 tx with two 128-bit loads.  The hardware serves a 128-bit shared load per
// quarter warp: its eight 16-byte pieces must fall into eight different 16-byte bank groups, which a stride of 32 bytes
// does not give (lanes 0 and 4 collide, ...).  Lanes 4..7 of every quarter therefore fetch their two halves in the
// opposite order: conflict free (the timeline probe showed the plain version spending 2400 of ~10 000 cycles per item
// in this copy).
__device__ __forceinline__ void tile_lds_row(double (&v)[4], uint32_t addr) {
    const uint32_t h = (threadIdx.x & 4u) << 2;           // 0 or 16
    const double2 a = lds_f64x2(addr + h), b = lds_f64x2(addr + (h ^ 16u));
    v[0] = h ? b.x : a.x; v[1] = h ? b.y : a.y; v[2] = h ? a.x : b.x; v[3] = h ? a.y : b.y;
}

// fp32 rows (row pitch P floats): thread tx reads the 16 bytes at base + 16 tx -- consecutive, conflict free
__device__ __forceinline__ void tile_lds_row_f32(double (&v)[4], uint32_t addr) {
    float a, b, c, d;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "r"(addr) : "memory");
    v[0] = double(a); v[1] = double(b); v[2] = double(c); v[3] = double(d);
}
__device__ __forceinline__ double lds_f32_as_f64(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return double(v);
}
// store a row as fp32 and keep the ROUNDED values in the registers (what the reader of the row will see)
__device__ __forceinline__ void tile_store_row_f32(float* row, double (&v)[4]) {
    const float a = float(v[0]), b = float(v[1]), c = float(v[2]), d = float(v[3]);
    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(row), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
    v[0] = double(a); v[1] = double(b); v[2] = double(c); v[3] = double(d);
}

// straight-line half sweep for a tile whose four rows all have the primary class (t.rt == 0x55)
template <int X, int MODE, bool NEED_DG, int CGT, bool CS>
__device__ __forceinline__ void tile_phase_fast(double (&z)[4][4], const double (&r)[4][4], const ColWeights<NEED_DG>& w,
                                                const TileThread<CGT, CS>& t) {
    const int CG = t.CG(), CGp = t.CGp();
    if (MODE == 2) {
#pragma unroll
        for (int I = 0; I < 4; ++I)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int j = ((I + X) & 1) + 2 * q;
                z[I][j] = r[I][j] * w.idg[j];
            }
        return;
    }
    double hx[4], zN0[2], zS3[2];
#pragma unroll
    for (int I = 0; I < 4; ++I)
        hx[I] = (((I + X) & 1) == 0) ? lds_f64(t.xr + (I * CGp - 1) * 8) : lds_f64(t.xl + (I * CGp + 1) * 8);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        zN0[q] = lds_f64(t.eb + ((X & 1) + 2 * q - 4) * CG * 8);
        zS3[q] = lds_f64(t.et + (((3 + X) & 1) + 2 * q + 4) * CG * 8);
    }
    double upd[4][2];
#pragma unroll
    for (int I = 0; I < 4; ++I)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int j = ((I + X) & 1) + 2 * q;
            const double zW = (j > 0) ? z[I][j > 0 ? j - 1 : 0] : hx[I];
            const double zE = (j < 3) ? z[I][j < 3 ? j + 1 : 3] : hx[I];
            const double zN = (I > 0) ? z[I > 0 ? I - 1 : 0][j] : zN0[q];
            const double zS = (I < 3) ? z[I < 3 ? I + 1 : 3][j] : zS3[q];
            upd[I][q] = fma(r[I][j], w.idg[j], fma(0.25, zN + zS, fma(w.hW[j], zW, w.hE[j] * zE)));
        }
#pragma unroll
    for (int I = 0; I < 4; ++I)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int j = ((I + X) & 1) + 2 * q;
            z[I][j] = (MODE == 1) ? w.dg[NEED_DG ? j : 0] * (upd[I][q] - z[I][j]) : upd[I][q];
        }
}

template <int X, int MODE, bool NEED_DG, int CGT, bool CS>
__device__ __forceinline__ void tile_phase_any(double (&z)[4][4], const double (&r)[4][4], const ColWeights<NEED_DG>& w,
                                               const TileThread<CGT, CS>& t, const TileArgs& a) {
    if (t.rt == 0x55) tile_phase_fast<X, MODE, NEED_DG>(z, r, w, t);
    else              tile_phase<X, MODE, NEED_DG>(z, r, w, t, a);
}

// one thread: TMA bulk copy of rows [row0, row0 + nrow) (clamped to the grid) of one system into a staging strip
// whose first row is stage_row0; returns nothing, the byte count is computed by tile_rows_bytes
__device__ __forceinline__ uint32_t tile_rows_bytes(int row0, int nrow, int R, int P) {
    const int lo = max(row0, 0), hi = min(row0 + nrow, R + 1);
    return hi > lo ? uint32_t(hi - lo) * uint32_t(P) * 8u : 0u;
}
__device__ __forceinline__ void tile_tma(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tile_tma_rows(uint32_t dst, const double* gsys, int P, int R, int row0, int nrow,
                                              int stage_row0, uint32_t bar) {
    const int lo = max(row0, 0), hi = min(row0 + nrow, R + 1);
    if (hi > lo)
        tile_tma(dst + uint32_t(lo - stage_row0) * uint32_t(P) * 8u, gsys + size_t(lo) * P,
                 uint32_t(hi - lo) * uint32_t(P) * 8u, bar);
}

// round-robin walk over the (system, strip) items of the active systems
struct TileWalk {
    int k, strip, K, ns, dk, ds;
    const int* active;
    __device__ __forceinline__ void init(int K_, int ns_, const int* act) {
        K = K_; ns = ns_; active = act;
        dk = int(gridDim.x) / ns; ds = int(gridDim.x) - dk * ns;
        k = int(blockIdx.x) / ns; strip = int(blockIdx.x) - k * ns;
        skip();
    }
    __device__ __forceinline__ void skip() { while (k < K && !active[k]) step(); }
    __device__ __forceinline__ void step() {
        k += dk; strip += ds;
        if (strip >= ns) { strip -= ns; ++k; }
    }
    __device__ __forceinline__ void next() { step(); skip(); }
    __device__ __forceinline__ bool valid() const { return k < K; }
    // split form of next(): step() + flag() early (the load overlaps the item), settle() when the successor is needed
    __device__ __forceinline__ int flag() const { return k < K ? active[k] : 1; }
    __device__ __forceinline__ void settle(int f) { if (!f) { step(); skip(); } }
};
template <int CGT>
__global__ void __launch_bounds__(TILE_MAXT, 1)
k_mgp_down(TileArgs a, const double* __restrict__ r_in, double* __restrict__ z_out, double* __restrict__ rc_out,
           const int* __restrict__ active, int K) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int CG = CGT > 0 ? CGT : int(blockDim.x), NRG = blockDim.y;
    const int P = 4 * CG, R = a.g.R;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const bool leader = tx == 0 && ty == 0;
    const uint32_t base = smem_u32(smem_raw);
    const TileStage s = tile_stage_carve(base, a, CG, NRG, false);
    TileThread<CGT, true> t;
    tile_exchange_init(t, a, s.ex);
    tile_stage_tables(s, a, P, 3 * a.ns * NRG);
    if (leader) {
        mbar_init(reinterpret_cast<uint64_t*>(smem_raw + (s.bar - base)), 1);
        mbar_fence_init();
    }
    // the copies of one item are issued by different warps (part 0: expected byte count + weight table, part 1: r strip)
    auto issue = [&](const TileWalk& wk, int stage, int parts) {
        const int row0 = wk.strip * a.TY - a.halo_top;
        if (parts & 1) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s.bar),
                         "r"(s.tb + tile_rows_bytes(row0, a.NR, R, P)) : "memory");
            tile_tma(s.T0 + stage * s.tb, a.tab + int64_t(wk.k) * a.ntab * TWD, s.tb, s.bar);
        }
        if (parts & 2) tile_tma_rows(s.Rs, r_in + int64_t(wk.k) * a.g.Dp, P, R, row0, a.NR, row0, s.bar);
    };
    const int tid_ = ty * int(blockDim.x) + tx, nt_ = int(blockDim.x * blockDim.y);
    const int my_parts = (tid_ == 0 ? 1 : 0) | (tid_ == (nt_ > 32 ? 32 : 0) ? 2 : 0);
    __syncthreads();
    TileWalk wk;
    wk.init(K, a.ns, active);
    if (my_parts && wk.valid()) issue(wk, 0, my_parts);
    uint32_t phase = 0;
    int stage = 0;
    const int nu = a.nu;
    const int CGp = CG + 2;
    const uint32_t rs_own = s.Rs + uint32_t((4 * ty) * P + 4 * tx) * 8;
    while (wk.valid()) {
        const int k = wk.k, y0 = wk.strip * a.TY;
        const int rho0 = y0 - a.halo_top + 4 * ty;
        TileWalk nx = wk;
        nx.step();
        const int nflag = nx.flag();                       // in flight while this item is set up
        const uint32_t ri = s.rinfo + (wk.strip * NRG + ty) * 12;
        const int rinfo = lds_s32(ri);
        mbar_wait(reinterpret_cast<uint64_t*>(smem_raw + (s.bar - base)), phase);
        phase ^= 1u;
        double z[4][4], r[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) z[i][j] = 0.0;
        if ((rinfo & 0xff) == 0x55) {
#pragma unroll
            for (int i = 0; i < 4; ++i) tile_lds_row(r[i], rs_own + i * P * 8);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int j = 0; j < 4; ++j) r[i][j] = 0.0;
                if (unsigned(rho0 + i) <= unsigned(R)) tile_lds_row(r[i], rs_own + i * P * 8);
            }
        }
        ColWeights<true> w;
        tile_item_init(t, w, a, s.T0 + stage * s.tb, rho0, rinfo, ri, s.colv);
        __syncthreads();                                   // everybody has left the staging strip
        nx.settle(nflag);
        if (my_parts && nx.valid()) issue(nx, stage ^ 1, my_parts);
        tile_phase_any<0, 2, true>(z, r, w, t, a);
        tile_publish<0>(z, t);
        __syncthreads();
        tile_phase_any<1, 0, true>(z, r, w, t, a);
        tile_publish<1>(z, t);
        __syncthreads();
        for (int sw = 1; sw < nu; ++sw) {
            tile_phase_any<0, 0, true>(z, r, w, t, a);
            tile_publish<0>(z, t);
            __syncthreads();
            tile_phase_any<1, 0, true>(z, r, w, t, a);
            tile_publish<1>(z, t);
            __syncthreads();
        }
        double* zo = z_out + int64_t(k) * a.g.Dp + int64_t(rho0) * P + 4 * tx;
        const int own_lo = y0 - rho0, own_hi = min(y0 + a.TY, R + 1) - rho0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i >= own_lo && i < own_hi) {
                if (a.in_f32) tile_store_row_f32(reinterpret_cast<float*>(z_out) + (zo - z_out) + i * P, z[i]);
                else          tile_store_row(zo + i * P, z[i]);
            }
        if (a.has_coarse) {
            tile_phase_any<0, 1, true>(z, r, w, t, a);
            tile_publish<0>(z, t);
            __syncthreads();
            const LevelGeo& gc = a.gc;
            const double dN31 = lds_f64(t.eb - 3 * CG * 8), dN33 = lds_f64(t.eb - 1 * CG * 8);
            const double dW13 = lds_f64(t.xr + (1 * CGp - 1) * 8), dW33 = lds_f64(t.xr + (3 * CGp - 1) * 8);
            double rc[2][2];
            rc[0][0] = z[0][0] + 0.5 * (dN31 + dW13);
            rc[0][1] = z[0][2] + 0.5 * (dN33 + z[1][1]);
            rc[1][0] = z[2][0] + 0.5 * (z[1][1] + dW33);
            rc[1][1] = z[2][2] + 0.5 * (z[1][3] + z[3][1]);
            const int J0 = 2 * tx;
            const bool okJ0 = J0 >= 1 && J0 <= gc.C - 1, okJ1 = J0 + 1 <= gc.C - 1;
            double* co = rc_out + int64_t(k) * gc.Dp + J0;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int rho = rho0 + 2 * q;
                const int I = rho >> 1;
                if (rho >= y0 && rho < y0 + a.TY && I <= gc.R) {
                    const bool okI = I >= 1 && I <= gc.R - 1;
                    *reinterpret_cast<double2*>(co + size_t(I) * gc.P) =
                        make_double2(okI && okJ0 ? rc[q][0] : 0.0, okI && okJ1 ? rc[q][1] : 0.0);
                }
            }
        }
        else if (a.emit_res) {
            tile_phase_any<0, 1, true>(z, r, w, t, a);          // residual on the red points (black: zero, just relaxed)
            double* dout = rc_out + int64_t(k) * a.g.Dp + int64_t(rho0) * (P / 2) + 2 * tx;   // red points only, pitch P / 2
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (i >= own_lo && i < own_hi)
                    *reinterpret_cast<double2*>(dout + i * (P / 2)) = make_double2(z[i][i & 1], z[i][(i & 1) + 2]);
        }
        // no barrier here: the next item writes exchange slots only after its own staging barrier
        wk = nx;
        stage ^= 1;
    }
}

// ---- PCG update fused into the going-down kernel of the finest level --------------------------------------------------
// x += alpha p ; r -= alpha A p ; then exactly k_mgp_down on the new residual.  The residual never makes the round trip
// through HBM between the two steps (one stream of 7.25 saved, and one launch), at the price of recomputing A p on the
// halo rows of the region.  Neighbouring strips read the OLD residual on their halo rows, so the new one goes to a second
// buffer (the host swaps the two after the launch).  A p is evaluated in difference form, sum_nb w_nb (p - p_nb), like k_pcg_update; the
// neighbours outside the tile are read straight from the staged p strip (no exchange round needed).
// Region rows: r_new is valid on local rows [1, NR - 2], the smoother's validity cone starts from there.
template <bool GENERAL, int CGT>
__device__ __forceinline__ void tile_apply_row(double (&r)[4][4], const double (&p)[4][4], const ColWeights<true>& w,
                                               const TileThread<CGT, true>& t, const TileArgs& a, int I, double nalpha,
                                               double pWh, double pEh, const double (&pN)[4], const double (&pS)[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double pc = p[I][j];
        const double pW = (j > 0) ? p[I][j > 0 ? j - 1 : 0] : pWh;
        const double pE = (j < 3) ? p[I][j < 3 ? j + 1 : 3] : pEh;
        if (GENERAL) {
            const uint32_t q = tile_entry(t, a, I, j);
            const double2 we = lds_f64x2(q), ns = lds_f64x2(q + 16), wv = lds_f64x2(q + 32);
            const double s = (we.x * (pc - pW) + we.y * (pc - pE)) + (ns.x * (pc - pN[j]) + ns.y * (pc - pS[j]));
            r[I][j] = fma(nalpha * wv.y, s, r[I][j]);
        } else {
            const double s = fma(0.25, (pc - pN[j]) + (pc - pS[j]), fma(w.hW[j], pc - pW, w.hE[j] * (pc - pE)));
            r[I][j] = fma(nalpha * w.dg[j], s, r[I][j]);
        }
    }
}

template <int CGT, bool P32, int XMODE>
__global__ void __launch_bounds__(TILE_MAXT, 1)
k_mgp_update_down(TileArgs a, const double* __restrict__ p_in, double* __restrict__ x_io, const double* __restrict__ r_in,
                  double* __restrict__ r_out, const double* __restrict__ alpha, double* __restrict__ z_out, double* __restrict__ rc_out,
                  const int* __restrict__ active, int K) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int CG = CGT > 0 ? CGT : int(blockDim.x), NRG = blockDim.y;
    const int P = 4 * CG, R = a.g.R, NR = a.NR;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const bool leader = tx == 0 && ty == 0;
    const uint32_t base = smem_u32(smem_raw);
    const TileStage s = tile_stage_carve(base, a, CG, NRG, true);          // Zs holds the p strip
    TileThread<CGT, true> t;
    tile_exchange_init(t, a, s.ex);
    tile_stage_tables(s, a, P, 3 * a.ns * NRG);
    if (leader) {
        mbar_init(reinterpret_cast<uint64_t*>(smem_raw + (s.bar - base)), 1);
        mbar_fence_init();
    }
    // the copies of one item are issued by different warps (part 0: expected byte count, weight table, p strip;
    // part 1: r strip; part 2: L2 prefetch of the x rows, which are read straight from global memory)
    auto issue = [&](const TileWalk& wk, int stage, int parts) {
        const int row0 = wk.strip * a.TY - a.halo_top;
        if (parts & 1) {
            const uint32_t rb = tile_rows_bytes(row0, NR, R, P);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s.bar),
                         "r"(s.tb + rb + (P32 ? rb / 2 : rb)) : "memory");
            tile_tma(s.T0 + stage * s.tb, a.tab + int64_t(wk.k) * a.ntab * TWD, s.tb, s.bar);
            if (P32) {                                        // fp32 rows: P floats each, same element offsets
                const int lo = max(row0, 0), hi = min(row0 + NR, R + 1);
                if (hi > lo)
                    tile_tma(s.Zs + uint32_t(lo - row0) * uint32_t(P) * 4u,
                             reinterpret_cast<const float*>(p_in) + int64_t(wk.k) * a.g.Dp + size_t(lo) * P,
                             uint32_t(hi - lo) * uint32_t(P) * 4u, s.bar);
            } else {
                tile_tma_rows(s.Zs, p_in + int64_t(wk.k) * a.g.Dp, P, R, row0, NR, row0, s.bar);
            }
        }
        if (parts & 2) tile_tma_rows(s.Rs, r_in + int64_t(wk.k) * a.g.Dp, P, R, row0, NR, row0, s.bar);
        if ((parts & 4) && XMODE != 1) {
            tile_prefetch_rows(x_io, a.g, wk.k, wk.strip * a.TY, a.TY);
            if (XMODE == 2) {
                const int lo = max(wk.strip * a.TY, 0), hi = min(wk.strip * a.TY + a.TY, R + 1);
                if (hi > lo) bulk_prefetch_l2(a.p_prev + int64_t(wk.k) * a.g.Dp + size_t(lo) * P, uint32_t(hi - lo) * uint32_t(P) * 4u);
            }
        }
    };
    const int tid_ = ty * int(blockDim.x) + tx, nt_ = int(blockDim.x * blockDim.y);
    const int my_parts = (tid_ == 0 ? 1 : 0) | (tid_ == (nt_ > 32 ? 32 : 0) ? 2 : 0) | (tid_ == (nt_ > 64 ? 64 : 0) ? 4 : 0);
    __syncthreads();
    TileWalk wk;
    wk.init(K, a.ns, active);
    if (my_parts && wk.valid()) issue(wk, 0, my_parts);
    uint32_t phase = 0;
    int stage = 0;
    const int nu = a.nu;
    const int CGp = CG + 2;
    const uint32_t own = uint32_t((4 * ty) * P + 4 * tx) * 8;
    while (wk.valid()) {
        const int k = wk.k, y0 = wk.strip * a.TY;
        const int rho0 = y0 - a.halo_top + 4 * ty;
        TileWalk nx = wk;
        nx.step();
        const int nflag = nx.flag();                       // in flight while this item is set up
        const uint32_t ri = s.rinfo + (wk.strip * NRG + ty) * 12;
        const int rinfo = lds_s32(ri);
        const double al = __ldg(alpha + k);
        const int own_lo = y0 - rho0, own_hi = min(y0 + a.TY, R + 1) - rho0;   // owned tile rows: own_lo <= i < own_hi
        const int64_t goff = int64_t(k) * a.g.Dp + int64_t(rho0) * P + 4 * tx;
        double z[4][4], r[4][4];                                             // z holds p until the residual is updated
        constexpr int xmode = P32 ? XMODE : 0;       // compile time: the plain update (0) is the kernel as it was
        if (xmode != 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i)                                       // x rows: in flight during the staging wait
                if (i >= own_lo && i < own_hi) tile_load_row(r[i], x_io + goff + i * P);
        }
        if (xmode == 2) {
            // the deferred direction of the previous iteration goes in first (and is gone from the registers before the
            // staged strips are unpacked: the kernel runs at the 128-register cap)
            const double alp = __ldg(a.alpha_prev + k);
            const float* pq = a.p_prev + goff;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (i >= own_lo && i < own_hi) {
                    float p0, p1, p2, p3;
                    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(p0), "=f"(p1), "=f"(p2), "=f"(p3) : "l"(pq + i * P));
                    r[i][0] = fma(alp, double(p0), r[i][0]); r[i][1] = fma(alp, double(p1), r[i][1]);
                    r[i][2] = fma(alp, double(p2), r[i][2]); r[i][3] = fma(alp, double(p3), r[i][3]);
                }
        }
        mbar_wait(reinterpret_cast<uint64_t*>(smem_raw + (s.bar - base)), phase);
        phase ^= 1u;
        const bool all_rows = (rinfo & 0xff) == 0x55;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (all_rows || unsigned(rho0 + i) <= unsigned(R)) {
                if (P32) tile_lds_row_f32(z[i], s.Zs + own / 2 + i * P * 4);
                else         tile_lds_row(z[i], s.Zs + own + i * P * 8);
            } else { z[i][0] = z[i][1] = z[i][2] = z[i][3] = 0.0; }
        }
        // x += alpha p on the owned rows
        if (xmode != 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (i >= own_lo && i < own_hi) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) r[i][j] = fma(al, z[i][j], r[i][j]);
                    tile_store_row(x_io + goff + i * P, r[i]);
                }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (all_rows || unsigned(rho0 + i) <= unsigned(R)) tile_lds_row(r[i], s.Rs + own + i * P * 8);
            else { r[i][0] = r[i][1] = r[i][2] = r[i][3] = 0.0; }
        }
        ColWeights<true> w;
        tile_item_init(t, w, a, s.T0 + stage * s.tb, rho0, rinfo, ri, s.colv);
        // r -= alpha A p; neighbours of the tile from the staged p strip (rows outside the region: zero, those region
        // rows are outside the validity cone anyway)
        {
            const uint32_t pb = P32 ? s.Zs + own / 2 : s.Zs + own;   // fp32 strip: row pitch P * 4 bytes
            const uint32_t prow = P32 ? uint32_t(P) * 4u : uint32_t(P) * 8u;
            double pN[4], pS[4];
            if (ty > 0) { if (P32) tile_lds_row_f32(pN, pb - prow); else tile_lds_row(pN, pb - prow); }
            else { pN[0] = pN[1] = pN[2] = pN[3] = 0.0; }
#pragma unroll
            for (int I = 0; I < 4; ++I) {
                const int rtype = (t.rt >> (2 * I)) & 3;
                if (I == 3) {
                    if (ty < NRG - 1) { if (P32) tile_lds_row_f32(pS, pb + 4 * prow); else tile_lds_row(pS, pb + 4 * prow); }
                    else { pS[0] = pS[1] = pS[2] = pS[3] = 0.0; }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) pS[j] = z[I + 1 < 4 ? I + 1 : 3][j];
                }
                if (I > 0) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) pN[j] = z[I > 0 ? I - 1 : 0][j];
                }
                if (rtype != 0) {
                    const double pWh = P32 ? lds_f32_as_f64(pb + I * prow - 4) : lds_f64(pb + I * prow - 8);
                    const double pEh = P32 ? lds_f32_as_f64(pb + I * prow + 16) : lds_f64(pb + I * prow + 32);
                    if (rtype == 1) tile_apply_row<false>(r, z, w, t, a, I, -al, pWh, pEh, pN, pS);
                    else            tile_apply_row<true>(r, z, w, t, a, I, -al, pWh, pEh, pN, pS);
                }
            }
        }
        double* ro = r_out + goff;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i >= own_lo && i < own_hi) tile_store_row(ro + i * P, r[i]);
        __syncthreads();                                   // everybody has left the staging strips
        nx.settle(nflag);
        if (my_parts && nx.valid()) issue(nx, stage ^ 1, my_parts);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) z[i][j] = 0.0;
        tile_phase_any<0, 2, true>(z, r, w, t, a);
        tile_publish<0>(z, t);
        __syncthreads();
        tile_phase_any<1, 0, true>(z, r, w, t, a);
        tile_publish<1>(z, t);
        __syncthreads();
        for (int sw = 1; sw < nu; ++sw) {
            tile_phase_any<0, 0, true>(z, r, w, t, a);
            tile_publish<0>(z, t);
            __syncthreads();
            tile_phase_any<1, 0, true>(z, r, w, t, a);
            tile_publish<1>(z, t);
            __syncthreads();
        }
        double* zo = z_out + goff;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i >= own_lo && i < own_hi) {
                if (a.in_f32) tile_store_row_f32(reinterpret_cast<float*>(z_out) + (zo - z_out) + i * P, z[i]);
                else          tile_store_row(zo + i * P, z[i]);
            }
        if (a.has_coarse) {
            tile_phase_any<0, 1, true>(z, r, w, t, a);
            tile_publish<0>(z, t);
            __syncthreads();
            const LevelGeo& gc = a.gc;
            const double dN31 = lds_f64(t.eb - 3 * CG * 8), dN33 = lds_f64(t.eb - 1 * CG * 8);
            const double dW13 = lds_f64(t.xr + (1 * CGp - 1) * 8), dW33 = lds_f64(t.xr + (3 * CGp - 1) * 8);
            double rc[2][2];
            rc[0][0] = z[0][0] + 0.5 * (dN31 + dW13);
            rc[0][1] = z[0][2] + 0.5 * (dN33 + z[1][1]);
            rc[1][0] = z[2][0] + 0.5 * (z[1][1] + dW33);
            rc[1][1] = z[2][2] + 0.5 * (z[1][3] + z[3][1]);
            const int J0 = 2 * tx;
            const bool okJ0 = J0 >= 1 && J0 <= gc.C - 1, okJ1 = J0 + 1 <= gc.C - 1;
            double* co = rc_out + int64_t(k) * gc.Dp + J0;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int rho = rho0 + 2 * q;
                const int I = rho >> 1;
                if (rho >= y0 && rho < y0 + a.TY && I <= gc.R) {
                    const bool okI = I >= 1 && I <= gc.R - 1;
                    *reinterpret_cast<double2*>(co + size_t(I) * gc.P) =
                        make_double2(okI && okJ0 ? rc[q][0] : 0.0, okI && okJ1 ? rc[q][1] : 0.0);
                }
            }
        }
        else if (a.emit_res) {
            tile_phase_any<0, 1, true>(z, r, w, t, a);          // residual on the red points (black: zero, just relaxed)
            double* dout = rc_out + int64_t(k) * a.g.Dp + int64_t(rho0) * (P / 2) + 2 * tx;   // red points only, pitch P / 2
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (i >= own_lo && i < own_hi)
                    *reinterpret_cast<double2*>(dout + i * (P / 2)) = make_double2(z[i][i & 1], z[i][(i & 1) + 2]);
        }
        wk = nx;
        stage ^= 1;
    }
}

template <int CGT>
__global__ void __launch_bounds__(TILE_MAXT, 1)
k_mgp_up(TileArgs a, const double* __restrict__ e_c, const double* __restrict__ z_in, const double* __restrict__ r_in,
         double* __restrict__ z_out, const int* __restrict__ active, double* __restrict__ part_rz, int K) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int CG = CGT > 0 ? CGT : int(blockDim.x), NRG = blockDim.y;
    const int P = 4 * CG, R = a.g.R, NR = a.NR;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const bool leader = tx == 0 && ty == 0;
    const uint32_t base = smem_u32(smem_raw);
    const TileStage s = tile_stage_carve(base, a, CG, NRG, true);
    TileThread<CGT, true> t;
    tile_exchange_init(t, a, s.ex);
    tile_stage_tables(s, a, P, 3 * a.ns * NRG);
    if (leader) {
        mbar_init(reinterpret_cast<uint64_t*>(smem_raw + (s.bar - base)), 1);
        mbar_fence_init();
    }
    const int Pc = a.gc.P, Rc = a.gc.R;
    // the copies of one item are issued by different warps (part 0: expected byte count, weight table, z strip;
    // part 1: r strip; part 2: coarse correction strip)
    auto issue = [&](const TileWalk& wk, int stage, int parts) {
        const int row0 = wk.strip * a.TY - a.halo_top;
        if (parts & 1) {
            const uint32_t rb = tile_rows_bytes(row0, NR, R, P);
            uint32_t tot = s.tb + rb + (a.in_f32 ? rb / 2 : rb);
            if (a.has_coarse) tot += tile_rows_bytes(row0 >> 1, NR / 2 + 1, Rc, Pc);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s.bar), "r"(tot) : "memory");
            tile_tma(s.T0 + stage * s.tb, a.tab + int64_t(wk.k) * a.ntab * TWD, s.tb, s.bar);
        }
        if (parts & 8) {
            if (a.in_f32) {                                      // fp32 rows: P floats each, same element offsets
                const int lo = max(row0, 0), hi = min(row0 + NR, R + 1);
                if (hi > lo)
                    tile_tma(s.Zs + uint32_t(lo - row0) * uint32_t(P) * 4u,
                             reinterpret_cast<const float*>(z_in) + int64_t(wk.k) * a.g.Dp + size_t(lo) * P,
                             uint32_t(hi - lo) * uint32_t(P) * 4u, s.bar);
            } else {
                tile_tma_rows(s.Zs, z_in + int64_t(wk.k) * a.g.Dp, P, R, row0, NR, row0, s.bar);
            }
        }
        if (parts & 2) tile_tma_rows(s.Rs, r_in + int64_t(wk.k) * a.g.Dp, P, R, row0, NR, row0, s.bar);
        if ((parts & 4) && a.has_coarse)
            tile_tma_rows(s.Es, e_c + int64_t(wk.k) * a.gc.Dp, Pc, Rc, row0 >> 1, NR / 2 + 1, row0 >> 1, s.bar);
    };
    __syncthreads();
    TileWalk wk;
    wk.init(K, a.ns, active);
    const int tid = ty * int(blockDim.x) + tx, nwarps = (int(blockDim.x * blockDim.y) + 31) >> 5;
    const int nt_ = int(blockDim.x * blockDim.y);
    // one bulk copy per issuing warp: issuing a copy costs the thread several hundred cycles (timeline probe)
    const int my_parts = (tid == 0 ? 1 : 0) | (tid == (nt_ > 32 ? 32 : 0) ? 2 : 0) | (tid == (nt_ > 64 ? 64 : 0) ? 4 : 0) |
                         (tid == (nt_ > 96 ? 96 : 0) ? 8 : 0);
    if (my_parts && wk.valid()) issue(wk, 0, my_parts);
    uint32_t phase = 0;
    int stage = 0;
    const int nu = a.nu;
    const uint32_t own = uint32_t((4 * ty) * P + 4 * tx) * 8;
    const uint32_t es_own = s.Es + uint32_t((2 * ty) * Pc + 2 * tx) * 8;
    int64_t prev_slot = -1;                                // part_rz slot of the previous item (its reduction is deferred)
    while (wk.valid()) {
        const int k = wk.k, strip = wk.strip;
        const int y0 = strip * a.TY;
        const int rho0 = y0 - a.halo_top + 4 * ty;
        TileWalk nx = wk;
        nx.step();
        const int nflag = nx.flag();                       // in flight while this item is set up
        const uint32_t ri = s.rinfo + (strip * NRG + ty) * 12;
        const int rinfo = lds_s32(ri);
        mbar_wait(reinterpret_cast<uint64_t*>(smem_raw + (s.bar - base)), phase);
        phase ^= 1u;
        double z[4][4], r[4][4];
        if ((rinfo & 0xff) == 0x55) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (a.in_f32) tile_lds_row_f32(z[i], s.Zs + own / 2 + i * P * 4);
                else          tile_lds_row(z[i], s.Zs + own + i * P * 8);
                tile_lds_row(r[i], s.Rs + own + i * P * 8);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int j = 0; j < 4; ++j) { z[i][j] = 0.0; r[i][j] = 0.0; }
                if (unsigned(rho0 + i) <= unsigned(R)) {
                    if (a.in_f32) tile_lds_row_f32(z[i], s.Zs + own / 2 + i * P * 4);
                    else          tile_lds_row(z[i], s.Zs + own + i * P * 8);
                    tile_lds_row(r[i], s.Rs + own + i * P * 8);
                }
            }
        }
        if (a.has_coarse) {
            // prolongation on the red points (see k_mgt_up); the staging strip's coarse row 2 ty is row (rho0 >> 1)
            const int I0 = rho0 >> 1, J0 = 2 * tx;
            double e[3][3];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                e[q][0] = e[q][1] = e[q][2] = 0.0;
                if (unsigned(I0 + q) <= unsigned(Rc)) {
                    const uint32_t ea = es_own + q * Pc * 8;
                    const double2 v = lds_f64x2(ea);
                    e[q][0] = v.x; e[q][1] = v.y;
                    if (J0 + 2 < Pc) e[q][2] = lds_f64(ea + 16);
                }
            }
            z[0][0] += e[0][0]; z[0][2] += e[0][1];
            z[2][0] += e[1][0]; z[2][2] += e[1][1];
            z[1][1] += 0.5 * (e[0][1] + e[1][0]); z[1][3] += 0.5 * (e[0][2] + e[1][1]);
            z[3][1] += 0.5 * (e[1][1] + e[2][0]); z[3][3] += 0.5 * (e[1][2] + e[2][1]);
        }
        ColWeights<false> w;
        tile_item_init(t, w, a, s.T0 + stage * s.tb, rho0, rinfo, ri, s.colv);
        tile_publish<0>(z, t);
        __syncthreads();                                   // staging strip free, red perimeter visible
        nx.settle(nflag);
        if (my_parts && nx.valid()) issue(nx, stage ^ 1, my_parts);
        if (part_rz && prev_slot >= 0 && (tid >> 5) == nwarps - 1) {
            // deferred deterministic reduction of the previous item's per-warp r.z partials (other parity), on the
            // last warp: the first one is busy issuing the copies
            const int l = tid & 31;
            double v = l < nwarps ? lds_f64(s.red + ((stage ^ 1) * 32 + l) * 8) : 0.0;
            v = warp_sum(v);
            if (l == 0) part_rz[prev_slot] = v;
        }
        for (int sw = 0; sw < nu; ++sw) {
            tile_phase_any<1, 0, false>(z, r, w, t, a);
            tile_publish<1>(z, t);
            __syncthreads();
            tile_phase_any<0, 0, false>(z, r, w, t, a);
            if (sw + 1 < nu) {
                tile_publish<0>(z, t);
                __syncthreads();
            }
        }
        double* zo = z_out + int64_t(k) * a.g.Dp + int64_t(rho0) * P + 4 * tx;
        const int own_lo = y0 - rho0, own_hi = min(y0 + a.TY, R + 1) - rho0;
        double acc4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (i >= own_lo && i < own_hi) {
                if (a.out_f32) {
                    const float f0 = float(z[i][0]), f1 = float(z[i][1]), f2 = float(z[i][2]), f3 = float(z[i][3]);
                    float* zf = reinterpret_cast<float*>(z_out) + int64_t(k) * a.g.Dp + int64_t(rho0 + i) * P + 4 * tx;
                    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(zf), "f"(f0), "f"(f1), "f"(f2), "f"(f3) : "memory");
                    z[i][0] = double(f0); z[i][1] = double(f1); z[i][2] = double(f2); z[i][3] = double(f3);
                } else {
                    tile_store_row(zo + i * P, z[i]);
                }
                acc4[i] = fma(r[i][0], z[i][0], r[i][1] * z[i][1]) + fma(r[i][2], z[i][2], r[i][3] * z[i][3]);
            }
        }
        if (part_rz) {
            double acc = (acc4[0] + acc4[1]) + (acc4[2] + acc4[3]);
            acc = warp_sum(acc);
            if ((tid & 31) == 0) sts_f64(s.red + (stage * 32 + (tid >> 5)) * 8, acc);
            prev_slot = int64_t(k) * a.ns + strip;
        }
        // no barrier here: the next item's staging barrier orders the exchange-buffer and reduction-slot reuse
        wk = nx;
        stage ^= 1;
    }
    if (part_rz && prev_slot >= 0) {
        __syncthreads();
        if (tid < 32) {
            double v = tid < nwarps ? lds_f64(s.red + ((stage ^ 1) * 32 + tid) * 8) : 0.0;
            v = warp_sum(v);
            if (tid == 0) part_rz[prev_slot] = v;
        }
    }
}

// =====================================================================================================
// Multigrid tail on register tiles: every level with <= ROMHC_TAIL_MAX_DP padded vertices of ONE system, handled by one
// CTA.  The level being relaxed lives in registers (4 x 4 tile per thread, as above), the other levels wait in
// shared memory as plain padded arrays.  Replaces k_mg_tail (solver.cu), which spent ~90 instructions per point update
// in generic shared-memory loops.
// Shared memory (doubles): [T][red 32][mbarrier 2][exchange buffers of the finest tail level][z_l, r_l arrays][factor]
// =====================================================================================================
struct TailTileArgs {
    int nlev;
    LevelGeo geo[ROMHC_MAX_LEVELS];
    const int* rowv[ROMHC_MAX_LEVELS];
    const int* colv[ROMHC_MAX_LEVELS];
    int off_z[ROMHC_MAX_LEVELS], off_r[ROMHC_MAX_LEVELS];   // doubles from the start of shared memory; off_r[0] < 0 if unused
    int off_ex, off_fac, n_doubles;
    const double* tab;
    int ntab, ncv;
    int direct, DL, LD, coarse_sweeps, nu;
};

// per-level thread setup: tile coordinates, exchange addresses + zero rim, row types, primary-class weights
template <bool NEED_DG>
__device__ __forceinline__ bool tail_level_enter(TileThread<0>& t, ColWeights<NEED_DG>& w, TileArgs& a, const TailTileArgs& p,
                                                 int li, uint32_t base, int tid) {
    const LevelGeo& g = p.geo[li];
    const int CG = g.P / 4, NRG = (g.R + 3) / 4, CGp = CG + 2;
    const int ty = tid / CG, tx = tid - ty * CG;
    const bool act = ty < NRG;
    a.rowv = p.rowv[li]; a.colv = p.colv[li]; a.ncv = p.ncv; a.ntab = p.ntab;
    const int NR = 4 * NRG;
    const uint32_t EL = base + uint32_t(p.off_ex) * 8;
    const uint32_t ER = EL + uint32_t(NR) * CGp * 8;
    const uint32_t ET = ER + uint32_t(NR) * CGp * 8;
    const uint32_t EB = ET + uint32_t(NRG + 2) * 4 * CG * 8;
    t.cg = CG; t.tx = tx; t.T = base; t.rho0 = 4 * ty; t.rt = 0;
    t.xl = EL + uint32_t((4 * ty) * CGp + tx + 1) * 8;
    t.xr = ER + uint32_t((4 * ty) * CGp + tx + 1) * 8;
    t.et = ET + uint32_t((ty + 1) * 4 * CG + tx) * 8;
    t.eb = EB + uint32_t((ty + 1) * 4 * CG + tx) * 8;
    if (!act) return false;
    if (tx == 0) {
#pragma unroll
        for (int I = 0; I < 4; ++I) sts_f64(t.xr + (I * CGp - 1) * 8, 0.0);
    }
    if (tx == CG - 1) {
#pragma unroll
        for (int I = 0; I < 4; ++I) sts_f64(t.xl + (I * CGp + 1) * 8, 0.0);
    }
    if (ty == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) sts_f64(t.eb + (j - 4) * CG * 8, 0.0);
    }
    if (ty == NRG - 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) sts_f64(t.et + (j + 4) * CG * 8, 0.0);
    }
    int rv[4], prim = -1;
#pragma unroll
    for (int I = 0; I < 4; ++I) {
        rv[I] = __ldg(a.rowv + 4 * ty + I);
        if (prim < 0 && rv[I] >= 0 && !(rv[I] & 1)) prim = rv[I];
    }
#pragma unroll
    for (int I = 0; I < 4; ++I) t.rt |= (rv[I] < 0 ? 0 : (rv[I] == prim ? 1 : 2)) << (2 * I);
    const int4 cv = __ldg(reinterpret_cast<const int4*>(a.colv) + tx);
    const int cvs[4] = {cv.x, cv.y, cv.z, cv.w};
    tile_pack_classes(t, rv, cvs);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int e = (cvs[j] >= 0 && prim >= 0) ? prim * a.ncv + cvs[j] : a.ntab - 1;
        const uint32_t q = t.T + uint32_t(e) * (TWD * 8);
        const double2 u = lds_f64x2(q), v = lds_f64x2(q + 32);
        w.hW[j] = u.x; w.hE[j] = u.y; w.idg[j] = v.x;
        if (NEED_DG) w.dg[j] = v.y;
    }
    return true;
}

// tile <-> padded array in shared memory (rows 0..R only)
__device__ __forceinline__ void tail_tile_load(double (&v)[4][4], uint32_t arr, const LevelGeo& g, int rho0, int tx) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (rho0 + i <= g.R) tile_lds_row(v[i], arr + uint32_t((rho0 + i) * g.P + 4 * tx) * 8);
        else { v[i][0] = v[i][1] = v[i][2] = v[i][3] = 0.0; }
    }
}
__device__ __forceinline__ void tail_tile_store(uint32_t arr, const double (&v)[4][4], const LevelGeo& g, int rho0, int tx) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (rho0 + i <= g.R) {
            const uint32_t a = arr + uint32_t((rho0 + i) * g.P + 4 * tx) * 8;
            asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v[i][0]), "d"(v[i][1]) : "memory");
            asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a + 16), "d"(v[i][2]), "d"(v[i][3]) : "memory");
        }
}
__device__ __forceinline__ void tail_tile_load_global(double (&v)[4][4], const double* gsys, const LevelGeo& g, int rho0,
                                                      int tx) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (rho0 + i <= g.R) tile_load_row(v[i], gsys + size_t(rho0 + i) * g.P + 4 * tx);
        else { v[i][0] = v[i][1] = v[i][2] = v[i][3] = 0.0; }
    }
}

// MAXT = 128: the usual case (a 32 x 32 tail level is 64 threads) may use up to 168 registers (three CTAs of 128 threads
// would still fit; shared memory allows ~6 CTAs of 64 threads) -- under the 128-register cap of the 256-thread variant the
// per-level TileArgs and tile state spilled to local memory inside every phase
template <int MAXT>
__global__ void __launch_bounds__(MAXT, MAXT == 128 ? 3 : 2)
k_mgt_tail(TailTileArgs p, const double* __restrict__ r_in, double* __restrict__ z_out, const double* __restrict__ cfac,
           const int* __restrict__ active, double* __restrict__ part_rz) {
    const int64_t k = blockIdx.x;
    if (!active[k]) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, nt = blockDim.x;
    const uint32_t base = smem_u32(smem_raw);
    double* S = reinterpret_cast<double*>(smem_raw);
    double* redp = S + size_t(p.ntab) * TWD;
    uint64_t* bar = reinterpret_cast<uint64_t*>(redp + 32);
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
        const uint32_t bytes = uint32_t(p.ntab) * TWD * 8u;
        mbar_expect_tx(bar, bytes);
        bulk_g2s(S, p.tab + k * int64_t(p.ntab) * TWD, bytes, bar);
    }
    // level arrays start out as zeros (boundary and padding slots stay zero for ever)
    for (int i = p.off_ex + tid; i < p.off_fac; i += nt) S[i] = 0.0;
    if (p.direct) {
        const double* src = cfac + k * size_t(p.DL) * p.LD;
        double* F = S + p.off_fac;
        for (int i = tid; i < p.DL * p.LD; i += nt) F[i] = src[i];
    }
    __syncthreads();
    mbar_wait(bar, 0);
    const int last = p.nlev - 1;
    const double* rg = r_in + k * p.geo[0].Dp;
    TileThread<0> t;
    TileArgs a;
    double z[4][4], r[4][4];
    // ---- down ----
    for (int li = 0; li < last; ++li) {
        const LevelGeo& g = p.geo[li];
        const LevelGeo& gc = p.geo[li + 1];
        ColWeights<true> w;
        const bool act = tail_level_enter(t, w, a, p, li, base, tid);
        if (act) {
            if (li == 0) tail_tile_load_global(r, rg, g, t.rho0, t.tx);
            else         tail_tile_load(r, base + uint32_t(p.off_r[li]) * 8, g, t.rho0, t.tx);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) z[i][j] = 0.0;
        }
        for (int sw = 0; sw < p.nu; ++sw) {
            if (act) {
                if (sw == 0) tile_phase_any<0, 2, true>(z, r, w, t, a);
                else         tile_phase_any<0, 0, true>(z, r, w, t, a);
                tile_publish<0>(z, t);
            }
            __syncthreads();
            if (act) {
                tile_phase_any<1, 0, true>(z, r, w, t, a);
                tile_publish<1>(z, t);
            }
            __syncthreads();
        }
        if (act) {
            tail_tile_store(base + uint32_t(p.off_z[li]) * 8, z, g, t.rho0, t.tx);
            tile_phase_any<0, 1, true>(z, r, w, t, a);
            tile_publish<0>(z, t);
        }
        __syncthreads();
        if (act) {
            const int CG = t.cg, CGp = CG + 2;
            const double dN31 = lds_f64(t.eb - 3 * CG * 8), dN33 = lds_f64(t.eb - 1 * CG * 8);
            const double dW13 = lds_f64(t.xr + (1 * CGp - 1) * 8), dW33 = lds_f64(t.xr + (3 * CGp - 1) * 8);
            double rc[2][2];
            rc[0][0] = z[0][0] + 0.5 * (dN31 + dW13);
            rc[0][1] = z[0][2] + 0.5 * (dN33 + z[1][1]);
            rc[1][0] = z[2][0] + 0.5 * (z[1][1] + dW33);
            rc[1][1] = z[2][2] + 0.5 * (z[1][3] + z[3][1]);
            const int J0 = 2 * t.tx;
            const bool okJ0 = J0 >= 1 && J0 <= gc.C - 1, okJ1 = J0 + 1 <= gc.C - 1;
            double* co = S + p.off_r[li + 1] + J0;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int I = (t.rho0 >> 1) + q;
                if (I <= gc.R) {
                    const bool okI = I >= 1 && I <= gc.R - 1;
                    co[size_t(I) * gc.P] = okI && okJ0 ? rc[q][0] : 0.0;
                    co[size_t(I) * gc.P + 1] = okI && okJ1 ? rc[q][1] : 0.0;
                }
            }
        }
        __syncthreads();
    }
    // ---- coarsest ----
    {
        const LevelGeo& g = p.geo[last];
        double* rs = S + p.off_r[last];
        double* zs = S + p.off_z[last];
        if (last == 0) {
            for (int i = tid; i < g.Dp; i += nt) rs[i] = rg[i];
            __syncthreads();
        }
        if (p.direct) {
            // L L^T z = r with the packed factor (diagonal stores 1 / L_ii); warp 0, lanes own rows lane, lane+32
            if (tid < 32) {
                const int D = p.DL, LD = p.LD, W = g.C - 1;
                const double* F = S + p.off_fac;
                const int j0 = tid, j1 = tid + 32;
                double b0 = 0.0, b1 = 0.0;
                if (j0 < D) b0 = rs[(1 + j0 / W) * g.P + 1 + j0 % W];
                if (j1 < D) b1 = rs[(1 + j1 / W) * g.P + 1 + j1 % W];
                for (int i = 0; i < D; ++i) {
                    const double bi = __shfl_sync(0xffffffffu, i < 32 ? b0 : b1, i & 31);
                    const double yi = bi * F[i * LD + i];
                    if (j0 == i) b0 = yi;
                    if (j1 == i) b1 = yi;
                    if (j0 > i && j0 < D) b0 = fma(-F[j0 * LD + i], yi, b0);
                    if (j1 > i && j1 < D) b1 = fma(-F[j1 * LD + i], yi, b1);
                }
                for (int i = D - 1; i >= 0; --i) {
                    const double bi = __shfl_sync(0xffffffffu, i < 32 ? b0 : b1, i & 31);
                    const double xi = bi * F[i * LD + i];
                    if (j0 == i) b0 = xi;
                    if (j1 == i) b1 = xi;
                    if (j0 < i) b0 = fma(-F[i * LD + j0], xi, b0);
                    if (j1 < i) b1 = fma(-F[i * LD + j1], xi, b1);
                }
                if (j0 < D) zs[(1 + j0 / W) * g.P + 1 + j0 % W] = b0;
                if (j1 < D) zs[(1 + j1 / W) * g.P + 1 + j1 % W] = b1;
            }
            __syncthreads();
        } else {
            ColWeights<false> w;
            const bool act = tail_level_enter(t, w, a, p, last, base, tid);
            if (act) {
                tail_tile_load(r, base + uint32_t(p.off_r[last]) * 8, g, t.rho0, t.tx);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) z[i][j] = 0.0;
            }
            for (int sw = 0; sw < p.coarse_sweeps; ++sw) {
                if (act) {
                    if (sw == 0) tile_phase_any<0, 2, false>(z, r, w, t, a);
                    else         tile_phase_any<0, 0, false>(z, r, w, t, a);
                    tile_publish<0>(z, t);
                }
                __syncthreads();
                if (act) { tile_phase_any<1, 0, false>(z, r, w, t, a); tile_publish<1>(z, t); }
                __syncthreads();
            }
            for (int sw = 0; sw < p.coarse_sweeps; ++sw) {
                if (act) { tile_phase_any<1, 0, false>(z, r, w, t, a); tile_publish<1>(z, t); }
                __syncthreads();
                if (act) { tile_phase_any<0, 0, false>(z, r, w, t, a); tile_publish<0>(z, t); }
                __syncthreads();
            }
            if (act) tail_tile_store(base + uint32_t(p.off_z[last]) * 8, z, g, t.rho0, t.tx);
            __syncthreads();
        }
    }
    // ---- up ----
    double acc = 0.0;
    for (int li = last - 1; li >= 0; --li) {
        const LevelGeo& g = p.geo[li];
        const LevelGeo& gc = p.geo[li + 1];
        ColWeights<false> w;
        const bool act = tail_level_enter(t, w, a, p, li, base, tid);
        if (act) {
            tail_tile_load(z, base + uint32_t(p.off_z[li]) * 8, g, t.rho0, t.tx);
            if (li == 0) tail_tile_load_global(r, rg, g, t.rho0, t.tx);
            else         tail_tile_load(r, base + uint32_t(p.off_r[li]) * 8, g, t.rho0, t.tx);
            const int I0 = t.rho0 >> 1, J0 = 2 * t.tx;
            const uint32_t es = base + uint32_t(p.off_z[li + 1] + I0 * gc.P + J0) * 8;
            double e[3][3];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                e[q][0] = e[q][1] = e[q][2] = 0.0;
                if (I0 + q <= gc.R) {
                    const double2 v = lds_f64x2(es + q * gc.P * 8);
                    e[q][0] = v.x; e[q][1] = v.y;
                    if (J0 + 2 < gc.P) e[q][2] = lds_f64(es + q * gc.P * 8 + 16);
                }
            }
            z[0][0] += e[0][0]; z[0][2] += e[0][1];
            z[2][0] += e[1][0]; z[2][2] += e[1][1];
            z[1][1] += 0.5 * (e[0][1] + e[1][0]); z[1][3] += 0.5 * (e[0][2] + e[1][1]);
            z[3][1] += 0.5 * (e[1][1] + e[2][0]); z[3][3] += 0.5 * (e[1][2] + e[2][1]);
            tile_publish<0>(z, t);
        }
        __syncthreads();
        for (int sw = 0; sw < p.nu; ++sw) {
            if (act) { tile_phase_any<1, 0, false>(z, r, w, t, a); tile_publish<1>(z, t); }
            __syncthreads();
            if (act) {
                tile_phase_any<0, 0, false>(z, r, w, t, a);
                if (sw + 1 < p.nu) tile_publish<0>(z, t);
            }
            if (sw + 1 < p.nu) __syncthreads();
        }
        if (li > 0) {
            if (act) tail_tile_store(base + uint32_t(p.off_z[li]) * 8, z, g, t.rho0, t.tx);
            __syncthreads();
        } else if (act) {
            double* zo = z_out + k * g.Dp;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (t.rho0 + i <= g.R) {
                    tile_store_row(zo + size_t(t.rho0 + i) * g.P + 4 * t.tx, z[i]);
                    acc += fma(r[i][0], z[i][0], r[i][1] * z[i][1]) + fma(r[i][2], z[i][2], r[i][3] * z[i][3]);
                }
        }
    }
    if (last == 0) {
        // the whole hierarchy is the coarsest level: the solution sits in the shared-memory array
        const LevelGeo& g = p.geo[0];
        const double* zs = S + p.off_z[0];
        const double* rs = S + p.off_r[0];
        double* zo = z_out + k * g.Dp;
        for (int i = tid; i < g.Dp; i += nt) {
            const double v = zs[i];
            zo[i] = v;
            acc = fma(rs[i], v, acc);
        }
    }
    if (part_rz) {
        const double tot = block_sum(acc, redp, tid, nt);
        if (tid == 0) part_rz[k] = tot;
    }
}

// ======================================================================================================
// host side
// ======================================================================================================
int Context::tile_setup() {
    if (tile_ready) return ROMHC_OK;
    tile_maxt_down = tile_maxt_up = TILE_MAXT;
    const void* fns[] = {(const void*)k_mgt_down<64>, (const void*)k_mgt_down<32>, (const void*)k_mgt_down<0>,
                         (const void*)k_mgt_up<64>,   (const void*)k_mgt_up<32>,   (const void*)k_mgt_up<0>};
    for (int i = 0; i < 6; ++i) {
        CK(cudaFuncSetAttribute(fns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        cudaFuncAttributes fa;
        CK(cudaFuncGetAttributes(&fa, fns[i]));
        int& m = i < 3 ? tile_maxt_down : tile_maxt_up;
        m = std::min(m, fa.maxThreadsPerBlock);
    }
    const void* ffns[] = {(const void*)k_mgp_update_down<64, false, 0>, (const void*)k_mgp_update_down<32, false, 0>,
                          (const void*)k_mgp_update_down<16, false, 0>, (const void*)k_mgp_update_down<0, false, 0>,
                          (const void*)k_mgp_update_down<64, true, 0>,  (const void*)k_mgp_update_down<32, true, 0>,
                          (const void*)k_mgp_update_down<16, true, 0>,  (const void*)k_mgp_update_down<0, true, 0>,
                          (const void*)k_mgp_update_down<64, true, 1>,  (const void*)k_mgp_update_down<32, true, 1>,
                          (const void*)k_mgp_update_down<16, true, 1>,  (const void*)k_mgp_update_down<0, true, 1>,
                          (const void*)k_mgp_update_down<64, true, 2>,  (const void*)k_mgp_update_down<32, true, 2>,
                          (const void*)k_mgp_update_down<16, true, 2>,  (const void*)k_mgp_update_down<0, true, 2>};
    for (int i = 0; i < 16; ++i) {
        CK(cudaFuncSetAttribute(ffns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        cudaFuncAttributes fa;
        CK(cudaFuncGetAttributes(&fa, ffns[i]));
        tile_maxt_down = std::min(tile_maxt_down, fa.maxThreadsPerBlock);
    }
    const void* pfns[] = {(const void*)k_mgp_down<64>, (const void*)k_mgp_down<32>, (const void*)k_mgp_down<16>, (const void*)k_mgp_down<0>,
                          (const void*)k_mgp_up<64>,   (const void*)k_mgp_up<32>,   (const void*)k_mgp_up<16>,   (const void*)k_mgp_up<0>};
    for (int i = 0; i < 8; ++i) {
        CK(cudaFuncSetAttribute(pfns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        cudaFuncAttributes fa;
        CK(cudaFuncGetAttributes(&fa, pfns[i]));
        int& m = i < 4 ? tile_maxt_down : tile_maxt_up;
        m = std::min(m, fa.maxThreadsPerBlock);
    }
    CK(cudaDeviceGetAttribute(&tile_nsm, cudaDevAttrMultiProcessorCount, device));
    CK(cudaFuncSetAttribute(k_mgt_tail<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(k_mgt_tail<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    // vertex-class tables of every level
    for (int* p : tile_rowv) cudaFree(p);
    for (int* p : tile_colv) cudaFree(p);
    tile_rowv.clear(); tile_colv.clear();
    auto vclass = [](int v, int n, int nblk) {
        if (v < 1 || v > nblk * n - 1) return -1;
        return (v % n) ? 2 * (v / n) : 2 * (v / n) - 1;
    };
    for (const LevelGeo& g : levels) {
        std::vector<int> rv(g.R + 1 + 2 * ROMHC_ROWV_PAD), cv(g.P);
        for (int i = 0; i < (int)rv.size(); ++i) rv[i] = vclass(i - ROMHC_ROWV_PAD, g.N, g.nrb);
        for (int c = 0; c < g.P; ++c) cv[c] = vclass(c, g.N, g.ncb);
        int *dr = nullptr, *dc = nullptr;
        CK(cudaMalloc(&dr, rv.size() * sizeof(int)));
        CK(cudaMalloc(&dc, cv.size() * sizeof(int)));
        CK(cudaMemcpy(dr, rv.data(), rv.size() * sizeof(int), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dc, cv.data(), cv.size() * sizeof(int), cudaMemcpyHostToDevice));
        tile_rowv.push_back(dr); tile_colv.push_back(dc);
    }
    tile_ready = true;
    return ROMHC_OK;
}

// CTAs of `func` resident on the whole GPU (the distance, in CTA ids, to the CTA that takes this one's slot next)
int Context::tile_pf_dist(const void* func, int threads, size_t smem) {
    if (!tile_prefetch) return 0;
    int per_sm = 1, nsm = 148;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, func, threads, smem) != cudaSuccess) per_sm = 1;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device);
    return std::max(1, per_sm) * nsm;
}

// per (strip, row group): row types of the tile rows (2 bits each: 0 not interior, 1 primary class, 2 other) and the
// primary class (the first non-interface interior row's), cached per (level, strip height, halo, region rows)
const int* Context::tile_rinfo(int l, int TY, int halo_top, int NR) {
    const std::array<int, 4> key{l, TY, halo_top, NR};
    auto it = tile_rinfo_cache.find(key);
    if (it != tile_rinfo_cache.end()) return it->second;
    const LevelGeo& g = levels[l];
    auto vclass = [&](int v) {
        if (v < 1 || v > g.R - 1) return -1;
        return (v % g.N) ? 2 * (v / g.N) : 2 * (v / g.N) - 1;
    };
    const int ns = (g.R + TY - 1) / TY, NRG = NR / 4;
    std::vector<int> tab(size_t(ns) * NRG * 3);
    for (int s = 0; s < ns; ++s)
        for (int ty = 0; ty < NRG; ++ty) {
            const int rho0 = s * TY - halo_top + 4 * ty;
            int rv[4], prim = -1, rt = 0;
            for (int i = 0; i < 4; ++i) {
                rv[i] = vclass(rho0 + i);
                if (prim < 0 && rv[i] >= 0 && !(rv[i] & 1)) prim = rv[i];
            }
            for (int i = 0; i < 4; ++i) rt |= (rv[i] < 0 ? 0 : (rv[i] == prim ? 1 : 2)) << (2 * i);
            int* o = &tab[(size_t(s) * NRG + ty) * 3];
            o[0] = (prim * 256) | rt;
            o[1] = int(unsigned(rv[0] + 1) | (unsigned(rv[1] + 1) << 16));
            o[2] = int(unsigned(rv[2] + 1) | (unsigned(rv[3] + 1) << 16));
        }
    int* d = nullptr;
    cudaError_t e = cudaMalloc(&d, tab.size() * sizeof(int));
    if (e == cudaSuccess) e = cudaMemcpy(d, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice);
    if (e != cudaSuccess || !d) {
        fprintf(stderr, "tile_rinfo(l=%d, TY=%d, halo=%d, NR=%d): %zu ints: %s\n", l, TY, halo_top, NR, tab.size(), cudaGetErrorString(e));
        if (d) cudaFree(d);
        return nullptr;
    }
    tile_rinfo_cache[key] = d;
    return d;
}

// grid of a persistent kernel: every CTA slot of the GPU (SMs x resident CTAs per SM), at most one CTA per item
int Context::tile_persistent_grid(const void* func, int threads, size_t smem, int64_t items) {
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, func, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    return int(std::min<int64_t>(int64_t(tile_nsm) * per_sm, items));
}

int Context::tile_ntab() const { return (2 * nrb - 1) * (2 * ncb - 1) + 1; }

int Context::tile_weight_table(const double* y, int Kc, cudaStream_t st) {
    ++g_launches; k_weight_table<<<Kc, 64, 0, st>>>(y, ws.wtab, nrb, ncb, Kc);
    CK(cudaGetLastError());
    return ROMHC_OK;
}

// multigrid tail on register tiles; returns ROMHC_ERR_ARG if the configuration does not fit (caller falls back)
int Context::tile_tail(const double* y, int Kc, double* part_rz, cudaStream_t st) {
    (void)y;
    const int L = int(levels.size()) - 1;
    TailTileArgs p;
    memset(&p, 0, sizeof(p));
    p.nlev = L - tail_level + 1;
    p.tab = ws.wtab; p.ntab = tile_ntab(); p.ncv = 2 * ncb - 1;
    p.direct = coarse_direct ? 1 : 0; p.DL = coarse_D; p.LD = coarse_LD; p.coarse_sweeps = coarse_sweeps; p.nu = nu_tail;
    const LevelGeo& g0 = levels[tail_level];
    const int CG0 = g0.P / 4, NRG0 = (g0.R + 3) / 4;
    int threads = ((CG0 * NRG0 + 31) / 32) * 32;
    int off = p.ntab * TWD + 34;
    p.off_ex = off;
    int exmax = 0;
    for (int l = 0; l < p.nlev; ++l) {
        const LevelGeo& g = levels[tail_level + l];
        const int CG = g.P / 4, NRG = (g.R + 3) / 4;
        exmax = std::max(exmax, 2 * 4 * NRG * (CG + 2) + 2 * (NRG + 2) * 4 * CG);
        threads = std::max(threads, ((CG * NRG + 31) / 32) * 32);
    }
    off += exmax;
    for (int l = 0; l < p.nlev; ++l) {
        p.geo[l] = levels[tail_level + l];
        p.rowv[l] = tile_rowv[tail_level + l] + ROMHC_ROWV_PAD;
        p.colv[l] = tile_colv[tail_level + l];
        p.off_z[l] = off; off += p.geo[l].Dp;
        if (l > 0 || p.nlev == 1) { p.off_r[l] = off; off += p.geo[l].Dp; } else p.off_r[l] = -1;
    }
    p.off_fac = off;
    if (coarse_direct) off += coarse_D * coarse_LD;
    p.n_doubles = off;
    const size_t sm = size_t(off) * 8;
    if (threads > 256 || sm > 227 * 1024) return ROMHC_ERR_ARG;
    ++g_launches;
    if (threads <= 128) k_mgt_tail<128><<<Kc, threads, sm, st>>>(p, ws.r[tail_level], ws.za[tail_level], ws.cfac, ws.active, part_rz);
    else                k_mgt_tail<256><<<Kc, threads, sm, st>>>(p, ws.r[tail_level], ws.za[tail_level], ws.cfac, ws.active, part_rz);
    return ROMHC_OK;
}

// can the tile kernels run level l?  (column groups per CTA, register budget, shared memory)
bool Context::tile_level_ok(int l) const {
    if (!use_tile) return false;
    const LevelGeo& g = levels[l];
    const int CG = g.P / 4;
    const int maxt = std::min(tile_maxt_down, tile_maxt_up);
    const int nu = nu_of(l);
    const int nrg_min = (2 + 4 * nu + 2 + 3) / 4;      // TY = 2 in the down kernel
    if (CG * nrg_min > maxt || CG > 128) return false;
    return true;
}

// strip height (even) and region rows (a multiple of 4, the tile height): the tallest region whose CTA fits the
// thread budget, capped by tile_ty_cap and by the grid height
static void tile_pick_ty(const LevelGeo& g, int extra_rows, int maxt, int cap, int* TY_out, int* NR_out) {
    const int CG = g.P / 4;
    const int nrg = std::max(1, std::min(maxt / CG, 16));
    int TY = 4 * nrg - extra_rows;
    TY = std::min(TY, cap);
    TY = std::max(TY & ~1, 2);
    const int ns = (g.R + TY - 1) / TY;                  // strips needed at the tallest height ...
    TY = std::min(TY, (((g.R + ns - 1) / ns) + 1) & ~1);  // ... then balance the rows over them
    *TY_out = TY;
    *NR_out = ((TY + extra_rows + 3) / 4) * 4;
}

// will the going-up kernel of level l run as the persistent tile kernel (the only reader that understands an fp32 z_A)?
bool Context::tile_up_persistent_ok(int l) const {
    if (!use_tile || !tile_persistent || l >= tail_level || !tile_level_ok(l)) return false;
    const LevelGeo& g = levels[l];
    const bool has_c = fused_coarse(l);
    const int CG = g.P / 4, nu = nu_of(l);
    for (int cap = tile_ty_cap; cap >= 2; cap -= 2) {
        int TY, NR;
        tile_pick_ty(g, 4 * nu, tile_maxt_up, cap, &TY, &NR);
        if (tile_stage_bytes(tile_ntab(), NR, CG, (g.R + TY - 1) / TY, true, has_c ? NR / 2 + 1 : 0, has_c ? levels[l + 1].P : g.P) <= 227 * 1024)
            return true;
    }
    return false;
}

int Context::tile_down(int l, const double* y, int Kc, cudaStream_t st) {
    (void)y;
    TileArgs a;
    memset(&a, 0, sizeof(a));
    a.g = levels[l];
    a.has_coarse = fused_coarse(l) ? 1 : 0;
    a.gc = a.has_coarse ? levels[l + 1] : levels[l];
    a.rowv = tile_rowv[l] + ROMHC_ROWV_PAD; a.colv = tile_colv[l];
    a.tab = ws.wtab; a.ntab = tile_ntab(); a.ncv = 2 * ncb - 1;
    const int nu = nu_of(l);
    a.nu = nu;
    a.halo_top = 2 * nu + 2;
    a.pf_dist = 0; a.rinfo = nullptr;
    const int CG = a.g.P / 4;
    if (tile_persistent) {
        for (int cap = tile_ty_cap; cap >= 2; cap -= 2) {
            tile_pick_ty(a.g, 4 * nu + 2, tile_maxt_down, cap, &a.TY, &a.NR);
            a.ns = (a.g.R + a.TY - 1) / a.TY;
            if (tile_stage_bytes(a.ntab, a.NR, CG, a.ns, false, 0, 0) <= 227 * 1024) break;
        }
        const size_t sm = tile_stage_bytes(a.ntab, a.NR, CG, a.ns, false, 0, 0);
        if (sm <= 227 * 1024) {
            a.ns = (a.g.R + a.TY - 1) / a.TY;
            auto fn = CG == 64 ? k_mgp_down<64> : (CG == 32 ? k_mgp_down<32> : (CG == 16 ? k_mgp_down<16> : k_mgp_down<0>));
            a.rinfo = tile_rinfo(l, a.TY, a.halo_top, a.NR);
            if (!a.rinfo) { set_error("tile kernels: row-info table allocation failed"); return ROMHC_ERR_CUDA; }
            const int grid = tile_persistent_grid((const void*)fn, CG * (a.NR / 4), sm, int64_t(Kc) * a.ns);
            ++g_launches;
            a.emit_res = (l == bridge_level) ? 1 : 0;             // residual into zb[l] (free until the way up)
            a.in_f32 = (l == 0 && use_z32 >= 2 && z32_want && l != bridge_level && tile_up_persistent_ok(l)) ? 1 : 0;
            if (l == 0) za_f32 = a.in_f32 != 0;
            fn<<<grid, dim3(CG, a.NR / 4), sm, st>>>(a, ws.r[l], ws.za[l], a.has_coarse ? ws.r[l + 1] : (a.emit_res ? ws.zb[l] : nullptr),
                                                     ws.active, Kc);
            bridge_res_emitted = a.emit_res != 0;
            return ROMHC_OK;
        }
    }
    tile_pick_ty(a.g, 4 * nu + 2, tile_maxt_down, tile_ty_cap, &a.TY, &a.NR);
    a.ns = (a.g.R + a.TY - 1) / a.TY;
    const size_t sm = tile_smem_bytes(a.ntab, a.NR, CG);
    if (sm > 227 * 1024) { set_error("tile kernels: shared memory"); return ROMHC_ERR_ARG; }
    auto fn = CG == 64 ? k_mgt_down<64> : (CG == 32 ? k_mgt_down<32> : k_mgt_down<0>);
    a.pf_dist = tile_pf_dist((const void*)fn, CG * (a.NR / 4), sm);
    ++g_launches;
    fn<<<dim3(a.ns, Kc), dim3(CG, a.NR / 4), sm, st>>>(a, ws.r[l], ws.za[l], a.has_coarse ? ws.r[l + 1] : nullptr, ws.active);
    return ROMHC_OK;
}

// x += alpha p, r -= alpha A p and the going-down kernel of level l in one launch (persistent tile kernels only);
// returns ROMHC_ERR_ARG if the configuration does not fit (the caller then runs the two kernels separately)
// can the fused update + going-down kernel run on level 0 (the only reader that understands an fp32 p)?
bool Context::tile_fused_ok() const {
    if (!use_tile || !tile_persistent || !use_fused || tail_level < 1 || !tile_level_ok(0)) return false;   // tail_level 0: level 0 lives in the tail kernel
    const LevelGeo& g = levels[0];
    const int CG = g.P / 4, nu = nu_of(0);
    int TY = 0, NR = 0;
    size_t sm = 0;
    for (int cap = tile_ty_cap; cap >= 2; cap -= 2) {
        tile_pick_ty(g, 4 * nu + 3, tile_maxt_down, cap, &TY, &NR);
        sm = tile_stage_bytes(tile_ntab(), NR, CG, (g.R + TY - 1) / TY, true, 0, 0);
        if (sm <= 227 * 1024) break;
    }
    return sm <= 227 * 1024 && NR >= 4 * nu + 6;
}

int Context::tile_update_down(int l, int Kc, const double* p, double* x, const double* alpha, cudaStream_t st) {
    if (!tile_persistent || !use_fused) return ROMHC_ERR_ARG;
    TileArgs a;
    memset(&a, 0, sizeof(a));
    a.g = levels[l];
    a.has_coarse = fused_coarse(l) ? 1 : 0;
    a.gc = a.has_coarse ? levels[l + 1] : levels[l];
    a.rowv = tile_rowv[l] + ROMHC_ROWV_PAD; a.colv = tile_colv[l];
    a.tab = ws.wtab; a.ntab = tile_ntab(); a.ncv = 2 * ncb - 1;
    const int nu = nu_of(l);
    a.nu = nu;
    a.halo_top = 2 * nu + 2;
    a.pf_dist = 0; a.rinfo = nullptr;
    const int CG = a.g.P / 4;
    auto bytes = [&]() { return tile_stage_bytes(a.ntab, a.NR, CG, (a.g.R + a.TY - 1) / a.TY, true, 0, 0); };
    for (int cap = tile_ty_cap; cap >= 2; cap -= 2) {
        tile_pick_ty(a.g, 4 * nu + 3, tile_maxt_down, cap, &a.TY, &a.NR);
        if (bytes() <= 227 * 1024) break;
    }
    const size_t sm = bytes();
    if (sm > 227 * 1024 || a.NR < 4 * nu + 6) return ROMHC_ERR_ARG;
    a.ns = (a.g.R + a.TY - 1) / a.TY;
#define ROMHC_UD(P32_, XM_) (CG == 64 ? k_mgp_update_down<64, P32_, XM_> : (CG == 32 ? k_mgp_update_down<32, P32_, XM_> : \
                             (CG == 16 ? k_mgp_update_down<16, P32_, XM_> : k_mgp_update_down<0, P32_, XM_>)))
    const int xm = p_f32 ? x_mode : 0;       // deferred update of the iterate: own instantiations, the plain kernel is untouched
    auto fn = !p_f32 ? ROMHC_UD(false, 0) : (xm == 1 ? ROMHC_UD(true, 1) : (xm == 2 ? ROMHC_UD(true, 2) : ROMHC_UD(true, 0)));
#undef ROMHC_UD
    a.rinfo = tile_rinfo(l, a.TY, a.halo_top, a.NR);
    if (!a.rinfo) { set_error("tile kernels: row-info table allocation failed"); return ROMHC_ERR_CUDA; }
    const int grid = tile_persistent_grid((const void*)fn, CG * (a.NR / 4), sm, int64_t(Kc) * a.ns);
    ++g_launches;
    if (l != 0) return ROMHC_ERR_ARG;
    // systems that are no longer active keep their (converged) residual in the old buffer: nobody reads it again
    a.emit_res = (l == bridge_level) ? 1 : 0;
    a.in_f32 = (use_z32 >= 2 && z32_want && l != bridge_level && tile_up_persistent_ok(l)) ? 1 : 0;
    za_f32 = a.in_f32 != 0;
    a.p_f32 = p_f32 ? 1 : 0;
    a.p_prev = x_p_prev;
    a.alpha_prev = x_alpha_prev;
    fn<<<grid, dim3(CG, a.NR / 4), sm, st>>>(a, p, x, ws.r[0], ws.r_alt, alpha, ws.za[0],
                                             a.has_coarse ? ws.r[1] : (a.emit_res ? ws.zb[0] : nullptr), ws.active, Kc);
    bridge_res_emitted = a.emit_res != 0;
    std::swap(ws.r[0], ws.r_alt);
    return ROMHC_OK;
}

int Context::tile_up(int l, const double* y, int Kc, const double* e, double* part_rz, int* ns_out, cudaStream_t st) {
    (void)y;
    TileArgs a;
    memset(&a, 0, sizeof(a));
    a.g = levels[l];
    a.has_coarse = fused_coarse(l) ? 1 : 0;
    a.gc = a.has_coarse ? levels[l + 1] : levels[l];
    a.rowv = tile_rowv[l] + ROMHC_ROWV_PAD; a.colv = tile_colv[l];
    a.tab = ws.wtab; a.ntab = tile_ntab(); a.ncv = 2 * ncb - 1;
    const int nu = nu_of(l);
    a.nu = nu;
    a.halo_top = 2 * nu;
    a.pf_dist = 0; a.rinfo = nullptr;
    const int CG = a.g.P / 4;
    if (tile_persistent) {
        auto bytes = [&]() { return tile_stage_bytes(a.ntab, a.NR, CG, (a.g.R + a.TY - 1) / a.TY, true, a.has_coarse ? a.NR / 2 + 1 : 0, a.gc.P); };
        for (int cap = tile_ty_cap; cap >= 2; cap -= 2) {
            tile_pick_ty(a.g, 4 * nu, tile_maxt_up, cap, &a.TY, &a.NR);
            if (bytes() <= 227 * 1024) break;
        }
        const size_t sm = bytes();
        if (sm <= 227 * 1024) {
            a.ns = (a.g.R + a.TY - 1) / a.TY;
            auto fn = CG == 64 ? k_mgp_up<64> : (CG == 32 ? k_mgp_up<32> : (CG == 16 ? k_mgp_up<16> : k_mgp_up<0>));
            a.rinfo = tile_rinfo(l, a.TY, a.halo_top, a.NR);
            if (!a.rinfo) { set_error("tile kernels: row-info table allocation failed"); return ROMHC_ERR_CUDA; }
            const int grid = tile_persistent_grid((const void*)fn, CG * (a.NR / 4), sm, int64_t(Kc) * a.ns);
            ++g_launches;
            a.out_f32 = (l == 0 && use_z32 && z32_want) ? 1 : 0;
            a.in_f32 = (l == 0 && za_f32) ? 1 : 0;
            fn<<<grid, dim3(CG, a.NR / 4), sm, st>>>(a, e, ws.za[l], ws.r[l], ws.zb[l], ws.active, part_rz, Kc);
            if (l == 0) z32_out = a.out_f32 != 0;
            *ns_out = a.ns;
            return ROMHC_OK;
        }
    }
    if (l == 0 && za_f32) { set_error("tile kernels: fp32 z_A without the persistent going-up kernel"); return ROMHC_ERR_ARG; }
    tile_pick_ty(a.g, 4 * nu, tile_maxt_up, tile_ty_cap, &a.TY, &a.NR);
    a.ns = (a.g.R + a.TY - 1) / a.TY;
    const size_t sm = tile_smem_bytes(a.ntab, a.NR, CG);
    if (sm > 227 * 1024) { set_error("tile kernels: shared memory"); return ROMHC_ERR_ARG; }
    auto fn = CG == 64 ? k_mgt_up<64> : (CG == 32 ? k_mgt_up<32> : k_mgt_up<0>);
    a.pf_dist = tile_pf_dist((const void*)fn, CG * (a.NR / 4), sm);
    ++g_launches;
    fn<<<dim3(a.ns, Kc), dim3(CG, a.NR / 4), sm, st>>>(a, e, ws.za[l], ws.r[l], ws.zb[l], ws.active, part_rz);
    *ns_out = a.ns;
    return ROMHC_OK;
}

}  // namespace romhc
