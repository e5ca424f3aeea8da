"""Sparse CPU restatement of the reference's FEM layer (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/src/lib/SolutionsManagers.py:
  * mesh / element stiffness ......... :146-219  (`SolutionsManagerFEM.__init__`, inner `A(a)` :187-215)
  * load vector f == 1 ................ :177-185
  * full-order solve .................. :17-40, :64-68 (`galerkin`, `generate_solutions`)
  * H10 / l2 norms .................... :49, :56-62
  * reduced Galerkin .................. :88-106 (`generate_fm_solutions`)
  * H10 projection .................... :108-139 (`project_solutions`)
  * P1 point evaluation ............... :221-244 (`evaluate_solutions`), :70-86 (`generate_riesz`)

The reference stores every subdomain stiffness as a dense (D, D) array, which
cannot exist beyond D ~ 4k.  Here the same matrices are assembled element by
element into scipy CSR (identical entries, see tests/test_oracle_golden.py) and
solved with SuperLU (`splu`) -- the same factorisation the reference's
`method="lsqsparse"` reaches through `spsolve`.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spl

# P1 stiffness of a right isosceles triangle (legs h, unit coefficient): vertex 0 is the
# right-angle corner.  Independent of h in 2-D.  (reference :196-214, before the global "/ 2")
_KE = 0.5 * np.array([[2.0, -1.0, -1.0], [-1.0, 1.0, 0.0], [-1.0, 0.0, 1.0]])


def _triangles(nrb: int, ncb: int, N: int):
    """Vertex triples (right-angle vertex first) and owning block of every triangle.

    Vertex grid is (R+1) x (C+1), row-major, rows = y (reference :156-163).  Each cell
    (line, column) is cut by the diagonal joining (line, column+1)-(line+1, column): the
    "even" triangle has its right angle at (line, column) (:193-203), the "odd" one at
    (line+1, column+1) (:204-214).
    """
    R, C = nrb * N, ncb * N
    ncv = C + 1
    line, col = np.meshgrid(np.arange(R), np.arange(C), indexing="ij")
    line, col = line.ravel(), col.ravel()
    pos = ncv * line + col
    even = np.stack([pos, pos + 1, pos + ncv], axis=1)
    pos2 = ncv * (line + 1) + col + 1
    odd = np.stack([pos2, pos2 - 1, pos2 - ncv], axis=1)
    blk = (line // N) * ncb + (col // N)          # a[line // N, column // N]  (:190-192)
    return np.concatenate([even, odd]), np.concatenate([blk, blk])


def _interior_index(nrb: int, ncb: int, N: int):
    R, C = nrb * N, ncb * N
    idx = -np.ones((R + 1, C + 1), dtype=np.int64)
    idx[1:R, 1:C] = np.arange((R - 1) * (C - 1)).reshape(R - 1, C - 1)   # (:160-163)
    return idx.ravel()


def block_stiffness_csr(nrb: int, ncb: int, N: int):
    """List of the nrb*ncb sparse matrices A_pq (row-major over (p, q)); reference :215-218."""
    tri, blk = _triangles(nrb, ncb, N)
    idx = _interior_index(nrb, ncb, N)
    D = (nrb * N - 1) * (ncb * N - 1)
    ti = idx[tri]                                   # (T, 3) interior numbers or -1
    mats = []
    for b in range(nrb * ncb):
        sel = ti[blk == b]
        rows = np.repeat(sel, 3, axis=1).ravel()
        cols = np.tile(sel, (1, 3)).ravel()
        vals = np.tile(_KE.ravel(), len(sel))
        ok = (rows >= 0) & (cols >= 0)
        mats.append(sp.csr_matrix((vals[ok], (rows[ok], cols[ok])), shape=(D, D)))
    return mats


def stiffness_csr(a: np.ndarray, N: int, blocks=None):
    """A(a) = sum_pq a_pq A_pq  (reference `galerkin` :19-23) as CSR."""
    a = np.asarray(a, dtype=np.float64)
    nrb, ncb = a.shape
    blocks = block_stiffness_csr(nrb, ncb, N) if blocks is None else blocks
    A = None
    for coef, M in zip(a.ravel(), blocks):
        A = coef * M if A is None else A + coef * M
    return A.tocsr()


def load_vector(nrb: int, ncb: int, N: int):
    """f == 1 load vector; loops restated from reference :177-185 (vectorised)."""
    R, C = nrb * N, ncb * N
    area = (1.0 / N) * (1.0 / N)
    B = np.zeros((R + 1, C + 1))
    B[:-1, :-1] += area / 6
    B[1:, :-1] += area / 3
    B[:-1, 1:] += area / 3
    B[1:, 1:] += area / 6
    return B[1:-1, 1:-1].reshape(-1).copy()


class FEMOracle:
    """Mirror of `SolutionsManagerFEM` on sparse matrices (never builds A_preassembled)."""

    def __init__(self, blocks_geometry, N: int):
        nrb, ncb = blocks_geometry
        self.blocks_geometry = (nrb, ncb)
        self.N = N
        self.x_domain = (-ncb / 2.0, ncb / 2.0)
        self.y_domain = (-nrb / 2.0, nrb / 2.0)
        self.nc_inner_vertices = ncb * N - 1
        self.nr_inner_vertices = nrb * N - 1
        self.nc_cells = ncb * N + 1
        self.nr_cells = nrb * N + 1
        self.vspace_dim = self.nc_inner_vertices * self.nr_inner_vertices
        self.points_c = np.linspace(*self.x_domain, self.nc_cells)
        self.points_r = np.linspace(*self.y_domain, self.nr_cells)
        self.blocks = block_stiffness_csr(nrb, ncb, N)
        self.B_total = load_vector(nrb, ncb, N)
        self.A1 = stiffness_csr(np.ones((nrb, ncb)), N, self.blocks)   # A_preassembled4h1_norm (:49)

    # -- full-order solves ------------------------------------------------------------
    def matrix(self, a):
        return stiffness_csr(a, self.N, self.blocks)

    def generate_solutions(self, a2try):
        a2try = np.asarray(a2try, dtype=np.float64)
        out = np.empty((len(a2try), self.vspace_dim))
        for k, a in enumerate(a2try):
            out[k] = spl.splu(self.matrix(a).tocsc()).solve(self.B_total)
        return out

    # -- norms --------------------------------------------------------------------------
    def H10norm(self, solutions):
        S = np.asarray(solutions, dtype=np.float64)
        return np.sqrt(np.einsum("kd,kd->k", S, (self.A1 @ S.T).T))

    @staticmethod
    def l2norm(solutions):
        return np.sqrt(np.sum(np.square(solutions), axis=1))

    # -- reduced problems ---------------------------------------------------------------
    def reduced_operators(self, basis):
        Phi = np.asarray(basis, dtype=np.float64)
        nrb, ncb = self.blocks_geometry
        Ahat = np.stack([Phi @ (M @ Phi.T) for M in self.blocks]).reshape(nrb, ncb, len(Phi), len(Phi))
        return Ahat, Phi @ self.B_total

    def reduced_coefficients(self, a, basis):
        Ahat, bhat = self.reduced_operators(basis)
        a = np.asarray(a, dtype=np.float64)
        Ak = np.einsum("pqij,kpq->kij", Ahat, a)
        return np.linalg.solve(Ak, np.broadcast_to(bhat, (len(a), len(bhat)))[..., None])[..., 0]

    def generate_fm_solutions(self, a, coefficients_rom):
        if len(coefficients_rom) == 0:
            return np.zeros((len(a), self.vspace_dim))
        Phi = np.asarray(coefficients_rom, dtype=np.float64)
        return self.reduced_coefficients(a, Phi) @ Phi

    def projection_coefficients(self, solutions, basis):
        Phi = np.asarray(basis, dtype=np.float64)
        U = np.asarray(solutions, dtype=np.float64)
        W = (self.A1 @ Phi.T).T                      # (n, D)
        return np.linalg.solve(Phi @ W.T, W @ U.T).T  # (K, n)

    def project_solutions(self, solutions, coefficients_rom):
        if len(coefficients_rom) == 0:
            return np.zeros((len(solutions), self.vspace_dim))
        Phi = np.asarray(coefficients_rom, dtype=np.float64)
        return self.projection_coefficients(solutions, Phi) @ Phi

    # -- P1 point evaluation --------------------------------------------------------------
    def interpolation_matrix(self, points):
        """(m, D) sparse matrix E with E @ u == evaluate_solutions(points, [u])[0]  (:221-244)."""
        points = np.asarray(points, dtype=np.float64).reshape(-1, 2)
        nrc, ncc = self.nr_cells, self.nc_cells
        idx = -np.ones((nrc, ncc), dtype=np.int64)
        idx[1:-1, 1:-1] = np.arange(self.vspace_dim).reshape(self.nr_inner_vertices, self.nc_inner_vertices)
        rows, cols, vals = [], [], []
        for i, (x, y) in enumerate(points):
            px = int(np.searchsorted(self.points_c, x)) - 1
            py = int(np.searchsorted(self.points_r, y)) - 1
            qx = (x - self.points_c[px]) / (self.points_c[px + 1] - self.points_c[px])
            qy = (y - self.points_r[py]) / (self.points_r[py + 1] - self.points_r[py])
            if qx + qy < 1:
                terms = [(1 - qx - qy, px, py), (qx, px + 1, py), (qy, px, py + 1)]
            else:
                terms = [(qx + qy - 1, px + 1, py + 1), (1 - qx, px, py + 1), (1 - qy, px + 1, py)]
            for w, ix, iy in terms:                   # val.T[ix, iy] == grid[iy, ix]
                j = idx[iy, ix]
                if j >= 0:
                    rows.append(i), cols.append(j), vals.append(w)
        return sp.csr_matrix((vals, (rows, cols)), shape=(len(points), self.vspace_dim))

    def evaluate_solutions(self, points, solutions):
        E = self.interpolation_matrix(points)
        return np.asarray((E @ np.asarray(solutions, dtype=np.float64).T).T)

    def generate_riesz(self, x, norm="h10"):
        if norm == "l2":
            return self.interpolation_matrix(x).toarray()
        raise Exception("Not implemented.")
