"""Drop-in for the reference's `src/lib/SolutionsManagers.py` backed by the B200 kernels.

Same names, signatures, argument meaning and error behaviour as
/root/reference/src/lib/SolutionsManagers.py (line numbers cited per method); numpy in, fresh C-contiguous
float64 numpy out.  The dense tensor `A_preassembled` (nrb, ncb, D, D) of the reference is never formed: the
stiffness is applied matrix-free on the device (it remains available as a lazily built property for the small
sizes where it fits).  Extra, purely additive keyword arguments (`return_coefs`, `as_device`) let large batches
stay on the device.  There is no CPU fallback.
"""
from __future__ import annotations

from typing import List, Tuple, Union

import numpy as np

_METHODS = ("lsq", "lsqsparse", "ridge")   # the reference's three solvers agree to <= 1e-12: one GPU path


def _check_method(method):
    if not isinstance(method, str) or method.lower() not in _METHODS:
        raise Exception(f"Method {method} Not implemented.")        # reference :39


def h1_error(v: List[np.ndarray]):                                    # reference :13-14 (unused helper)
    return np.sqrt(np.mean(np.sum(np.power(np.gradient(v, axis=(1, 2)), 2), axis=0), axis=(1, 2)))


class _DenseOps:
    """Handle-free device primitives of libromhc (GEMMs, batched SPD solves, row norms / dots) for callers that hold
    dense operators: galerkin() and the generic SolutionsManager(A_preassembled, B_total).  No CPU fallback."""

    def __init__(self):
        import torch
        from .. import _lib
        if not torch.cuda.is_available():
            raise _lib.RomhcError("no CUDA device: the ROMHighContrast B200 path has no CPU fallback")
        self.torch, self._lib = torch, _lib
        self.device = torch.device("cuda", torch.cuda.current_device())

    def dev(self, a):
        return self.torch.as_tensor(np.ascontiguousarray(np.asarray(a, dtype=np.float64)), device=self.device)

    def _p(self, t):
        import ctypes as C
        return C.c_void_p(t.data_ptr())

    def _st(self):
        import ctypes as C
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def empty(self, *shape, dtype=None):
        return self.torch.empty(*shape, dtype=dtype or self.torch.float64, device=self.device)

    def gemm_nt(self, A, B):
        out = self.empty(A.shape[0], B.shape[0])
        self._lib.call("romhc_gemm_nt", self._p(A), A.stride(0), self._p(B), B.stride(0), self._p(out), out.stride(0), A.shape[0],
                       B.shape[0], A.shape[1], 0, self._st())
        return out

    def gemm_nn(self, A, B):
        out = self.empty(A.shape[0], B.shape[1])
        self._lib.call("romhc_gemm_nn", self._p(A), A.stride(0), self._p(B), B.stride(0), self._p(out), out.stride(0), A.shape[0],
                       B.shape[1], A.shape[1], self._st())
        return out

    def solve_spd(self, y, Aq, rhs):
        """(sum_q y[k][q] Aq[q]) c_k = rhs (shared (n,) or per system (K, n)) -> (K, n); LinAlgError if not SPD."""
        K, nb, n = y.shape[0], y.shape[1], Aq.shape[-1]
        out = self.empty(K, n)
        info = self.empty(K, dtype=self.torch.int32)
        self._lib.call("romhc_reduced_solve", self._p(y), nb, self._p(Aq), self._p(rhs), 1 if rhs.dim() == 2 else 0, n, K,
                       self._p(out), self._p(info), self._st())
        if bool(info.any().item()):
            raise np.linalg.LinAlgError("Matrix is not positive definite.")
        return out

    def row_dots(self, X, Y):
        out = self.empty(X.shape[0])
        self._lib.call("romhc_row_dots", self._p(X), X.stride(0), self._p(Y), Y.stride(0), X.shape[0], X.shape[1], self._p(out),
                       self._st())
        return out


def galerkin(a, B_total, A_preassembled, method="lsq"):
    """One dense system (sum_pq a_pq A_pq) c = B_total on the caller's operators (reference :17-40), any size: the
    batched device solver with K = 1 (n <= 64: register / shared-memory Cholesky kernels; larger: blocked Cholesky).
    All three reference methods solve the same SPD system (they agree to 1e-12) and share this path."""
    _check_method(method)
    ops = _DenseOps()
    a = np.asarray(a, dtype=np.float64)
    A = np.asarray(A_preassembled, dtype=np.float64)
    b = np.asarray(B_total, dtype=np.float64)
    n = b.shape[0]
    return ops.solve_spd(ops.dev(a.reshape(1, -1)), ops.dev(A.reshape(-1, n, n)), ops.dev(b))[0].cpu().numpy()


class SolutionsManager:
    """Base class of the reference (:43-142).  Constructed directly it is the reference's generic manager over dense
    preassembled operators `A_preassembled` (nrb, ncb, D, D) and a load vector `B_total` (D,): the operators are uploaded
    once and every method runs on the device primitives of libromhc (batched SPD solves, GEMMs).  `SolutionsManagerFEM`
    overrides the data path with the matrix-free kernels and never forms the dense tensor."""

    def __init__(self, A_preassembled, B_total, num_cores=1, method="lsq"):      # reference :44-51
        self.method = method
        self.mapfunction = map                    # num_cores is accepted and ignored: the batch runs on the GPU
        if A_preassembled is None:                # matrix-free subclass (SolutionsManagerFEM) sets its own attributes
            return
        self.vspace_dim = len(B_total)
        self.blocks_geometry = np.shape(A_preassembled)[:2]
        self.A_preassembled = A_preassembled
        self.A_preassembled4h1_norm = np.einsum("abij->ij", self.A_preassembled)
        self.B_total = B_total

    # ---- dense-operator data path (generic manager) ----------------------------------------------------------------
    def _dense_ops_(self):
        st = self.__dict__.get("_dense_state")
        if st is None:
            ops = _DenseOps()
            D = self.vspace_dim
            st = {"ops": ops, "A": ops.dev(np.asarray(self.A_preassembled, dtype=np.float64).reshape(-1, D, D)),
                  "A4": ops.dev(self.A_preassembled4h1_norm), "b": ops.dev(self.B_total)}
            self.__dict__["_dense_state"] = st
        return st

    def _dense_reduced_operators(self, Phi):
        """Phi (n, D) device -> Ahat (nb, n, n) = Phi A_q Phi^T (reference :93-103)."""
        st = self._dense_ops_()
        ops, A = st["ops"], st["A"]
        n = Phi.shape[0]
        Ahat = ops.empty(A.shape[0], n, n)
        for q in range(A.shape[0]):
            Ahat[q] = ops.gemm_nt(ops.gemm_nn(Phi, A[q]), Phi)
        return Ahat

    def __str__(self):
        return self.__class__.__name__

    # ---- device plumbing ---------------------------------------------------------------------------------
    def _engine_(self):
        eng = self.__dict__.get("_engine")
        if eng is None:
            from ..engine import Engine
            eng = Engine(self.blocks_geometry, self.N)
            self.__dict__["_engine"] = eng
        return eng

    def __getstate__(self):                       # joblib / pickle: drop device handles
        d = dict(self.__dict__)
        d.pop("_engine", None)
        d.pop("_dense_cache", None)
        d.pop("_dense_state", None)
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)

    def _pad_rows(self, rows):
        """list / array of (D,) vectors -> (K, Dp) device tensor."""
        eng = self._engine_()
        arr = np.asarray(rows, dtype=np.float64)
        return eng.pad(arr.reshape(-1, self.vspace_dim))

    # ---- norms -----------------------------------------------------------------------------------------------
    def H10norm(self, solutions: List[np.ndarray]):
        """sqrt(u^T A_1 u), A_1 = stiffness with a == 1 (reference :56-58); edge form, never NaN."""
        if "A_preassembled" in self.__dict__:                      # generic manager: u^T A4 u through GEMM + row dots
            st = self._dense_ops_()
            S = st["ops"].dev(np.asarray(solutions, dtype=np.float64).reshape(-1, self.vspace_dim))
            return np.sqrt(st["ops"].row_dots(S, st["ops"].gemm_nt(S, st["A4"])).cpu().numpy())
        eng = self._engine_()
        return eng.h10_norm(self._pad_rows(solutions)).cpu().numpy()

    @staticmethod
    def l2norm(solutions: List[np.ndarray]):
        """Euclidean norm of the coefficient vectors (reference :60-62)."""
        import ctypes as C
        import torch
        from .. import _lib
        if not torch.cuda.is_available():
            raise _lib.RomhcError("no CUDA device: the ROMHighContrast B200 path has no CPU fallback")
        X = torch.as_tensor(np.ascontiguousarray(np.asarray(solutions, dtype=np.float64)),
                            device=torch.device("cuda", torch.cuda.current_device()))
        X = X.reshape(X.shape[0], -1)
        out = torch.empty(X.shape[0], dtype=torch.float64, device=X.device)
        _lib.call("romhc_row_norms", C.c_void_p(X.data_ptr()), X.stride(0), X.shape[0], X.shape[1],
                  C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream(X.device).cuda_stream))
        return out.cpu().numpy()

    # ---- full-order solves -----------------------------------------------------------------------------------------
    def generate_solutions(self, a2try, *, return_stats=False):
        """Snapshots u(a) for every parameter (reference :64-68): batched GMG-PCG on the device."""
        _check_method(self.method)
        a = np.asarray(a2try, dtype=np.float64)
        if a.size == 0:
            return np.zeros((0, self.vspace_dim))
        a = a.reshape((-1,) + tuple(self.blocks_geometry))
        if "A_preassembled" in self.__dict__:                      # generic manager: K dense SPD solves of size D
            st = self._dense_ops_()
            U = st["ops"].solve_spd(st["ops"].dev(a.reshape(len(a), -1)), st["A"], st["b"]).cpu().numpy()
            return (U, None, None) if return_stats else U
        if not np.all(a > 0) or not np.all(np.isfinite(a)):
            raise np.linalg.LinAlgError("diffusion coefficients must be positive and finite")
        eng = self._engine_()
        U, iters, relres = eng.generate_solutions_host(a, return_stats=True)
        self.last_solver_report = {"iterations": iters, "relative_residual": relres}
        return (U, iters, relres) if return_stats else U

    def generate_riesz(self, x, norm="h10"):
        """Riesz representers of point evaluations, shape (m, D) (reference :70-86)."""
        if norm == "l2":
            return self._interpolation_matrix(x)
        elif norm.lower() == "h10":
            raise Exception("Not implemented.")                        # reference :78-79
        else:
            raise Exception("Not implemented.")                        # reference :86

    # ---- reduced Galerkin ---------------------------------------------------------------------------------------------
    def generate_riesz_h10(self, x, a=None):
        """Extension (the reference stubs this out at :78-84): Riesz representers of the point
        evaluations at x (m, 2) with respect to the H10 inner product u^T A_1 v, shape (m, D): row j solves
        A_1 w_j = l_j with l_j[i] = phi_i(x_j).  `a` (nrb, ncb) selects the energy inner product of A(a) instead.
        m batched GMG-PCG solves on the device with point-functional right-hand sides."""
        eng = self._engine_()
        L = self._interpolation_matrix(np.asarray(x, dtype=np.float64))            # (m, D) point functionals
        y = None
        if a is not None:
            y = eng.params(np.broadcast_to(np.asarray(a, dtype=np.float64), (L.shape[0],) + tuple(self.blocks_geometry)))
        w, iters, relres = eng.solve(y, rhs=eng.pad(L))
        self.last_solver_report = {"iterations": iters.cpu().numpy(), "relative_residual": relres.cpu().numpy()}
        return eng.unpad_host(w)

    def generate_fm_solutions(self, a: Union[np.ndarray, List[np.ndarray]], coefficients_rom: List[np.ndarray], *,
                              return_coefs=False):
        """Galerkin projection onto span(coefficients_rom) for every parameter (reference :88-106)."""
        if len(coefficients_rom) == 0:
            return np.zeros((len(a), self.vspace_dim))                  # reference :89-91
        _check_method(self.method)
        if "A_preassembled" in self.__dict__:                      # generic manager
            st = self._dense_ops_()
            ops = st["ops"]
            Phi = ops.dev(np.asarray(coefficients_rom, dtype=np.float64).reshape(len(coefficients_rom), -1))
            Ahat = self._dense_reduced_operators(Phi)
            bhat = ops.gemm_nt(Phi, st["b"].reshape(1, -1)).reshape(-1).contiguous()
            Cc = ops.solve_spd(ops.dev(np.asarray(a, dtype=np.float64).reshape(len(a), -1)), Ahat, bhat)
            return Cc.cpu().numpy() if return_coefs else ops.gemm_nn(Cc, Phi).cpu().numpy()
        eng = self._engine_()
        Phi = self._pad_rows(coefficients_rom)
        y = eng.params(np.asarray(a, dtype=np.float64).reshape((-1,) + tuple(self.blocks_geometry)))
        Ahat, bhat = eng.project_operators(Phi)                         # :93-103
        Cc = eng.reduced_solve(y, Ahat, bhat)                           # :104-105
        if return_coefs:
            return Cc.cpu().numpy()
        return eng.unpad_host(eng.gemm_nn(Cc, Phi))            # :106

    def project_solutions(self, solutions: List[np.ndarray], coefficients_rom: List[np.ndarray], *,
                          return_coefs=False):
        """H10-orthogonal projection of every solution onto span(coefficients_rom) (reference :108-139)."""
        if len(coefficients_rom) == 0:
            return np.zeros((len(solutions), self.vspace_dim))          # reference :109-111
        _check_method(self.method)
        if "A_preassembled" in self.__dict__:                      # generic manager
            st = self._dense_ops_()
            ops = st["ops"]
            Phi = ops.dev(np.asarray(coefficients_rom, dtype=np.float64).reshape(len(coefficients_rom), -1))
            S = ops.dev(np.asarray(solutions, dtype=np.float64).reshape(len(solutions), -1))
            B = ops.gemm_nt(S, ops.gemm_nn(Phi, st["A4"]))         # (K, n): B_km of :113-124, summed over the blocks
            Ahat = self._dense_reduced_operators(Phi)
            ones = ops.torch.ones(S.shape[0], Ahat.shape[0], dtype=ops.torch.float64, device=ops.device)
            Cc = ops.solve_spd(ones, Ahat, B)
            return Cc.cpu().numpy() if return_coefs else ops.gemm_nn(Cc, Phi).cpu().numpy()
        eng = self._engine_()
        import torch
        Phi = self._pad_rows(coefficients_rom)
        U = self._pad_rows(solutions)
        Cc = self._projection_coefficients_dev(eng, U, Phi)
        if return_coefs:
            return Cc.cpu().numpy()
        return eng.unpad_host(eng.gemm_nn(Cc, Phi))            # :139

    @staticmethod
    def _projection_coefficients_dev(eng, U_pad, Phi_pad):
        import torch
        W = eng.apply(None, Phi_pad)                                    # A_1 Phi^T
        B = eng.gemm_nt(U_pad, W, splitk=True)                          # (K, n): B_km of :113-124 (79 row tiles: split over the contraction)
        Ahat, _ = eng.project_operators(Phi_pad)                        # :125-133
        ones = torch.ones(U_pad.shape[0], eng.nb, dtype=torch.float64, device=eng.device)   # a = ones, :136
        return eng.reduced_solve(ones, Ahat, B)                         # :135-138

    def evaluate_solutions(self, points: np.ndarray, solutions: List[np.ndarray]) -> np.ndarray:
        raise Exception("Not implemented.")                             # reference :141-142


class SolutionsManagerFEM(SolutionsManager):
    """P1 FEM on the uniform right-triangle mesh of an (nrb x ncb) checkerboard (reference :145-244)."""

    def __init__(self, blocks_geometry: Tuple[int, int], N: int, num_cores=1, method="lsq"):
        nrb, ncb = blocks_geometry
        self.N = N
        self.x_domain = (-ncb / 2.0, ncb / 2.0)
        self.y_domain = (-nrb / 2.0, nrb / 2.0)
        self.nc_inner_vertices = ncb * self.N - 1
        self.nr_inner_vertices = nrb * self.N - 1
        self.nc_cells = ncb * self.N + 1
        self.nr_cells = nrb * self.N + 1
        self.points_c = np.linspace(*self.x_domain, self.nc_cells)
        self.points_r = np.linspace(*self.y_domain, self.nr_cells)
        self.vspace_dim = self.nc_inner_vertices * self.nr_inner_vertices
        self.blocks_geometry = (nrb, ncb)
        # f == 1 load vector: every interior entry collects area/6, area/3, area/3, area/6 from its four cells in the
        # reference's loop order (:177-185)
        area = (1 / self.N) * (1 / self.N)
        self.B_total = np.full(self.vspace_dim, ((area / 6 + area / 3) + area / 3) + area / 6)
        self.num_cores = num_cores
        super().__init__(None, None, num_cores=num_cores, method=method)      # matrix-free: no dense operators to hand over

    # the dense reference tensors exist only on demand and only where they fit (nobody outside L1 reads them)
    def _dense(self):
        cache = self.__dict__.get("_dense_cache")
        if cache is None:
            D, nb = self.vspace_dim, self.blocks_geometry[0] * self.blocks_geometry[1]
            if nb * D * D * 8 > 2 << 30:
                raise MemoryError(f"A_preassembled would need {nb * D * D * 8 / 2**30:.1f} GiB; the B200 path is "
                                  "matrix-free and never forms it")
            eng = self._engine_()
            eye = eng.pad(np.eye(D))
            import torch
            out = np.empty((nb, D, D))
            for q in range(nb):
                y = torch.zeros(D, nb, dtype=torch.float64, device=eng.device)
                y[:, q] = 1.0
                out[q] = eng.unpad(eng.apply(y, eye)).cpu().numpy()
            cache = out.reshape(self.blocks_geometry + (D, D))
            self.__dict__["_dense_cache"] = cache
        return cache

    @property
    def A_preassembled(self):
        return self._dense()

    @property
    def A_preassembled4h1_norm(self):
        return np.einsum("abij->ij", self._dense())

    def _interpolation_matrix(self, x):
        import ctypes as C
        import torch
        from .. import _lib
        eng = self._engine_()
        pts = eng.dev(np.asarray(x, dtype=np.float64).reshape(-1, 2))
        m = pts.shape[0]
        idx = torch.empty(m, 3, dtype=torch.int32, device=eng.device)
        w = torch.empty(m, 3, dtype=torch.float64, device=eng.device)
        _lib.check(eng.lib.romhc_interp_weights(eng.handle, C.c_void_p(pts.data_ptr()), m, C.c_void_p(idx.data_ptr()),
                                                C.c_void_p(w.data_ptr()), eng.stream()))
        idx, w = idx.cpu().numpy(), w.cpu().numpy()
        out = np.zeros((m, self.vspace_dim))
        r, c = idx // eng.P, idx % eng.P
        comp = (r - 1) * (eng.C - 1) + (c - 1)
        for j in range(m):
            for t in range(3):
                if idx[j, t] >= 0:
                    out[j, comp[j, t]] += w[j, t]
        return out

    def evaluate_solutions(self, points: np.ndarray, solutions: List[np.ndarray]) -> np.ndarray:
        """P1 interpolation of every solution at every point: (n, m) (reference :221-244)."""
        eng = self._engine_()
        pts = np.asarray(points, dtype=np.float64).reshape(-1, 2)
        if len(solutions) == 0:
            return np.zeros((0, len(pts)))
        return eng.evaluate(pts, self._pad_rows(solutions)).cpu().numpy()


class SolutionsManagerPolynomial(SolutionsManager):
    """The reference's 2x2-only spectral variant (:247-343) is never instantiated by the reference and is outside
    the accelerated path."""

    def __init__(self, lagrange_polynomials_degree):
        raise Exception("Not implemented.")
