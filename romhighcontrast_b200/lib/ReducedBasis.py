"""Drop-in for the reference's `src/lib/ReducedBasis.py` backed by the B200 kernels.

Same names, signatures and quirks as /root/reference/src/lib/ReducedBasis.py (line numbers cited).  Host-side
bookkeeping (selection order, the double permutation of `sort_orthogonalize_base`, numpy's legacy RNG, the
inf-split) is kept literally because the stored results depend on it; everything that touches (K, D) data runs
on the device and stays there between greedy steps.
"""
from logging import warning
from typing import List

import numpy as np

from .Estimators import EstimatorInv, EstimatorLinear
from .SolutionsManagers import SolutionsManager

try:                                    # tqdm is optional here; the reference uses it for a progress bar (:120)
    from tqdm import tqdm
except Exception:                       # pragma: no cover
    def tqdm(it, **kwargs):
        return it

INFINIT_A = 1e10  # reference :11


def get_high_contrast_coefficient(a):
    return np.array([np.max(coefs, axis=(-1, -2)) for coefs in a])           # reference :14-15


def orthonormalize_base(rb):
    q, r = np.linalg.qr(np.array(rb).T)                                     # reference :18-21 (host LAPACK, (D, n))
    rb = q.T
    return rb


def sort_orthogonalize_base(a_selected, rb):
    order = np.argsort(1 / a_selected)                                       # reference :24-29, `order` applied twice
    a_selected = a_selected[order]
    rb = rb[order, :]
    rb = orthonormalize_base(rb[order, :])
    return a_selected, rb


class BaseReducedBasis:
    def __init__(self):
        self.basis = None
        self.a = None
        self.inverse_parameter_estimator = None
        self.linear_parameter_estimator = None

    def build(self, **kwargs):
        raise Exception("Not implemented.")

    def set(self, basis, a):
        self.basis = basis
        self.a = a
        self.inverse_parameter_estimator = EstimatorInv(a)
        self.linear_parameter_estimator = EstimatorLinear(a)

    @property
    def dim(self):
        return np.shape(self.basis)[0]

    @property
    def ambient_space_dim(self):
        return np.shape(self.basis)[1]

    def __str__(self):
        return self.__class__.__name__

    def forward_modeling(self, sm: SolutionsManager, a: np.ndarray, **kwargs):
        return sm.generate_fm_solutions(a=a, coefficients_rom=self.basis, **kwargs)      # reference :59-60

    def projection(self, sm: SolutionsManager, true_solutions: np.ndarray, **kwargs):
        return sm.project_solutions(true_solutions, self.basis, **kwargs)                # reference :62-63

    def state_estimation(self, sm: SolutionsManager, measurement_points: np.ndarray, measurements: np.ndarray,
                         return_coefs=False, *, reconstruct=True):
        """Least-squares fit of the basis to point measurements (reference :65-70).

        The (m, n) collocation matrix is factorised once on the host (np.linalg.lstsq against the identity gives its
        pseudo-inverse with the same gelsd/rcond=-1 semantics); applying it to the K measurement vectors and
        reconstructing c^T basis are device GEMMs.  Additive keyword `reconstruct=False` returns only c (n, K)."""
        eng = sm._engine_()
        rb_evaluations_in_points = sm.evaluate_solutions(measurement_points, self.basis)           # (n, m)
        Z = np.asarray(measurements, dtype=np.float64)
        m = rb_evaluations_in_points.shape[1]
        pinv = np.linalg.lstsq(rb_evaluations_in_points.T, np.eye(m), rcond=-1)[0]                # (n, m)
        single = Z.ndim == 1           # one measurement vector (m,): lstsq(E^T, z) gives c (n,), c^T basis gives (D,)
        Zd = eng.dev(Z.reshape(-1, m))
        c_dev = eng.gemm_nt(eng.dev(pinv), Zd)                                                    # (n, K)
        if not reconstruct:            # large observation batches: (K, D) fields would not fit; coefficients only
            c = c_dev.cpu().numpy()
            return c[:, 0] if single else c
        basis_pad = sm._pad_rows(self.basis)
        solution_estimations = eng.unpad_host(eng.gemm_nn(c_dev.T.contiguous(), basis_pad))       # c^T basis
        c = c_dev.cpu().numpy()
        if single:
            c, solution_estimations = c[:, 0], solution_estimations[0]
        return (c, solution_estimations) if return_coefs else solution_estimations

    def parameter_estimation_inverse(self, c):
        return self.inverse_parameter_estimator.estimate_parameter(c_values=c)                     # reference :72-78

    def parameter_estimation_linear(self, c):
        return self.linear_parameter_estimator.estimate_parameter(c_values=c)                      # reference :80-86

    def __getitem__(self, item):
        rb = BaseReducedBasis()                                                                    # reference :88-92
        rb.set(basis=self.basis[item], a=self.a[item])
        return rb

    def orthonormalize(self):
        _, self.basis = sort_orthogonalize_base(                                                   # reference :94-98
            get_high_contrast_coefficient(self.a),
            np.reshape(self.basis, (-1, self.ambient_space_dim))
        )


GREEDY_FOR_H10 = r"$H^1_0$"
GREEDY_FOR_GALERKIN = "galerkin"


class ReducedBasisGreedy(BaseReducedBasis):
    def __init__(self, greedy_for=GREEDY_FOR_GALERKIN):
        self.greedy_for = greedy_for
        self.name = "Greedy " + self.greedy_for
        self.linestyle = "solid" if greedy_for == GREEDY_FOR_H10 else "dashed"
        super().__init__()

    def build(self, n: int, sm: SolutionsManager, solutions2train, a2train: List[np.ndarray] = (()),
              solutions2train_h1norm=1, **kwargs):
        """Weak greedy with the true H10 error (reference :112-139).

        The whole loop is device resident (romhighcontrast_b200/greedy_core.py): per round the reduced operators of the
        current orthonormal basis, K reduced solves (Galerkin) or K projections (H10), the fused error sweep
        || c Phi - u ||_{A_1} over the resident snapshots, an argmax with np.argmax semantics, a row gather and one
        Gram-Schmidt step; the host sees the selected indices once, at the end.  `solutions2train` may also be the padded
        device tensor (K, Dp) that `Engine.solve` returns (additive: snapshots -> build without crossing PCIe); `a2train`
        then may be a (K, nb) device tensor as well."""
        if self.greedy_for not in (GREEDY_FOR_H10, GREEDY_FOR_GALERKIN):
            raise Exception(f"Not implemented greedy for {self.greedy_for}, "
                            f"should be one of [{GREEDY_FOR_H10}, {GREEDY_FOR_GALERKIN}]")
        import torch
        from ..greedy_core import greedy_select
        eng = sm._engine_()
        dev_in = isinstance(solutions2train, torch.Tensor)
        if dev_in:
            U = solutions2train
            K = U.shape[0]
        else:
            solutions2train = np.asarray(solutions2train, dtype=np.float64)
            K = len(solutions2train)
            U = eng.pad(solutions2train)                                 # resident for the whole build
        y = a2train.reshape(K, -1).contiguous() if isinstance(a2train, torch.Tensor) else eng.params(np.asarray(a2train, dtype=np.float64))
        if isinstance(solutions2train_h1norm, torch.Tensor):
            norm = solutions2train_h1norm.to(eng.device).expand(K).contiguous()
        else:
            norm = eng.dev(np.broadcast_to(np.asarray(solutions2train_h1norm, dtype=np.float64), (K,)).copy())
        picked, max_errors, rows, params, _ = greedy_select(
            sm, eng, n, U, y, norm, self.greedy_for, progress=lambda it: tqdm(it, desc="Obtaining greedy basis."))
        self.selected_indices = picked
        self.max_errors = max_errors
        if dev_in:
            basis = eng.unpad(rows).cpu().numpy()
        else:
            basis = np.reshape(solutions2train[picked], (n, -1)) if n else np.empty((0, 0))
        if isinstance(a2train, torch.Tensor):
            pa = params.cpu().numpy().reshape((n,) + tuple(sm.blocks_geometry))
            a = [pa[i] for i in range(n)]
        else:
            a = [a2train[i] for i in picked]                             # :131 (the caller's own objects)
        super().set(basis=basis, a=a)
        return self


def get_inf_solutions_starting_basis(solutions2train, a2train, only_one_block=True):
    """Split off the snapshots with INFINIT_A blocks (reference :142-150)."""
    num_hc_blocks = np.sum(np.array(a2train) == INFINIT_A, axis=(-1, -2))
    chosen_ix = np.ravel(np.where(num_hc_blocks == 1 if only_one_block else num_hc_blocks != 0))
    free_ix = np.ravel(np.where(num_hc_blocks != 1 if only_one_block else num_hc_blocks == 0))
    return solutions2train[chosen_ix], a2train[chosen_ix], solutions2train[free_ix], a2train[free_ix]


def get_starting_basis(solutions2train, a2train, add_inf_solutions=True):
    basis, a, solutions2train, a2train = get_inf_solutions_starting_basis(solutions2train, a2train,   # :153-164
                                                                          only_one_block=False)
    if not add_inf_solutions:
        basis = np.empty((0, np.shape(solutions2train)[1]))
        a = np.empty((0,) + np.shape(a2train)[1:])
    return basis, a, solutions2train, a2train


class ReducedBasisRandom(BaseReducedBasis):
    def __init__(self, add_inf_solutions=True):
        self.add_inf_solutions = add_inf_solutions
        self.name = "Random" + (r" $\infty$" if add_inf_solutions else "")
        super().__init__()

    def build(self, n: int, sm: SolutionsManager, solutions2train, a2train: List[np.ndarray] = (()),
              solutions2train_h1norm=1, seed=42, **kwargs):
        """Host only: numpy's legacy global RNG must pick the same rows as the reference (:173-180)."""
        solutions2train, a2train = np.asarray(solutions2train), np.asarray(a2train)
        basis, a, solutions2train, a2train = get_starting_basis(solutions2train, a2train, self.add_inf_solutions)
        np.random.seed(seed)
        chosen_ix = np.random.choice(len(solutions2train), size=n, replace=False)
        super().set(basis=np.vstack((basis, solutions2train[chosen_ix]))[:n],
                    a=np.vstack((a, a2train[chosen_ix]))[:n])
        return self


class ReducedBasisPCA(BaseReducedBasis):
    GRAM_MAX_SNAPSHOTS = 32768      # pod_method="auto": Gram route up to here (G = 8.6 GB), Gram-free block Lanczos above

    def __init__(self, add_inf_solutions=True):
        self.add_inf_solutions = add_inf_solutions
        self.name = "PCA" + (r" $\infty$" if add_inf_solutions else "")
        super().__init__()

    def build(self, n: int, sm: SolutionsManager, solutions2train, a2train: List[np.ndarray] = (()),
              solutions2train_h1norm=1, add_inf_solutions=True, seed=42, pod_method="auto", **kwargs):
        """POD of the (inf-stripped) snapshots (reference :189-200).

        sklearn's PCA(n_components=n) is replaced by the method of snapshots on the device (column mean, centred
        Gram matrix on the fp64 tensor cores, top-n eigenpairs, back-projection) with sklearn's sign convention; it
        is deterministic, whereas the reference's randomized solver (random_state=None) moves by ~1e-8 run to run.
        pod_method (additive keyword; the reference's build swallows unknown keywords, :189): "gram" is the route above,
        "krylov" the Gram-free block-Lanczos route of pod.krylov_pca (same components and singular values to ~1e-15),
        "auto" (default) takes the Gram route up to GRAM_MAX_SNAPSHOTS snapshots and the Gram-free one beyond, where
        the K x K matrix (80 GB at K = 100 000) and its K^2 D flop stop being practical."""
        from ..pod import krylov_pca, pca_components
        if pod_method not in ("auto", "gram", "krylov"):
            raise ValueError(f"pod_method={pod_method!r} (expected 'auto', 'gram' or 'krylov')")
        solutions2train, a2train = np.asarray(solutions2train, dtype=np.float64), np.asarray(a2train)
        basis, a, solutions2train, a2train = get_starting_basis(solutions2train, a2train, self.add_inf_solutions)
        K, D = solutions2train.shape
        if not 0 <= n <= min(K, D):
            raise ValueError(f"n_components={n!r} must be between 0 and min(n_samples, n_features)={min(K, D)!r} "
                             "with svd_solver='full'")
        eng = sm._engine_()
        if pod_method == "auto":
            pod_method = "gram" if K <= self.GRAM_MAX_SNAPSHOTS else "krylov"
        comps_pad, sing, mean = (pca_components if pod_method == "gram" else krylov_pca)(
            eng, eng.pad(solutions2train), n, center_in_place=True)
        self.singular_values_ = sing.cpu().numpy()
        components = eng.unpad(comps_pad.contiguous()).cpu().numpy()
        super().set(basis=np.vstack((basis, components))[:n],
                    a=np.vstack((a, a2train))[:n])
        warning("PCA method has not been adapted for inverse parameter estimation, the a coefficients are not correct.")
        return self
