"""CPU restatement of the reference's reduced-basis layer (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/src/lib/ReducedBasis.py (:14-29 helpers, :65-70 state estimation,
:112-139 greedy, :142-164 inf-split, :173-180 random, :189-200 PCA) and
/root/reference/src/lib/Estimators.py (:24-37).  `sm` is anything with the
`FEMOracle` / `SolutionsManager` method set.
"""
from __future__ import annotations

import numpy as np

INFINIT_A = 1e10   # ReducedBasis.py:11


def high_contrast_coefficient(a):
    """max over the block axes of every parameter (ReducedBasis.py:14-15)."""
    return np.array([np.max(c, axis=(-1, -2)) for c in a])


def sort_orthogonalize_base(a_selected, rb):
    """ReducedBasis.py:24-29, including the double application of `order`."""
    order = np.argsort(1 / a_selected)
    a_selected = a_selected[order]
    rb = rb[order, :]
    q, _ = np.linalg.qr(np.array(rb[order, :]).T)
    return a_selected, q.T


def greedy_build(sm, n, solutions2train, a2train, solutions2train_h1norm=1, greedy_for="galerkin",
                 return_trace=False):
    """Weak greedy with the TRUE H10 error (ReducedBasis.py:112-139).

    Returns (basis, a_list, selected_indices[, per-step error arrays])."""
    hc = high_contrast_coefficient(a2train)
    basis = np.empty((0, 0))
    basis_orth = basis.copy()
    a_selected, a, picked, trace = [], [], [], []
    for _ in range(n):
        if greedy_for == "galerkin":
            approx = sm.generate_fm_solutions(a=a2train, coefficients_rom=basis_orth)
        elif greedy_for == r"$H^1_0$":
            approx = sm.project_solutions(solutions=solutions2train, coefficients_rom=basis_orth)
        else:
            raise Exception(f"Not implemented greedy for {greedy_for}")
        err = sm.H10norm(approx - solutions2train) / solutions2train_h1norm
        ix = int(np.argmax(err))
        picked.append(ix)
        trace.append(err)
        row = np.reshape(solutions2train[ix], (1, -1))
        basis = row if len(basis) == 0 else np.concatenate((basis, row), axis=0)
        a.append(a2train[ix])
        a_selected = np.append(a_selected, np.ravel(hc[ix]))
        a_selected, basis_orth = sort_orthogonalize_base(a_selected, np.reshape(basis, (len(basis), -1)))
    if return_trace:
        return basis, a, picked, trace
    return basis, a, picked


def split_inf_solutions(solutions2train, a2train):
    """ReducedBasis.py:142-164 with only_one_block=False (both branches call it that way)."""
    a2train = np.asarray(a2train)
    num_hc = np.sum(a2train == INFINIT_A, axis=(-1, -2))
    chosen = np.ravel(np.where(num_hc != 0))
    free = np.ravel(np.where(num_hc == 0))
    return solutions2train[chosen], a2train[chosen], solutions2train[free], a2train[free]


def _starting_basis(solutions2train, a2train, add_inf_solutions):
    basis, a, s_free, a_free = split_inf_solutions(solutions2train, a2train)
    if not add_inf_solutions:
        basis = np.empty((0, np.shape(s_free)[1]))
        a = np.empty((0,) + np.shape(a_free)[1:])
    return basis, a, s_free, a_free


def random_build(n, solutions2train, a2train, add_inf_solutions=True, seed=42):
    basis, a, s_free, a_free = _starting_basis(solutions2train, a2train, add_inf_solutions)
    np.random.seed(seed)
    ix = np.random.choice(len(s_free), size=n, replace=False)
    return np.vstack((basis, s_free[ix]))[:n], np.vstack((a, a_free[ix]))[:n]


def pca_components(X, n):
    """Deterministic equivalent of sklearn `PCA(n_components=n, svd_solver="full").fit(X)`:
    centred thin SVD, rows of Vt sign-fixed so that the max-|entry| of each row is positive
    (sklearn.utils.extmath.svd_flip(u_based_decision=False)).  Returns components, singular values, mean."""
    X = np.asarray(X, dtype=np.float64)
    mean = X.mean(axis=0)
    _, s, vt = np.linalg.svd(X - mean, full_matrices=False)
    vt, s = vt[:n], s[:n]
    sign = np.sign(vt[np.arange(len(vt)), np.argmax(np.abs(vt), axis=1)])
    sign[sign == 0] = 1
    return vt * sign[:, None], s, mean


def pca_build(n, solutions2train, a2train, add_inf_solutions=True):
    basis, a, s_free, a_free = _starting_basis(solutions2train, a2train, add_inf_solutions)
    comps, s, _ = pca_components(s_free, n)
    return np.vstack((basis, comps))[:n], np.vstack((a, a_free))[:n], s


def state_estimation(sm, basis, measurement_points, measurements):
    """ReducedBasis.py:65-70: least squares on point values, returns (c (n,K), estimates (K,D))."""
    E = sm.evaluate_solutions(measurement_points, basis)          # (n, m)
    c = np.linalg.lstsq(E.T, np.asarray(measurements).T, rcond=-1)[0]
    return c, c.T @ np.array(basis)


def estimator_linear(c_values, a_basis):
    """Estimators.py:24-27."""
    return np.einsum("bi,b...->i...", c_values, np.asarray(a_basis, dtype=np.float64))


def estimator_inv(c_values, a_basis):
    """Estimators.py:30-37."""
    return 1.0 / np.einsum("bi,b...->i...", c_values, 1.0 / np.asarray(a_basis, dtype=np.float64))


def pbdw_correction(sm, points, measurements, v):
    """InverseProblemPipeline.ipynb cell 52: u* = v + z R^T - (v R) R^T with R = evaluate(points, eye(D)) (D, m)."""
    E = sm.interpolation_matrix(points)                            # (m, D) == R^T
    v = np.asarray(v, dtype=np.float64)
    z = np.asarray(measurements, dtype=np.float64)
    return v + np.asarray((E.T @ (z - np.asarray((E @ v.T).T)).T).T)
