"""Notebook-level inverse-problem methods on the device.

The reference defines these only inside `src/notebooks/InverseProblemPipeline.ipynb` (cells 35, 44, 52), as functions of
a global `sm`.  Here they take `sm` explicitly; names and argument meaning follow the notebook:

* cell 52  `state_estimation_fitting_method_least_squares`, `pbdw_correction`, `state_estimation_fitting_method_pbdw`,
           `state_estimation_fitting_method_weighted_least_squares`, `polynomial_state_estimation_fitting_method_least_squares`
* cell 44  `inverse_christoffel_function`, `measurements_sampling_method_optimal`
* cell 35  `reduced_basis_generator_greedy` (the notebook's l2 / H10 greedy: least-squares residual, argmax)

Every O(K D) operation (point evaluation, reconstruction c^T basis, Riesz correction, residual norms) runs through the
C ABI (`romhc_evaluate`, `romhc_gemm_nn`, `romhc_energy_norm`, ...); the (m, n) collocation systems are factorised on
the host with the same LAPACK call the notebook uses (`np.linalg.lstsq(..., rcond=-1)`).
"""
from __future__ import annotations

import numpy as np
import torch

from .lib.ReducedBasis import orthonormalize_base


def _basis_array(reduced_basis):
    return np.ascontiguousarray(np.asarray(reduced_basis, dtype=np.float64)).reshape(len(reduced_basis), -1)


def _ls_coefficients(sm, measurement_points, measurements, basis, weights=None):
    """c (n, K) minimising || W (E^T c - z) || for every measurement vector z (rows of `measurements`)."""
    eng = sm._engine_()
    E = sm.evaluate_solutions(measurement_points, basis)                    # (n, m)
    m = E.shape[1]
    A = E.T if weights is None else E.T * weights[:, None]
    pinv = np.linalg.lstsq(A, np.eye(m), rcond=-1)[0]                         # (n, m): x = pinv @ (W z)
    if weights is not None:
        pinv = pinv * weights[None, :]
    Z = eng.dev(np.asarray(measurements, dtype=np.float64).reshape(-1, m))
    return eng.gemm_nt(eng.dev(pinv), Z)                                     # (n, K) device


def state_estimation_fitting_method_least_squares(sm, measurement_points, measurements, reduced_basis, **kwargs):
    """cell 52: least-squares fit of the basis to the point values, returns the (K, D) approximations."""
    eng = sm._engine_()
    basis = _basis_array(reduced_basis)
    c = _ls_coefficients(sm, measurement_points, measurements, basis)
    return eng.unpad_host(eng.gemm_nn(c.T.contiguous(), sm._pad_rows(basis)))


def pbdw_correction(sm, measurement_points, measurements, approximate_solutions, **kwargs):
    """cell 52: u* = v + z R^T - (v R) R^T with the l2 Riesz representers R = evaluate(points, eye(D)) (D, m).

    v R is the point evaluation of v, so no D x m matrix-matrix product is needed for it; the correction
    (z - v R) R^T is one (K, m) x (m, D) product with the (m, D) interpolation matrix."""
    eng = sm._engine_()
    V = sm._pad_rows(np.asarray(approximate_solutions, dtype=np.float64))
    Z = eng.dev(np.asarray(measurements, dtype=np.float64).reshape(V.shape[0], -1))
    Rt = sm._pad_rows(sm.generate_riesz(measurement_points, norm="l2"))     # (m, Dp)
    diff = (Z - eng.evaluate(measurement_points, V)).contiguous()            # (K, m)
    return eng.unpad_host(V + eng.gemm_nn(diff, Rt))


def state_estimation_fitting_method_pbdw(sm, measurement_points, measurements, reduced_basis, **kwargs):
    v = state_estimation_fitting_method_least_squares(sm, measurement_points, measurements, reduced_basis)
    return pbdw_correction(sm, measurement_points, measurements, v)


def inverse_christoffel_function(basis, sm, measurement_points):
    """cell 44: sum of squares of the ORTHONORMALISED basis functions at the points, shape (m,)."""
    q = orthonormalize_base(_basis_array(basis))
    vals = sm.evaluate_solutions(measurement_points, q)                      # (n, m)
    return np.sum(vals ** 2, axis=0)


def state_estimation_fitting_method_weighted_least_squares(sm, measurement_points, measurements, reduced_basis, **kwargs):
    """cell 52: the same least squares with rows weighted by 1 / inverse_christoffel_function."""
    eng = sm._engine_()
    basis = _basis_array(reduced_basis)
    weights = 1.0 / inverse_christoffel_function(basis, sm, measurement_points)
    c = _ls_coefficients(sm, measurement_points, measurements, basis, weights=weights)
    return eng.unpad_host(eng.gemm_nn(c.T.contiguous(), sm._pad_rows(basis)))


def polynomial_feature_terms(n, degree):
    """Index tuples of sklearn's PolynomialFeatures(degree, include_bias=False) in its output order: all degree-1 terms,
    then degree 2 in combinations_with_replacement order, ...; rows padded with -1 to `degree` factors."""
    from itertools import combinations_with_replacement
    terms = []
    for d in range(1, degree + 1):
        for comb in combinations_with_replacement(range(n), d):
            terms.append(list(comb) + [-1] * (degree - d))
    return np.asarray(terms, dtype=np.int32).reshape(-1, degree)


def polynomial_state_estimation_fitting_method_least_squares(sm, measurement_points, measurements, reduced_basis, degree=2,
                                                             **kwargs):
    """cell 52: v*(x) = sum_j c_j phi_j(x) + sum_jk d_jk phi_j(x) phi_k(x) (+ higher degrees): a linear regression of
    the measurements on the monomials of the basis values at the points (sklearn Pipeline(PolynomialFeatures(degree,
    include_bias=False), LinearRegression(fit_intercept=False)) in the notebook), predicted at every DOF.

    Host: the (m, F) feature matrix at the points and its least-squares solve (scipy.linalg.lstsq, as LinearRegression
    does).  Device: the (F, Dp) monomials of the basis at every DOF (`romhc_poly_features`) and the (K, F) x (F, Dp)
    product."""
    from scipy import linalg as sla
    eng = sm._engine_()
    basis = _basis_array(reduced_basis)
    n = basis.shape[0]
    E = sm.evaluate_solutions(measurement_points, basis)                    # (n, m)
    terms = polynomial_feature_terms(n, int(degree))
    X = np.ones((E.shape[1], len(terms)))
    for f, t in enumerate(terms):
        for j in t:
            if j >= 0:
                X[:, f] *= E[j]
    Z = np.asarray(measurements, dtype=np.float64).reshape(-1, E.shape[1])
    coef = sla.lstsq(X, Z.T)[0].T                                            # (K, F) == LinearRegression.coef_
    Fm = eng.poly_features(sm._pad_rows(basis), terms)                       # (F, Dp)
    return eng.unpad_host(eng.gemm_nn(eng.dev(coef), Fm))


def measurements_sampling_method_optimal(number_of_measures, xlim, ylim, basis, sm, seed=42, discretization=5, **kwargs):
    """cell 44: sample measurement points from the density given by the inverse Christoffel function on a grid."""
    np.random.seed(seed)
    n_per_dim = int(discretization * np.sqrt(number_of_measures))
    x, y = np.meshgrid(*[np.linspace(*xlim, num=n_per_dim), np.linspace(*ylim, num=n_per_dim)])
    pts = np.concatenate([x.reshape((-1, 1)), y.reshape((-1, 1))], axis=1)
    weights = inverse_christoffel_function(basis, sm, pts)
    weights /= np.sum(weights)
    return pts[np.random.choice(len(pts), size=number_of_measures, p=weights, replace=False)]


def reduced_basis_generator_greedy(sm, solutions_offline, number_of_reduced_base_elements, norm="l2"):
    """cell 35: start from the snapshot of largest norm; then repeatedly add the snapshot with the largest norm of its
    Euclidean least-squares residual against the current basis (np.argmax: first maximum).  Returns (basis, indices).

    The residual S - (S B^+) B is formed on the device; its norms come from the stencil-energy / l2 kernels."""
    if norm not in ("l2", "h10"):
        raise Exception(f"Norm {norm} not implemented.")
    eng = sm._engine_()
    S_host = np.asarray(solutions_offline, dtype=np.float64)
    S = sm._pad_rows(S_host)                                                  # (K, Dp) device
    f_norm = eng.l2_norm if norm == "l2" else eng.h10_norm
    picked = [int(np.argmax(f_norm(S).cpu().numpy()))]
    for _ in range(1, number_of_reduced_base_elements):
        B = S[picked].contiguous()                                            # (n, Dp)
        G = eng.gemm_nt(B, B).cpu().numpy()                                  # (n, n) Gram of the basis
        rhs = eng.gemm_nt(B, S)                                               # (n, K)
        Ginv = torch.as_tensor(np.ascontiguousarray(np.linalg.lstsq(G, np.eye(len(picked)), rcond=None)[0]), device=S.device)
        X = eng.gemm_nn(Ginv, rhs)                                            # (n, K) least-squares coefficients
        resid = S - eng.gemm_nn(X.T.contiguous(), B)
        picked.append(int(np.argmax(f_norm(resid).cpu().numpy())))
    return [S_host[i] for i in picked], picked


def observation_batch_estimation(sm, rb, measurement_points, a_observed, chunk=10000, pbdw=True, timings=None):
    """BASELINE configs[3] as one device pipeline: for every observed parameter (a batch of ~1e5) the snapshot is solved,
    measured at the m points, and from the measurements alone the state (least squares in the reduced space,
    ReducedBasis.py:65-70, optionally with the notebook's PBDW correction, cell 52) and the parameters (Estimators.py:24-37
    through `rb.parameter_estimation_inverse / _linear`) are estimated; the (K, D) fields live only chunk-wise on the
    device.  Returns a dict: measurements (K, m), coefficients c (n, K), a_inverse / a_linear (K, nrb, ncb), and the
    relative H10 errors of the state estimates against the true snapshots (`err_ls`, `err_pbdw`, (K,))."""
    import time
    eng = sm._engine_()
    pts = np.asarray(measurement_points, dtype=np.float64).reshape(-1, 2)
    a_obs = np.asarray(a_observed, dtype=np.float64)
    K, m = len(a_obs), len(pts)
    basis = _basis_array(rb.basis)
    n = basis.shape[0]
    Phi = sm._pad_rows(basis)
    E = sm.evaluate_solutions(pts, basis)                                     # (n, m)
    pinv = eng.dev(np.linalg.lstsq(E.T, np.eye(m), rcond=-1)[0])              # (n, m)
    Rt = sm._pad_rows(sm.generate_riesz(pts, norm="l2")) if pbdw else None    # (m, Dp)
    Z = torch.empty((K, m), dtype=torch.float64, device=eng.device)
    Cc = torch.empty((K, n), dtype=torch.float64, device=eng.device)
    err_ls = torch.empty(K, dtype=torch.float64, device=eng.device)
    err_pb = torch.empty(K, dtype=torch.float64, device=eng.device) if pbdw else None
    x = eng.empty(min(chunk, K), eng.Dp)
    tacc = {"solve_s": 0.0, "measure_s": 0.0, "state_ls_s": 0.0, "pbdw_s": 0.0}
    tick = (lambda: (torch.cuda.synchronize(), time.perf_counter())[1]) if timings is not None else (lambda: 0.0)
    for k0 in range(0, K, chunk):
        kc = min(chunk, K - k0)
        t0 = tick()
        xs, _, _ = eng.solve(eng.params(a_obs[k0:k0 + kc]), out=x[:kc])
        t1 = tick()
        z = eng.evaluate(pts, xs)                                             # (kc, m)
        Z[k0:k0 + kc] = z
        t2 = tick()
        c = eng.gemm_nt(z, pinv)                                              # (kc, n) = z pinv^T
        Cc[k0:k0 + kc] = c
        nrm = eng.h10_norm(xs)
        err_ls[k0:k0 + kc] = eng.error_norm(xs, c, Phi) / nrm
        t3 = tick()
        if pbdw:
            v = eng.gemm_nn(c, Phi)
            v = v + eng.gemm_nn((z - eng.evaluate(pts, v)).contiguous(), Rt)
            err_pb[k0:k0 + kc] = eng.h10_norm((v - xs).contiguous()) / nrm
        t4 = tick()
        tacc["solve_s"] += t1 - t0; tacc["measure_s"] += t2 - t1; tacc["state_ls_s"] += t3 - t2; tacc["pbdw_s"] += t4 - t3
    c_host = Cc.T.contiguous().cpu().numpy()                                  # (n, K) as np.linalg.lstsq returns it
    t5 = tick()
    out = {"measurements": Z.cpu().numpy(), "coefficients": c_host,
           "a_inverse": rb.parameter_estimation_inverse(c_host), "a_linear": rb.parameter_estimation_linear(c_host),
           "err_ls": err_ls.cpu().numpy(), "err_pbdw": err_pb.cpu().numpy() if pbdw else None}
    if timings is not None:
        tacc["parameter_estimation_s"] = tick() - t5
        timings.update(tacc)
    return out
