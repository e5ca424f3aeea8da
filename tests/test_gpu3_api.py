"""GPU parity tests of the drop-in `src.lib` API against vectors produced by the unmodified reference
(tests/golden, see oracle/gen_golden.py).  Import spellings are the reference's own."""
import pickle

import numpy as np
import pytest

from conftest import golden, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sm22():
    from lib.SolutionsManagers import SolutionsManagerFEM      # spelling of the reference's tests
    return SolutionsManagerFEM(blocks_geometry=(2, 2), N=10, num_cores=1, method="lsq")


def test_riesz_like_reference_test(sm22):
    """/root/reference/src/tests/test_Functions/test_SolutionsManager.py: shapes; the h10 variant raises in the
    reference itself (SolutionsManagers.py:78-79) and must raise the same here."""
    assert np.shape(sm22.generate_riesz([(0, 0)], norm="l2")) == (1, sm22.vspace_dim)
    with pytest.raises(Exception, match="Not implemented"):
        sm22.generate_riesz([(0, 0)], norm="h10")
    with pytest.raises(Exception, match="Not implemented"):
        sm22.generate_riesz([(0, 0)], norm="sobolev")


def test_attributes_match_reference(sm22):
    g = golden("g1_assembly_2x2_N10.npz")
    assert sm22.vspace_dim == 361 and sm22.blocks_geometry == (2, 2)
    np.testing.assert_array_equal(sm22.B_total, g["B_total"])
    np.testing.assert_array_equal(sm22.points_c, g["points_c"])
    np.testing.assert_array_equal(sm22.points_r, g["points_r"])
    assert sm22.x_domain == (-1.0, 1.0) and sm22.y_domain == (-1.0, 1.0)
    assert (sm22.nc_cells, sm22.nr_cells, sm22.nc_inner_vertices, sm22.nr_inner_vertices) == (21, 21, 19, 19)
    assert str(sm22) == "SolutionsManagerFEM"
    # the dense tensors stay available where they fit
    from src.lib.SolutionsManagers import SolutionsManagerFEM
    sm = SolutionsManagerFEM((3, 2), 4)
    g = golden("g1_assembly_3x2_N4.npz")
    np.testing.assert_allclose(sm.A_preassembled, g["A_pre"], atol=1e-15)
    np.testing.assert_allclose(sm.A_preassembled4h1_norm, g["A1"], atol=1e-14)


@pytest.mark.parametrize("method", ["lsq", "lsqsparse", "ridge", "LSQ"])
def test_generate_solutions_and_norms(method):
    from src.lib.SolutionsManagers import SolutionsManagerFEM
    g = golden("g2_solve_2x2_N10.npz")
    sm = SolutionsManagerFEM((2, 2), N=10, method=method)
    U = sm.generate_solutions(g["y"])
    assert U.shape == (5, 361) and U.flags["C_CONTIGUOUS"] and U.dtype == np.float64
    assert relerr(U, g["U_lsq"]) < 1e-9 and relerr(U, g["U_lsqsparse"]) < 1e-9
    np.testing.assert_allclose(sm.H10norm(U), g["h10"], rtol=1e-9)
    np.testing.assert_allclose(sm.l2norm(U), g["l2"], rtol=1e-9)
    np.testing.assert_allclose(SolutionsManagerFEM.l2norm(list(U)), g["l2"], rtol=1e-9)   # staticmethod, list input
    # list-of-arrays and integer inputs are accepted like the reference's np.einsum does
    Ui = sm.generate_solutions([np.array([[1, 7], [100, 1000]])])
    assert Ui.shape == (1, 361)
    assert sm.generate_solutions(np.empty((0, 2, 2))).shape == (0, 361)


def test_unknown_method_raises_like_reference():
    from src.lib.SolutionsManagers import SolutionsManagerFEM
    sm = SolutionsManagerFEM((2, 2), N=4, method="cholesky")
    with pytest.raises(Exception, match="Method cholesky Not implemented."):
        sm.generate_solutions(np.ones((1, 2, 2)))


def test_reduced_galerkin_projection_evaluation():
    from src.lib.SolutionsManagers import SolutionsManagerFEM
    g = golden("g3_reduced_3x2_N4.npz")
    sm = SolutionsManagerFEM((3, 2), N=4)
    for tag in ("", "_snap"):
        Phi = g["Phi" + tag]
        fm = sm.generate_fm_solutions(g["y"], Phi)
        assert relerr(fm, g["fm" + tag]) < 1e-9
        assert relerr(sm.generate_fm_solutions(list(g["y"]), list(Phi)), g["fm" + tag]) < 1e-9
        assert relerr(sm.project_solutions(g["U"], Phi), g["proj" + tag]) < 1e-9
        c = sm.generate_fm_solutions(g["y"], Phi, return_coefs=True)
        assert relerr(c @ Phi, g["fm" + tag]) < 1e-9
    z = sm.generate_fm_solutions(g["y"], np.empty((0, 0)))
    np.testing.assert_array_equal(z, g["fm_empty"])
    np.testing.assert_array_equal(sm.project_solutions(g["U"], []), np.zeros_like(g["U"]))
    np.testing.assert_allclose(sm.evaluate_solutions(g["pts"], g["U"]), g["ev"], rtol=1e-12, atol=1e-18)
    np.testing.assert_allclose(sm.evaluate_solutions(g["nodes"], list(g["U"][:2])), g["U"][:2], rtol=1e-12, atol=1e-18)
    np.testing.assert_allclose(sm.generate_riesz(g["pts"][:3], norm="l2"), g["riesz_l2"], atol=1e-14)


@pytest.mark.parametrize("N", [10, 32])
def test_greedy_selects_the_reference_indices(N):
    from src.lib.SolutionsManagers import SolutionsManagerFEM
    from src.lib.ReducedBasis import ReducedBasisGreedy, GREEDY_FOR_GALERKIN, GREEDY_FOR_H10
    g = golden(f"g4_greedy_2x2_N{N}.npz")
    sm = SolutionsManagerFEM((2, 2), N=N, method="lsqsparse")
    U = g["U"] if "U" in g else sm.generate_solutions(g["y"])
    h1 = sm.H10norm(U)
    np.testing.assert_allclose(h1, g["h1"], rtol=1e-9)
    for tag, crit in (("gal", GREEDY_FOR_GALERKIN), ("h10", GREEDY_FOR_H10)):
        rb = ReducedBasisGreedy(greedy_for=crit).build(n=10, sm=sm, solutions2train=U, a2train=g["y"],
                                                       optim_method="lsq", solutions2train_h1norm=h1)
        assert rb.selected_indices == list(g[f"idx_{tag}"]), (tag, rb.selected_indices)
        np.testing.assert_array_equal(np.array(rb.a), g[f"a_{tag}"])
        np.testing.assert_array_equal(rb.basis, U[rb.selected_indices])
        rb.orthonormalize()
        fm = rb.forward_modeling(sm, g["y"])
        pj = rb.projection(sm, U)
        np.testing.assert_allclose(sm.H10norm(fm - U) / h1, g[f"fm_err_{tag}"], rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(sm.H10norm(pj - U) / h1, g[f"pj_err_{tag}"], rtol=1e-7, atol=1e-9)
        if N == 10:
            assert relerr(np.abs(rb.basis), np.abs(g[f"basis_orth_{tag}"])) < 1e-9
    if N == 10:
        rb = ReducedBasisGreedy().build(n=4, sm=sm, solutions2train=U, a2train=g["y"])
        assert rb.selected_indices == list(g["idx_gal_unnormalised"])
    assert ReducedBasisGreedy(GREEDY_FOR_H10).linestyle == "solid" and ReducedBasisGreedy().name == "Greedy galerkin"
    with pytest.raises(Exception, match="Not implemented greedy for"):
        ReducedBasisGreedy(greedy_for="l2").build(n=1, sm=sm, solutions2train=U, a2train=g["y"])


def test_pca_random_builders_state_and_parameter_estimation(sm22):
    from src.lib.ReducedBasis import ReducedBasisPCA, ReducedBasisRandom, ReducedBasisGreedy, BaseReducedBasis
    g = golden("g5_builders_2x2_N10.npz")
    U, y = g["U"], g["y"]
    pca = ReducedBasisPCA().build(n=10, sm=sm22, solutions2train=U, a2train=y)
    np.testing.assert_allclose(pca.singular_values_, g["pca_full_singular_values"], rtol=1e-9)
    ratio = g["pca_full_singular_values"] / g["pca_full_singular_values"][0]
    for i in range(10):
        if ratio[i] > 1e-6:
            assert np.abs(pca.basis[i] - g["pca_full_components"][i]).max() < 1e-7, i
    assert pca.name == r"PCA $\infty$" and np.shape(pca.a) == (10, 2, 2)
    rnd = ReducedBasisRandom().build(n=10, sm=sm22, solutions2train=U, a2train=y, seed=42)
    np.testing.assert_array_equal(rnd.basis, g["random_basis"])
    np.testing.assert_array_equal(rnd.a, g["random_a"])
    rbg = ReducedBasisGreedy().build(n=6, sm=sm22, solutions2train=U, a2train=y, solutions2train_h1norm=sm22.H10norm(U))
    np.testing.assert_array_equal(rbg.basis, g["greedy6_basis"])
    c, est = rbg.state_estimation(sm22, g["points"], g["measurements"], return_coefs=True)
    assert c.shape == g["se_c"].shape and relerr(est, g["se_est"]) < 1e-8 and relerr(c, g["se_c"]) < 1e-6
    est2 = rbg.state_estimation(sm22, g["points"], g["measurements"])
    np.testing.assert_array_equal(est2, est)
    np.testing.assert_allclose(rbg.parameter_estimation_inverse(g["se_c"]), g["inv"], rtol=1e-12)
    np.testing.assert_allclose(rbg.parameter_estimation_linear(g["se_c"]), g["lin"], rtol=1e-12)
    sub = rbg[:3]
    assert type(sub) is BaseReducedBasis and sub.dim == int(g["sliced_dim"]) and sub.ambient_space_dim == 361
    with pytest.raises(Exception, match="Not implemented"):
        BaseReducedBasis().build()
    # builders and managers survive pickling (joblib checkpoint path of experiments/HighContrast.py:93-96)
    rb2, sm2 = pickle.loads(pickle.dumps((rbg, sm22)))
    np.testing.assert_array_equal(rb2.basis, rbg.basis)
    assert relerr(sm2.H10norm(U[:3]), sm22.H10norm(U[:3])) == 0.0


def test_inf_split_builders():
    from src.lib.SolutionsManagers import SolutionsManagerFEM
    from src.lib.ReducedBasis import ReducedBasisRandom, ReducedBasisGreedy, ReducedBasisPCA, INFINIT_A
    assert INFINIT_A == 1e10
    g = golden("g7_inf_2x2_N6.npz")
    sm = SolutionsManagerFEM((2, 2), N=6)
    U = sm.generate_solutions(g["a"])
    assert relerr(U[3:], g["U"][3:]) < 1e-9          # finite-contrast rows; the 1e10 rows agree to the reference's own accuracy
    assert relerr(U[:3], g["U"][:3]) < 1e-4
    r = ReducedBasisRandom(True).build(n=5, sm=sm, solutions2train=g["U"], a2train=g["a"])
    np.testing.assert_array_equal(r.basis, g["rand_inf_basis"])
    np.testing.assert_array_equal(r.a, g["rand_inf_a"])
    r = ReducedBasisRandom(False).build(n=3, sm=sm, solutions2train=g["U"], a2train=g["a"])
    np.testing.assert_array_equal(r.basis, g["rand_noinf_basis"])
    gr = ReducedBasisGreedy().build(n=5, sm=sm, solutions2train=g["U"], a2train=g["a"],
                                    solutions2train_h1norm=sm.H10norm(g["U"]))
    assert gr.selected_indices == list(g["greedy_idx"])
    p = ReducedBasisPCA(True).build(n=4, sm=sm, solutions2train=g["U"], a2train=g["a"])
    np.testing.assert_array_equal(p.basis[:3], g["U"][:3])
    with pytest.raises(ValueError):
        ReducedBasisPCA(True).build(n=40, sm=sm, solutions2train=g["U"], a2train=g["a"])


def test_pbdw_and_large_batch_online_stage():
    """state estimation on many observations + PBDW correction (InverseProblemPipeline.ipynb cell 52)."""
    from src.lib.SolutionsManagers import SolutionsManagerFEM
    from src.lib.ReducedBasis import ReducedBasisGreedy
    from oracle import FEMOracle, pbdw_correction, state_estimation
    geo, N, K = (4, 4), 8, 300
    sm = SolutionsManagerFEM(geo, N)
    o = FEMOracle(geo, N)
    y = 10 ** np.random.default_rng(44).uniform(0, 6, (K,) + geo)
    U = sm.generate_solutions(y)
    rb = ReducedBasisGreedy().build(n=8, sm=sm, solutions2train=U[:100], a2train=y[:100],
                                    solutions2train_h1norm=sm.H10norm(U[:100]))
    pts = np.random.default_rng(1).uniform(low=[-2, -2], high=[2, 2], size=(50, 2))
    Z = sm.evaluate_solutions(pts, U)
    np.testing.assert_allclose(Z, o.evaluate_solutions(pts, U), rtol=1e-12, atol=1e-16)
    c, est = rb.state_estimation(sm, pts, Z, return_coefs=True)
    co, esto = state_estimation(o, rb.basis, pts, Z)
    assert relerr(est, esto) < 1e-7
    # a single 1-D measurement vector: shapes (n,) and (D,) as np.linalg.lstsq(E.T, z.T) gives them (ReducedBasis.py:65-70)
    c1, est1 = rb.state_estimation(sm, pts, Z[3], return_coefs=True)
    assert c1.shape == (rb.dim,) and est1.shape == (sm.vspace_dim,)
    np.testing.assert_allclose(c1, c[:, 3], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(est1, est[3], rtol=1e-12, atol=1e-14)
    R = sm.generate_riesz(pts, norm="l2")                                  # (m, D)
    corrected = est + (Z - est @ R.T) @ R
    assert relerr(corrected, pbdw_correction(o, pts, Z, esto)) < 1e-7
    # the corrected state interpolates the data up to the conditioning of R R^T
    inv = rb.parameter_estimation_inverse(c)
    assert inv.shape == (K, 4, 4)


def test_riesz_h10_extension():
    """H10 Riesz representers (the reference raises 'Not implemented.'): <w_j, v>_{A_1} = v(x_j) for every v, and the
    representers equal the sparse direct solution of A_1 w = l_j; the energy variant uses A(a)."""
    from lib.SolutionsManagers import SolutionsManagerFEM
    from oracle import FEMOracle
    import scipy.sparse.linalg as spla
    geo, N = (3, 2), 16
    sm = SolutionsManagerFEM(geo, N)
    o = FEMOracle(geo, N)
    rng = np.random.default_rng(3)
    pts = np.column_stack([rng.uniform(*sm.x_domain, 7), rng.uniform(*sm.y_domain, 7)])
    with pytest.raises(Exception, match="Not implemented"):
        sm.generate_riesz(pts, norm="h10")                      # the reference-compatible entry point is untouched
    W = sm.generate_riesz_h10(pts)
    assert W.shape == (7, sm.vspace_dim)
    V = rng.standard_normal((5, sm.vspace_dim))
    lhs = V @ (o.A1 @ W.T)                                      # (5, 7) inner products
    rhs = sm.evaluate_solutions(pts, V)                         # (5, 7) point values
    assert relerr(lhs, rhs) < 1e-10
    Lmat = sm.generate_riesz(pts, norm="l2")                    # (7, D) point functionals
    Wo = spla.splu(o.A1.tocsc()).solve(Lmat.T).T
    assert relerr(W, Wo) < 1e-9
    a = 10 ** rng.uniform(0, 4, geo)
    Wa = sm.generate_riesz_h10(pts, a=a)
    Wao = spla.splu(o.matrix(a).tocsc()).solve(Lmat.T).T
    assert relerr(Wa, Wao) < 1e-9


def test_notebook_inverse_methods():
    """romhighcontrast_b200.inverse vs direct numpy restatements of InverseProblemPipeline.ipynb cells 35 / 44 / 52 run
    on the CPU oracle (same snapshots, same points)."""
    from lib.SolutionsManagers import SolutionsManagerFEM
    from lib.ReducedBasis import orthonormalize_base
    from oracle import FEMOracle
    from romhighcontrast_b200 import inverse as inv
    geo, N, K, n, m = (2, 2), 10, 60, 5, 30
    sm = SolutionsManagerFEM(geo, N)
    o = FEMOracle(geo, N)
    rng = np.random.default_rng(12)
    y = 10 ** rng.uniform(0, 3, (K,) + geo)
    U = sm.generate_solutions(y)
    pts = np.column_stack([rng.uniform(*sm.x_domain, m), rng.uniform(*sm.y_domain, m)])
    Z = o.evaluate_solutions(pts, U)
    # cell 35: notebook greedy, both norms
    for norm, f in (("l2", o.l2norm), ("h10", o.H10norm)):
        basis = [U[int(np.argmax(f(U), axis=0))]]
        idx = [int(np.argmax(f(U), axis=0))]
        for _ in range(1, n):
            x = np.linalg.lstsq(np.transpose(basis), np.transpose(U), rcond=None)[0]
            nxt = int(np.argmax(f((np.transpose(U) - np.transpose(basis) @ x).T)))
            basis.append(U[nxt]); idx.append(nxt)
        got, gidx = inv.reduced_basis_generator_greedy(sm, U, n, norm=norm)
        assert gidx == idx, (norm, gidx, idx)
        np.testing.assert_array_equal(np.asarray(got), np.asarray(basis))
    rb = basis
    # cell 52: least squares, PBDW, weighted least squares
    E = o.evaluate_solutions(pts, rb)
    ls = (np.linalg.lstsq(E.T, Z.T, rcond=-1)[0]).T @ np.array(rb)
    assert relerr(inv.state_estimation_fitting_method_least_squares(sm, pts, Z, rb), ls) < 1e-9
    R = o.evaluate_solutions(points=pts, solutions=np.eye(o.vspace_dim))        # (D, m)
    pb = ls + Z @ R.T - (ls @ R) @ R.T
    assert relerr(inv.pbdw_correction(sm, pts, Z, ls), pb) < 1e-9
    assert relerr(inv.state_estimation_fitting_method_pbdw(sm, pts, Z, rb), pb) < 1e-9
    w = np.sum(o.evaluate_solutions(pts, orthonormalize_base(np.asarray(rb))) ** 2, axis=0)
    np.testing.assert_allclose(inv.inverse_christoffel_function(rb, sm, pts), w, rtol=1e-10)
    wl = 1 / w
    wls = (np.linalg.lstsq(E.T * wl[:, None], Z.T * wl[:, None], rcond=-1)[0]).T @ np.array(rb)
    assert relerr(inv.state_estimation_fitting_method_weighted_least_squares(sm, pts, Z, rb), wls) < 1e-9
    # cell 52: polynomial least squares against the notebook's own sklearn pipeline on the oracle's evaluations
    from sklearn.linear_model import LinearRegression
    from sklearn.pipeline import Pipeline
    from sklearn.preprocessing import PolynomialFeatures
    for degree, nb_ in ((2, 5), (3, 3), (1, 5)):
        model = Pipeline([('PF', PolynomialFeatures(degree=degree, include_bias=False)), ('LR', LinearRegression(fit_intercept=False))])
        model.fit(o.evaluate_solutions(pts, rb[:nb_]).T, Z.T)
        ref_poly = model.predict(np.array(rb[:nb_]).T).T
        got_poly = inv.polynomial_state_estimation_fitting_method_least_squares(sm, pts, Z, rb[:nb_], degree=degree)
        assert relerr(got_poly, ref_poly) < 1e-8, (degree, relerr(got_poly, ref_poly))
    # cell 44: optimal sampling draws the same points (same seed, same density up to rounding)
    p1 = inv.measurements_sampling_method_optimal(8, sm.x_domain, sm.y_domain, rb, sm, seed=3)
    np.random.seed(3)
    npd = int(5 * np.sqrt(8))
    gx, gy = np.meshgrid(np.linspace(*sm.x_domain, num=npd), np.linspace(*sm.y_domain, num=npd))
    gp = np.concatenate([gx.reshape((-1, 1)), gy.reshape((-1, 1))], axis=1)
    ww = np.sum(o.evaluate_solutions(gp, orthonormalize_base(np.asarray(rb))) ** 2, axis=0)
    ww /= ww.sum()
    p2 = gp[np.random.choice(len(gp), size=8, p=ww, replace=False)]
    np.testing.assert_allclose(p1, p2)


def test_host_layout_transfers_match_device_path():
    """romhc_pack_host / romhc_unpack_host (pipelined pageable transfers behind Engine.pad / unpad_host) reproduce the
    plain copy + layout kernels bit for bit: several chunks, a ragged last chunk, pageable and pinned host memory."""
    import torch
    from romhighcontrast_b200.engine import Engine
    eng = Engine((3, 3), 43)                              # D = 16 384: 1024 rows per 128 MB chunk
    K = 2500
    rng = np.random.default_rng(0)
    U = rng.standard_normal((K, eng.D))
    ref = eng.empty(K, eng.Dp)
    from romhighcontrast_b200 import _lib
    Ud = torch.as_tensor(U).to(eng.device)
    _lib.check(eng.lib.romhc_pack(eng.handle, Ud.data_ptr(), ref.data_ptr(), K, eng.stream()))
    got = eng.pad(U)                                      # pageable numpy -> pipelined path
    assert torch.equal(got, ref)
    Upin = torch.as_tensor(U).pin_memory().numpy()
    assert torch.equal(eng.pad(Upin), ref)                # pinned: direct DMA
    back = eng.unpad_host(ref)
    np.testing.assert_array_equal(back, U)
    out_pin = torch.empty((K, eng.D), dtype=torch.float64, pin_memory=True).numpy()
    np.testing.assert_array_equal(eng.unpad_host(ref, out=out_pin), U)
    # small arrays take the plain path
    np.testing.assert_array_equal(eng.unpad_host(ref[:3]), U[:3])
    assert torch.equal(eng.pad(U[:3]), ref[:3])
    # twice in a row (slot reuse across calls)
    assert torch.equal(eng.pad(U * 2.0), ref * 2.0)


def test_generic_dense_manager_and_galerkin_on_reference_operators():
    """SolutionsManager(A_preassembled, B_total) (reference :43-139) and galerkin() (:17-40) on the reference's own dense
    operators (golden g1, D = 77: the blocked dense Cholesky path, n > 64) against the reference's outputs (golden g2 / g3);
    a D = 961 system against the matrix-free FEM path; a small reduced system (n = 5: the quad kernel)."""
    from src.lib.SolutionsManagers import SolutionsManager, SolutionsManagerFEM, galerkin
    g1, g2, g3 = golden("g1_assembly_3x2_N4.npz"), golden("g2_solve_3x2_N4.npz"), golden("g3_reduced_3x2_N4.npz")
    sm = SolutionsManager(g1["A_pre"], g1["B_total"], num_cores=1, method="lsq")
    assert sm.vspace_dim == 77 and tuple(sm.blocks_geometry) == (3, 2) and str(sm) == "SolutionsManager"
    np.testing.assert_array_equal(sm.A_preassembled4h1_norm, np.einsum("abij->ij", g1["A_pre"]))
    U = sm.generate_solutions(g2["y"])
    assert U.shape == (6, 77) and relerr(U, g2["U_lsq"]) < 1e-9
    np.testing.assert_allclose(sm.H10norm(g2["U_lsq"]), g2["h10"], rtol=1e-9)
    np.testing.assert_allclose(sm.l2norm(g2["U_lsq"]), g2["l2"], rtol=1e-12)
    assert relerr(sm.generate_fm_solutions(g3["y"], g3["Phi"]), g3["fm"]) < 1e-9
    assert relerr(sm.project_solutions(g3["U"], g3["Phi"]), g3["proj"]) < 1e-9
    assert relerr(sm.generate_fm_solutions(g3["y"], list(g3["Phi_snap"])), g3["fm_snap"]) < 1e-9
    np.testing.assert_array_equal(sm.generate_fm_solutions(g3["y"], []), g3["fm_empty"])
    with pytest.raises(Exception, match="Not implemented"):
        sm.evaluate_solutions(g3["pts"], g3["U"])
    import pickle
    sm_p = pickle.loads(pickle.dumps(sm))
    assert relerr(sm_p.generate_solutions(g2["y"][:2]), g2["U_lsq"][:2]) < 1e-9
    # galerkin(): one dense system, every reference method spelling, n = 77 (> 64: blocked Cholesky)
    for method in ("lsq", "lsqsparse", "ridge", "LSQ"):
        c = galerkin(g2["y"][1], g1["B_total"], g1["A_pre"], method=method)
        assert c.shape == (77,) and relerr(c, g2["U_lsq"][1]) < 1e-9
    with pytest.raises(Exception, match="Method cholesky Not implemented."):
        galerkin(g2["y"][1], g1["B_total"], g1["A_pre"], method="cholesky")
    # reduced operators of the reference's generate_fm_solutions (:93-103) through galerkin(), n = 5
    Phi = g3["Phi"]
    A_kl = np.einsum("...jk,dk->...jd", np.einsum("...jk,dj->...dk", g1["A_pre"], Phi), Phi)
    c5 = galerkin(g3["y"][2], Phi @ g1["B_total"], A_kl)
    assert relerr(c5 @ Phi, g3["fm"][2]) < 1e-9
    # a larger dense system (D = 961, 31 block columns) against the matrix-free solver
    fem = SolutionsManagerFEM((2, 2), 16)
    y = 10 ** np.random.default_rng(5).uniform(0, 4, (3, 2, 2))
    Uf = fem.generate_solutions(y)
    gen = SolutionsManager(fem.A_preassembled, fem.B_total)
    assert relerr(gen.generate_solutions(y), Uf) < 1e-9
    assert relerr(galerkin(y[0], fem.B_total, fem.A_preassembled), Uf[0]) < 1e-9
    np.testing.assert_allclose(gen.H10norm(Uf), fem.H10norm(Uf), rtol=1e-10)
    # not positive definite -> LinAlgError, as scipy.linalg.solve(assume_a='pos') raises
    with pytest.raises(np.linalg.LinAlgError):
        galerkin(-np.ones((3, 2)), g1["B_total"], g1["A_pre"])


def test_large_bases_n_above_32():
    """the basis dimension was silently capped at 32 (projection kernel, gemm_tn), 64 (reduced solve).
    Every builder and every online call now takes any n: greedy n = 40 against the oracle's greedy (identical indices while
    the error gap exceeds 1e-9), PCA n = 40 (back-projection in row blocks of 32), forward modelling / projection / state
    estimation with n = 33 and n = 70 (reduced operators through stencil apply + split-K DMMA, blocked dense Cholesky)."""
    from src.lib.ReducedBasis import ReducedBasisGreedy, ReducedBasisPCA, BaseReducedBasis, GREEDY_FOR_GALERKIN, GREEDY_FOR_H10
    from src.lib.SolutionsManagers import SolutionsManagerFEM
    from oracle import FEMOracle
    from oracle.rb import greedy_build, pca_components
    geo, N, K = (3, 3), 8, 160
    sm = SolutionsManagerFEM(geo, N)
    o = FEMOracle(geo, N)
    y = 10 ** np.random.default_rng(8).uniform(0, 4, (K,) + geo)
    U = sm.generate_solutions(y)
    h1 = sm.H10norm(U)
    n = 40
    for crit in (GREEDY_FOR_GALERKIN, GREEDY_FOR_H10):
        rb = ReducedBasisGreedy(greedy_for=crit).build(n=n, sm=sm, solutions2train=U, a2train=y, solutions2train_h1norm=h1)
        _, _, picked, trace = greedy_build(o, n, U, y, o.H10norm(U), greedy_for=crit, return_trace=True)
        agree = 0
        for step, (a_, b_) in enumerate(zip(picked, rb.selected_indices)):
            top = np.sort(trace[step])[-2:]
            gap = (top[1] - top[0]) / top[1]
            if a_ != b_:
                assert gap < 1e-9 or top[1] < 1e-7, (crit, step, a_, b_, gap, top[1])   # rounding-level ties / errors at the noise floor
                break
            agree += 1
        assert agree >= 33, (crit, agree)
        assert np.asarray(rb.basis).shape == (n, sm.vspace_dim)
    rbp = ReducedBasisPCA().build(n=n, sm=sm, solutions2train=U, a2train=y)
    comps, so, _ = pca_components(U, n)
    np.testing.assert_allclose(np.asarray(rbp.singular_values_), so, rtol=1e-9, atol=1e-9 * so[0])
    # online stage with n = 33 and n = 70 orthonormal bases
    for nb_ in (33, 70):
        Phi = np.linalg.qr(np.random.default_rng(nb_).standard_normal((sm.vspace_dim, nb_)))[0].T
        rbx = BaseReducedBasis()
        rbx.set(basis=Phi, a=y[:nb_])
        fm = rbx.forward_modeling(sm, y[:20])
        assert relerr(fm, o.generate_fm_solutions(y[:20], Phi)) < 1e-9
        pj = rbx.projection(sm, U[:20])
        assert relerr(pj, o.project_solutions(U[:20], Phi)) < 1e-9


def test_pinned_result_pool_keeps_live_arrays_intact():
    """Large results of the numpy API live in pooled pinned blocks (engine.PinnedPool): an array the caller keeps -- or any
    view of it -- must never be overwritten by a later call; a block returns to the pool only after the array and all its
    views are gone, and is then reused."""
    import gc
    from src.lib.SolutionsManagers import SolutionsManagerFEM
    from romhighcontrast_b200.engine import PINNED_POOL, Engine
    sm = SolutionsManagerFEM((3, 3), 43)                    # D = 16 384: 1200 snapshots = 157 MB > the pipeline threshold
    K = 1200
    y1 = 10 ** np.random.default_rng(1).uniform(0, 6, (K, 3, 3))
    y2 = 10 ** np.random.default_rng(2).uniform(0, 6, (K, 3, 3))
    U1 = sm.generate_solutions(y1)
    assert U1.nbytes >= Engine.HOST_PIPELINE_MIN_BYTES and U1.flags.writeable and U1.flags.c_contiguous
    keep = U1.copy()
    view = U1[100:200, ::7]
    U2 = sm.generate_solutions(y2)                           # must take another block
    np.testing.assert_array_equal(U1, keep)
    assert not np.shares_memory(U1, U2)
    addr1 = U1.ctypes.data
    del U1
    gc.collect()
    U3 = sm.generate_solutions(y2)                           # `view` still holds the first block
    assert U3.ctypes.data != addr1
    np.testing.assert_array_equal(view, keep[100:200, ::7])
    np.testing.assert_array_equal(U3, U2)
    del view
    gc.collect()
    n_free = len(PINNED_POOL.free)
    assert n_free >= 1                                       # the first block is back
    U4 = sm.generate_solutions(y1)
    assert U4.ctypes.data == addr1 or len(PINNED_POOL.free) < n_free   # ... and gets reused
    np.testing.assert_array_equal(U4, keep)
