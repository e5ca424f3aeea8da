"""GPU parity at BASELINE.json's full mesh size through size-independent properties (the oracle only spot-checks).

configs[2]: (4,4) subdomains, N = 64 (256 x 256 cells, D = 65 025), contrast 10^U(0,6).  Tolerances: 1e-9 relative
(north_star) on anything compared with the oracle; the algebraic identities are checked to the accuracy the PCG
stopping rule (rtol 1e-12 on the preconditioned residual) implies.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GEO, N = (4, 4), 64


def sample(K, seed):
    return 10 ** np.random.default_rng(seed).uniform(0, 6, (K,) + GEO)


@pytest.fixture(scope="module")
def eng():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from romhighcontrast_b200.engine import Engine
    return Engine(GEO, N)


def test_snapshots_full_size_identities(eng):
    import torch
    from oracle import FEMOracle
    K = 4096
    y = sample(K, 42)
    yd = eng.params(y)
    x, iters, relres = eng.solve(yd)
    assert eng.last_solve_stats["status"] == 0
    assert int(iters.max()) <= 30 and int(iters.min()) >= 5
    assert float(relres.max()) <= 1e-12 * 1.0000001
    # true residual b - A(y) u, b = 1 / N^2 on the interior vertices (SolutionsManagers.py:177-185).  With contrast
    # 1e6 the Euclidean norm of the residual is dominated by the stiff blocks (entries of A up to 4e6), so the
    # scale-free measure is the Jacobi-scaled one: |r_i| / diag_i relative to max |u|
    b = eng.pad(np.full((1, eng.D), 1.0 / N ** 2))
    res = b - eng.apply(yd, x)
    sub = slice(0, 64)
    kc = np.kron(y[sub], np.ones((1, N, N)))                               # cell coefficients (64, R, C)
    kp = np.pad(kc, ((0, 0), (1, 1), (1, 1)))
    diag = (kp[:, :-1, :-1] + kp[:, :-1, 1:] + kp[:, 1:, :-1] + kp[:, 1:, 1:])[:, 1:-1, 1:-1].reshape(64, -1)
    r_sub = eng.unpad(res[sub]).cpu().numpy()
    u_sub = eng.unpad(x[sub]).cpu().numpy()
    scaled = np.abs(r_sub / diag).max(axis=1) / np.abs(u_sub).max(axis=1)
    assert scaled.max() < 1e-11, scaled.max()
    rel = torch.linalg.vector_norm(res, dim=1) / torch.linalg.vector_norm(b)
    assert float(rel.max()) < 1e-6, float(rel.max())
    # energy identity u^T A u = b^T u
    en2 = eng.energy_norm(yd, x) ** 2
    bu = (x * b).sum(dim=1)
    assert float(((en2 - bu).abs() / bu).max()) < 1e-10
    # scaling: A(s y) = s A(y)  =>  u(s y) = u(y) / s, exactly representable for s = 4
    sel = slice(0, 512)
    x4, _, _ = eng.solve(eng.params(4.0 * y[sel]))
    d = torch.linalg.vector_norm(4.0 * x4 - x[sel], dim=1) / torch.linalg.vector_norm(x[sel], dim=1)
    assert float(d.max()) < 1e-10
    # oracle spot check (sparse LU restatement of the reference)
    pick = [0, K // 2, K - 1]
    Uo = FEMOracle(GEO, N).generate_solutions(y[pick])
    U = eng.unpad(x[pick]).cpu().numpy()
    err = np.linalg.norm(U - Uo, axis=1) / np.linalg.norm(Uo, axis=1)
    assert err.max() < 1e-9, err


def test_host_entry_full_size_matches_resident(eng):
    """chunked, pipelined host-buffer entry point == device-resident solve, bit for bit (same kernels, same order)"""
    K = 4100                                   # > 4096: exercises the multi-chunk pipeline and a ragged last chunk
    y = sample(K, 7)
    U_host, iters, relres = eng.generate_solutions_host(y, return_stats=True)
    x, it_d, rel_d = eng.solve(eng.params(y))
    np.testing.assert_array_equal(U_host, eng.unpad(x).cpu().numpy())
    np.testing.assert_array_equal(iters, it_d.cpu().numpy())
    np.testing.assert_array_equal(relres, rel_d.cpu().numpy())


def test_gram_and_online_stage_full_size(eng):
    import torch
    K, n = 2048, 20
    y = sample(K, 3)
    x, _, _ = eng.solve(eng.params(y))
    # centred Gram on the DMMA path: symmetric, trace = sum of squared row norms, sub-block equals a torch fp64 product
    mean = eng.column_mean(x)
    Xc = x - mean[None, :]
    G = eng.gemm_nt(Xc, Xc, symmetric=True)
    assert torch.equal(G, G.T)
    tr = float(torch.trace(G)); ref = float((Xc * Xc).sum())
    assert abs(tr - ref) <= 1e-12 * ref
    blk = Xc[:192] @ Xc[:64].T
    assert float((G[:192, :64] - blk).abs().max()) <= 1e-12 * float(blk.abs().max())
    # POD basis from the top eigenpairs, then 200k online reduced Galerkin solves: the reduced residual vanishes
    from romhighcontrast_b200.pod import top_eigenpairs
    lam, V = top_eigenpairs(eng, G, n)
    Phi = (eng.gemm_tn(V, Xc) / torch.sqrt(lam)[:, None]).contiguous()
    gram_phi = Phi @ Phi.T
    assert float((gram_phi - torch.eye(n, dtype=torch.float64, device=Phi.device)).abs().max()) < 1e-9
    Ahat, bhat = eng.project_operators(Phi)
    Ko = 200000
    yo = sample(Ko, 43)
    C = eng.reduced_solve(eng.params(yo), Ahat, bhat)
    idx = np.random.default_rng(0).choice(Ko, 300, replace=False)
    Ah, bh, Ch = Ahat.cpu().numpy(), bhat.cpu().numpy(), C.cpu().numpy()
    for k in idx:
        A = np.einsum("q,qij->ij", yo[k].ravel(), Ah)
        r = A @ Ch[k] - bh
        assert np.linalg.norm(r) <= 1e-10 * np.linalg.norm(bh)
    # scaling property of the reduced problem: c(2 y) = c(y) / 2
    # (sqrt(2 d) and sqrt(2) sqrt(d) round differently, so the two Cholesky solves agree to eps * cond(A_k) per system,
    #  not to a fixed 1e-12: cond reaches 1e6+ at contrast 10^U(0,6))
    C2 = eng.reduced_solve(eng.params(2.0 * yo[:1000]), Ahat, bhat).cpu().numpy()
    A1k = np.einsum("kq,qij->kij", yo[:1000].reshape(1000, -1), Ah)
    cond = np.linalg.cond(A1k)
    eps = np.finfo(float).eps
    dev = np.linalg.norm(2.0 * C2 - Ch[:1000], axis=1)
    assert np.all(dev <= 50 * eps * cond * np.linalg.norm(Ch[:1000], axis=1)), float((dev / (eps * cond * np.linalg.norm(Ch[:1000], axis=1))).max())
    # and both agree with LAPACK on the same systems to the north-star tolerance
    Cl = np.linalg.solve(A1k, np.broadcast_to(bh, (1000, n))[..., None])[..., 0]
    assert (np.linalg.norm(Ch[:1000] - Cl, axis=1) / np.linalg.norm(Cl, axis=1)).max() < 1e-9


def test_config1_greedy_and_pca_parity_at_full_size():
    """BASELINE configs[1]: (3,3) subdomains, 128 x 128 interior nodes (N = 43: prime, single-level solver), 1000 samples,
    contrast up to 1e6, n = 20: both greedy criteria pick the oracle's index sequence, PCA singular values agree."""
    from lib.ReducedBasis import ReducedBasisGreedy, ReducedBasisPCA, GREEDY_FOR_GALERKIN, GREEDY_FOR_H10
    from lib.SolutionsManagers import SolutionsManagerFEM
    from oracle import FEMOracle
    from oracle.rb import greedy_build, pca_components
    geo, Nb, K, n = (3, 3), 43, 1000, 20
    y = 10 ** np.random.default_rng(42).uniform(0, 6, (K,) + geo)
    sm = SolutionsManagerFEM(geo, Nb, method="lsqsparse")
    U = sm.generate_solutions(y)
    o = FEMOracle(geo, Nb)
    Uo = o.generate_solutions(y[:8])
    assert (np.linalg.norm(U[:8] - Uo, axis=1) / np.linalg.norm(Uo, axis=1)).max() < 1e-9
    h1 = sm.H10norm(U)
    np.testing.assert_allclose(h1, o.H10norm(U), rtol=1e-12)
    for crit in (GREEDY_FOR_GALERKIN, GREEDY_FOR_H10):
        rb = ReducedBasisGreedy(greedy_for=crit).build(n=n, sm=sm, solutions2train=U, a2train=y, solutions2train_h1norm=h1)
        _, _, picked, trace = greedy_build(o, n, U, y, o.H10norm(U), greedy_for=crit, return_trace=True)
        for step, (a_, b_) in enumerate(zip(picked, rb.selected_indices)):
            top = np.sort(trace[step])[-2:]
            gap = (top[1] - top[0]) / top[1]
            assert a_ == b_ or gap < 1e-9, (crit, step, a_, b_, gap)     # bit-identical wherever the gap exceeds 1e-9
            if a_ != b_:
                break
    rbp = ReducedBasisPCA().build(n=n, sm=sm, solutions2train=U, a2train=y)
    _, so, _ = pca_components(U, n)
    assert np.max(np.abs(np.asarray(rbp.singular_values_) - so) / so) < 1e-9


def test_config4_geometry_parity():
    """BASELINE configs[4]: (8,8) subdomains, N = 64 (512 x 512 cells, D = 261 121), contrast 10^U(0,6): three snapshots
    against the sparse direct oracle (<= 1e-9 relative, north_star), the energy identity on all of them, the reduced
    Galerkin stage with nb = 64 blocks against LAPACK."""
    import torch
    from oracle import FEMOracle
    from romhighcontrast_b200.engine import Engine
    geo, Nb, K, n = (8, 8), 64, 96, 12
    eng = Engine(geo, Nb)
    assert eng.D == 261121
    y = 10 ** np.random.default_rng(11).uniform(0, 6, (K,) + geo)
    yd = eng.params(y)
    x, iters, relres = eng.solve(yd)
    assert eng.last_solve_stats["status"] == 0 and float(relres.max()) <= 1e-12 * 1.0000001
    assert int(iters.max()) <= 40
    b = eng.pad(np.full((1, eng.D), 1.0 / Nb ** 2))
    en2 = eng.energy_norm(yd, x) ** 2
    bu = (x * b).sum(dim=1)
    assert float(((en2 - bu).abs() / bu).max()) < 1e-10
    pick = [0, K // 2, K - 1]
    Uo = FEMOracle(geo, Nb).generate_solutions(y[pick])
    U = eng.unpad(x[pick].contiguous()).cpu().numpy()
    err = np.linalg.norm(U - Uo, axis=1) / np.linalg.norm(Uo, axis=1)
    assert err.max() < 1e-9, err
    # online stage with 64 blocks: orthonormal basis from the snapshots, reduced operators, reduced solves vs LAPACK
    Phi = torch.linalg.qr(x[:n].T)[0].T.contiguous()
    Ahat, bhat = eng.project_operators(Phi)
    C = eng.reduced_solve(yd, Ahat, bhat).cpu().numpy()
    A = np.einsum("kq,qij->kij", y.reshape(K, -1), Ahat.cpu().numpy())
    Cl = np.linalg.solve(A, np.broadcast_to(bhat.cpu().numpy(), (K, n))[..., None])[..., 0]
    assert (np.linalg.norm(C - Cl, axis=1) / np.linalg.norm(Cl, axis=1)).max() < 1e-9


def test_config3_observation_batch_100k():
    """BASELINE configs[3]: state and parameter estimation from m = 50 point measurements over a 100 000-observation batch
    on the (4,4), N = 64 model (romhighcontrast_b200.inverse.observation_batch_estimation); a sub-sample against the CPU
    oracle's restatement of ReducedBasis.py:65-86 / Estimators.py:24-37 and of the notebook's PBDW correction."""
    from lib.ReducedBasis import ReducedBasisGreedy
    from lib.SolutionsManagers import SolutionsManagerFEM
    from oracle import FEMOracle, pbdw_correction
    from oracle.rb import state_estimation, estimator_inv, estimator_linear
    from romhighcontrast_b200.inverse import observation_batch_estimation
    geo, Nb, n, m, Kobs, Ktrain = (4, 4), 64, 20, 50, 100000, 1000
    sm = SolutionsManagerFEM(geo, Nb)
    ytr = 10 ** np.random.default_rng(42).uniform(0, 6, (Ktrain,) + geo)
    Utr = sm.generate_solutions(ytr)
    rb = ReducedBasisGreedy().build(n=n, sm=sm, solutions2train=Utr, a2train=ytr, solutions2train_h1norm=sm.H10norm(Utr))
    pts = np.random.default_rng(1).uniform(low=[sm.x_domain[0], sm.y_domain[0]], high=[sm.x_domain[1], sm.y_domain[1]], size=(m, 2))
    yobs = 10 ** np.random.default_rng(44).uniform(0, 6, (Kobs,) + geo)
    res = observation_batch_estimation(sm, rb, pts, yobs)
    Z, c = res["measurements"], res["coefficients"]
    assert Z.shape == (Kobs, m) and c.shape == (n, Kobs) and res["a_inverse"].shape == (Kobs,) + geo
    o = FEMOracle(geo, Nb)
    sel = np.arange(0, Kobs, Kobs // 64)[:64]
    Uo = o.generate_solutions(yobs[sel[:4]])
    Zo = o.evaluate_solutions(pts, Uo)
    assert np.abs(Z[sel[:4]] - Zo).max() <= 1e-9 * np.abs(Zo).max()
    co, esto = state_estimation(o, np.asarray(rb.basis), pts, Z[sel])
    assert np.linalg.norm(c[:, sel] - co) <= 1e-9 * np.linalg.norm(co)
    np.testing.assert_allclose(res["a_inverse"][sel], estimator_inv(c[:, sel], np.asarray(rb.a)), rtol=1e-10)
    np.testing.assert_allclose(res["a_linear"][sel], estimator_linear(c[:, sel], np.asarray(rb.a)), rtol=1e-10, atol=1e-300)
    # state errors of the first four observations against the oracle's own estimates of the same quantities
    h = o.H10norm(Uo)
    e_ls = o.H10norm(esto[:4] - Uo) / h
    np.testing.assert_allclose(res["err_ls"][sel[:4]], e_ls, rtol=1e-6)
    e_pb = o.H10norm(pbdw_correction(o, pts, Z[sel[:4]], esto[:4]) - Uo) / h
    np.testing.assert_allclose(res["err_pbdw"][sel[:4]], e_pb, rtol=1e-6)
    assert np.all(np.isfinite(res["err_ls"])) and np.all(np.isfinite(res["err_pbdw"]))
