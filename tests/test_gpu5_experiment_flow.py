"""The call sequence of the reference's experiment driver on the GPU-backed classes.

Restates /root/reference/src/experiments/HighContrast.py:99-115 (training-set sampler with the INFINIT_A corner
points) and :138-214 (snapshots -> norms -> builders -> per-n statistics -> joblib checkpoint) as a test and checks
the resulting error curves against the CPU oracle running the same flow on the same snapshots.
"""
import io

import joblib
import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu

INFINIT_A = 1e10


def get_a2test_and_train(blocks_geometry, high_contrast_blocks, refinement, max_samples, seed):
    d = len(high_contrast_blocks)
    num = min(refinement * int(np.log2(INFINIT_A)), int(np.ceil(max_samples ** (1 / d))))
    a_hc = np.transpose(list(map(np.ravel, np.meshgrid(*[1 / np.linspace(1 / INFINIT_A, 1, num=num, endpoint=False)] * d))))
    np.random.seed(seed)
    a_inf = np.transpose(list(map(np.ravel, np.meshgrid(*[[INFINIT_A, 1]] * d))))
    if len(a_hc) > max_samples - len(a_inf):
        a_hc = a_hc[np.random.choice(len(a_hc), size=max((0, max_samples - len(a_inf))), replace=False)]
    a_hc = np.vstack((a_inf, a_hc))
    a = np.ones((len(a_hc),) + tuple(blocks_geometry))
    for a_vec, same in zip(a_hc.T, high_contrast_blocks):
        for ix in same:
            a[:, ix[0], ix[1]] = a_vec
    return a, a_hc


def test_experiment_flow_matches_oracle():
    from lib.ReducedBasis import ReducedBasisGreedy, ReducedBasisRandom, GREEDY_FOR_GALERKIN, GREEDY_FOR_H10
    from lib.SolutionsManagers import SolutionsManagerFEM
    from oracle import FEMOracle, state_estimation, estimator_inv, estimator_linear, sort_orthogonalize_base, \
        high_contrast_coefficient
    geo, N, vn = (4, 4), 6, 8
    hcb = [[(0, 1)], [(1, 3)], [(2, 1), (2, 2), (2, 3)]]           # __main__ of HighContrast.py:512
    a, a_hc = get_a2test_and_train(geo, hcb, refinement=10, max_samples=120, seed=42)
    assert a[0].max() == INFINIT_A and len(a) == 120
    sm = SolutionsManagerFEM(geo, N=N, num_cores=1, method="lsqsparse")
    o = FEMOracle(geo, N)
    data = {"solutions": sm.generate_solutions(a2try=a)}
    data["solutions_H1norm"] = sm.H10norm(solutions=data["solutions"])
    U, h1 = data["solutions"], data["solutions_H1norm"]
    # snapshots: finite-contrast rows to 1e-9; rows with a 1e10 block only to the reference's own accuracy
    Uo = o.generate_solutions(a)
    fin = a.max(axis=(1, 2)) < 1e7
    assert relerr(U[fin], Uo[fin]) < 1e-9 and relerr(U, Uo) < 1e-3
    np.random.seed(0)
    pts = np.random.uniform(size=(30, 2))
    meas = sm.evaluate_solutions(pts, U)
    builders = [ReducedBasisRandom(), ReducedBasisRandom(False), ReducedBasisGreedy(greedy_for=GREEDY_FOR_H10),
                ReducedBasisGreedy(greedy_for=GREEDY_FOR_GALERKIN)]
    for b in builders:
        data[b.name] = {"errors": {}, "basis": b.build(n=vn, sm=sm, solutions2train=U, a2train=a, optim_method="lsq",
                                                        solutions2train_h1norm=h1)}
    for n in range(1, vn + 1):
        for b in builders:
            rb = data[b.name]["basis"][:n]
            c, se = rb.state_estimation(sm=sm, measurement_points=pts, measurements=meas, return_coefs=True)
            inv = rb.parameter_estimation_inverse(c=c)
            lin = rb.parameter_estimation_linear(c=c)
            rb.orthonormalize()
            fm = rb.forward_modeling(sm=sm, a=a)
            pj = rb.projection(sm=sm, true_solutions=U)
            e_fm = sm.H10norm(solutions=fm - U) / h1
            e_pj = sm.H10norm(solutions=pj - U) / h1
            e_se = sm.H10norm(solutions=se - U) / h1
            data[b.name]["errors"][n] = (e_fm, e_pj, e_se, np.abs(1 - np.array(inv) / a), np.abs(1 - np.array(lin) / a))
            # Cea: the H10 projection is the best approximation in the space (up to the a != 1 norm equivalence the
            # Galerkin solution is quasi-optimal); both shrink with n
            assert np.all(e_pj <= e_fm * (1 + 1e-6) + 1e-9)
            # oracle on the same basis
            basis_raw = np.asarray(data[b.name]["basis"].basis[:n])
            a_raw = np.asarray(data[b.name]["basis"].a[:n])
            _, Phi = sort_orthogonalize_base(high_contrast_coefficient(a_raw), basis_raw.reshape(n, -1))
            assert relerr(np.abs(rb.basis), np.abs(Phi)) < 1e-9
            fo = o.generate_fm_solutions(a, Phi)
            po = o.project_solutions(U, Phi)
            scale = np.linalg.norm(U)
            assert np.linalg.norm(fm[fin] - fo[fin]) / scale < 1e-8, (b.name, n)
            assert np.linalg.norm(pj - po) / scale < 1e-8, (b.name, n)
            # raw snapshot bases are nearly dependent: least-squares coefficients are only defined up to
            # cond(E) * eps (beyond ~1e10 the reference's own lstsq result is rounding noise), so compare there
            # only when the collocation matrix is reasonably conditioned
            E = o.evaluate_solutions(pts, basis_raw)
            if np.linalg.cond(E) < 1e8:
                co, so = state_estimation(o, basis_raw, pts, meas)
                assert np.linalg.norm(se - so) / scale < 1e-6 * np.linalg.cond(E)
    g = data["Greedy galerkin"]["errors"]
    assert g[vn][0].max() < g[1][0].max()
    # greedy picks the all-1e10 corner first (index 0), as in the reference (SURVEY 8a row a8 quirk i)
    assert data["Greedy galerkin"]["basis"].selected_indices[0] == 0
    assert data[r"Greedy $H^1_0$"]["basis"].selected_indices[0] == 0
    # checkpoint round trip (HighContrast.py:93-96, 214)
    buf = io.BytesIO()
    joblib.dump(data, buf)
    buf.seek(0)
    back = joblib.load(buf)
    np.testing.assert_array_equal(back["Greedy galerkin"]["basis"].basis, data["Greedy galerkin"]["basis"].basis)
    assert back["Greedy galerkin"]["basis"][:3].dim == 3


def test_experiment_driver_against_the_reference_drivers_record(tmp_path, capsys):
    """The experiment driver (tests/experiment_driver.py: the reference's `experiment()` restated and pinned bit for bit
    on the unmodified reference classes by tests/test_experiment_driver_cpu.py) runs on the GPU classes, and its `data`
    dictionary is compared with what the reference's OWN driver produced on the reference's classes
    (tests/golden/g9_experiment_4x4_N6.npz): same keys and timing entries, same training set and measurement points,
    same selected snapshots, same error curves, same AttributeError on the cached path (HighContrast.py:172)."""
    import types
    import experiment_driver as drv
    import lib.ReducedBasis as RB
    import lib.SolutionsManagers as SM
    from conftest import golden
    g = golden("g9_experiment_4x4_N6.npz")
    ns = types.SimpleNamespace(SolutionsManagerFEM=SM.SolutionsManagerFEM, ReducedBasisGreedy=RB.ReducedBasisGreedy,
                               ReducedBasisRandom=RB.ReducedBasisRandom, INFINIT_A=RB.INFINIT_A,
                               GREEDY_FOR_H10=RB.GREEDY_FOR_H10, GREEDY_FOR_GALERKIN=RB.GREEDY_FOR_GALERKIN)
    kw = dict(N=6, refinement=10, vn_max_dim=8, num_measurements=30, blocks_geometry=(4, 4),
              high_contrast_blocks=[[(0, 1)], [(1, 3)], [(2, 1), (2, 2), (2, 3)]], max_samples=120, seed=42,
              method="lsqsparse")                                  # oracle/gen_golden_experiment.py CONFIG
    builders = drv.default_builders(ns)
    sm, data, a, a_hc, points = drv.experiment(ns, tmp_path / "exp", builders, recalculate=True, recalculate_basis=True, **kw)
    # host-side bookkeeping: identical
    assert [b.name for b in builders] == list(g["names"])
    assert sorted(data.keys()) == list(g["data_keys"])
    np.testing.assert_array_equal(a, g["a"])
    np.testing.assert_array_equal(a_hc, g["a_high_contrast"])
    np.testing.assert_array_equal(points, g["points"])
    assert isinstance(data["time2calculate_solutions"], float) and isinstance(data["time2calculate_h1norm"], float)
    # snapshots: rows without a 1e10 block to 1e-9 against the reference's SuperLU path, the others to the reference's own
    # accuracy there (cond ~ 1e12, the reference is ~1e-5 off the exact discrete solution)
    U, Ug = data["solutions"], g["solutions"]
    fin = a.max(axis=(1, 2)) < 1e7
    row_err = np.linalg.norm(U - Ug, axis=1) / np.linalg.norm(Ug, axis=1)
    assert row_err[fin].max() < 1e-9 and row_err.max() < 1e-3, (row_err[fin].max(), row_err.max())
    np.testing.assert_allclose(data["solutions_H1norm"][fin], g["solutions_H1norm"][fin], rtol=1e-9)
    np.testing.assert_allclose(data["solutions_H1norm"], g["solutions_H1norm"], rtol=1e-3)
    report, bad = {}, []

    def compare(tag, mine, ref, rows, rtol, atol):
        """|mine - ref| <= rtol |ref| + atol on `rows`; the worst ratio to that bound goes into the report"""
        if not rows.any():
            return
        ratio = float((np.abs(mine - ref)[rows] / (rtol * np.abs(ref)[rows] + atol)).max())
        report[tag] = max(report.get(tag, 0.0), ratio)
        if not ratio <= 1.0:
            bad.append((tag, ratio))

    for i, name in enumerate(g["names"]):
        d = data[str(name)]
        assert sorted(d.keys()) == list(g[f"b{i}_keys"]), name
        rb = d["basis"]
        assert isinstance(d["time2build"], float) and rb.dim == kw["vn_max_dim"]
        idx = np.array([int(np.argmin(np.abs(U - b).sum(axis=1))) for b in np.asarray(rb.basis)])
        np.testing.assert_array_equal(idx, g[f"b{i}_idx"], err_msg=str(name))          # same snapshots, same order
        np.testing.assert_array_equal(np.asarray(rb.basis), U[idx])                    # raw rows, bit for bit
        np.testing.assert_array_equal(np.asarray(rb.a), g[f"b{i}_a"])
        assert sorted(d["errors"].keys()) == list(g[f"b{i}_ns"]) == sorted(d["times"].keys())
        for n in g[f"b{i}_ns"]:
            e, t = d["errors"][n], d["times"][n]
            assert type(e).__name__ == type(t).__name__ == "TypeOfProblems" and e._fields == tuple(g["fields"])
            assert all(isinstance(v, float) for v in t)
            clean = bool(fin[idx][:n].all())        # no 1e10 snapshot in the basis: nothing of the reference's 1e-5 there
            kind = "finite basis" if clean else "1e10 rows in basis"
            for f in ("forward_modeling", "projection"):
                mine, ref = np.asarray(getattr(e, f)), g[f"b{i}_n{n}_{f}"]
                assert mine.shape == ref.shape == (len(a),)
                # relative H10 errors of the approximations (values between 1e-16, a snapshot of the span, and 1)
                if clean:
                    compare(f"{f} | {kind} | finite rows", mine, ref, fin, 1e-7, 1e-9)
                else:
                    compare(f"{f} | {kind} | finite rows", mine, ref, fin, 1e-5, 1e-8)
                compare(f"{f} | {kind} | 1e10 rows", mine, ref, ~fin, 2e-2, 1e-5)
            # state estimation: least squares on raw snapshots; only defined up to cond(E) eps
            E = sm.evaluate_solutions(points, np.asarray(rb.basis)[:n])
            cond = np.linalg.cond(E)
            if cond < 1e6:
                mine, ref = np.asarray(e.state_estimation), g[f"b{i}_n{n}_state_estimation"]
                compare(f"state_estimation | {kind} | finite rows, cond(E) < 1e6", mine, ref, fin,
                        1e-7 if clean else 2e-3, 1e-10 * cond if clean else 1e-6)
                for f in ("parameter_estimation_inverse", "parameter_estimation_linear"):
                    mine, ref = np.asarray(getattr(e, f)), g[f"b{i}_n{n}_{f}"]
                    assert mine.shape == ref.shape == a.shape
                    compare(f"{f} | {kind} | finite rows, cond(E) < 1e6", mine, ref, np.broadcast_to(fin[:, None, None], a.shape),
                            1e-7 if clean else 2e-3, 1e-10 * cond if clean else 1e-6)
    with capsys.disabled():
        print("\n[experiment driver vs reference record] worst |mine - ref| / (rtol |ref| + atol) per comparison:")
        for k, v in sorted(report.items()):
            print("   %-90s %.2e" % (k, v))
    assert not bad, bad
    # cached path: the second call finds everything in data.compressed and trips over `.marker`, like the reference
    with pytest.raises(AttributeError) as ei:
        drv.experiment(ns, tmp_path / "exp", builders, **kw)
    assert f"{type(ei.value).__name__}: {ei.value}" == str(g["cached_call_exception"])
