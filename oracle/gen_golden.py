#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (TEST INFRASTRUCTURE ONLY).

Run in the build container, where /root/reference exists:
    python oracle/gen_golden.py [--ref /root/reference] [--out tests/golden]

The reference imports `pathos` at module top (src/lib/SolutionsManagers.py:3), which is not
installed; a 2-line shim package is written to a temp dir and put on sys.path.  Nothing is
written under /root/reference and `src.config` (which mkdirs at import) is never imported.
The GPU box has no /root/reference: tests only read the committed .npz files.
"""
from __future__ import annotations

import argparse
import os
import sys
import tempfile
import warnings

import numpy as np


def import_reference(ref_root: str):
    shim = tempfile.mkdtemp(prefix="pathos_shim_")
    os.makedirs(os.path.join(shim, "pathos"))
    open(os.path.join(shim, "pathos", "__init__.py"), "w").close()
    with open(os.path.join(shim, "pathos", "multiprocessing.py"), "w") as f:
        f.write("from multiprocessing import Pool, cpu_count\n")
    # drop the repo root / cwd so that the repo's own `src` / `lib` drop-in packages do not shadow the reference
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") not in (repo, os.path.join(repo, "src"))]
    sys.path[:0] = [shim, ref_root, os.path.join(ref_root, "src")]
    for m in [m for m in sys.modules if m == "src" or m.startswith(("src.", "lib."))or m == "lib"]:
        del sys.modules[m]
    import src.lib.SolutionsManagers as SM
    import src.lib.ReducedBasis as RB
    import src.lib.Estimators as ES
    assert SM.__file__.startswith(ref_root), SM.__file__
    return SM, RB, ES


def mp_truth(A_csr, b, x0, dps=60, iters=12):
    """Iterative refinement with mpmath residuals -> the exactly-rounded discrete solution."""
    import mpmath as mp
    import scipy.sparse.linalg as spl
    mp.mp.dps = dps
    A = A_csr.tocsr()
    lu = spl.splu(A.tocsc())
    n = A.shape[0]
    x = [mp.mpf(float(v)) for v in x0]
    dmp = [mp.mpf(float(v)) for v in A.data]
    bmp = [mp.mpf(float(v)) for v in b]
    for _ in range(iters):
        r = np.empty(n)
        for i in range(n):
            s = bmp[i]
            for k in range(A.indptr[i], A.indptr[i + 1]):
                s -= dmp[k] * x[A.indices[k]]
            r[i] = float(s)
        dx = lu.solve(r)
        x = [xi + mp.mpf(float(d)) for xi, d in zip(x, dx)]
        if np.abs(dx).max() < 1e-30:
            break
    return np.array([float(v) for v in x])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                  "tests", "golden"))
    ap.add_argument("--skip-large", action="store_true", help="skip the N=32 (D=3969) greedy runs (~2 min)")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    warnings.filterwarnings("ignore")
    SM, RB, ES = import_reference(args.ref)
    save = lambda name, **kw: (np.savez_compressed(os.path.join(args.out, name), **kw), print("wrote", name))

    # ---- G1: assembly facts, (2,2) N=10 and a non-square (3,2) N=4 ------------------------
    sm = SM.SolutionsManagerFEM((2, 2), N=10)
    y = np.array([[1.0, 7.0], [100.0, 1e6]])
    Ay = np.einsum("pqij,pq->ij", sm.A_preassembled, y)
    save("g1_assembly_2x2_N10.npz", vspace_dim=sm.vspace_dim, B_total=sm.B_total,
         A1_diag=np.diag(sm.A_preassembled4h1_norm), y=y, Ay_diag=np.diag(Ay),
         Ay_row180=Ay[180], points_c=sm.points_c, points_r=sm.points_r)
    sm32 = SM.SolutionsManagerFEM((3, 2), N=4)
    save("g1_assembly_3x2_N4.npz", A_pre=sm32.A_preassembled, B_total=sm32.B_total,
         A1=sm32.A_preassembled4h1_norm, points_c=sm32.points_c, points_r=sm32.points_r)

    # ---- G2: snapshot solves, three reference methods ---------------------------------------------------
    ys = np.random.default_rng(0).uniform(1, 100, (5, 2, 2))
    out = {}
    for method in ("lsq", "lsqsparse", "ridge"):
        sm.method = method
        out[method] = sm.generate_solutions(ys)
    sm.method = "lsq"
    save("g2_solve_2x2_N10.npz", y=ys, **{f"U_{k}": v for k, v in out.items()},
         h10=sm.H10norm(out["lsq"]), l2=sm.l2norm(out["lsq"]))
    ys32 = 10 ** np.random.default_rng(7).uniform(0, 6, (6, 3, 2))
    U32 = sm32.generate_solutions(ys32)
    sm32.method = "lsqsparse"
    U32s = sm32.generate_solutions(ys32)
    sm32.method = "lsq"
    save("g2_solve_3x2_N4.npz", y=ys32, U_lsq=U32, U_lsqsparse=U32s, h10=sm32.H10norm(U32), l2=sm32.l2norm(U32))

    # ---- G3: reduced Galerkin / projection / point evaluation on (3,2) N=4 ---------------------------------
    rng = np.random.default_rng(3)
    Phi = np.linalg.qr(rng.standard_normal((sm32.vspace_dim, 5)))[0].T
    Phi_snap = RB.orthonormalize_base(U32[:4])
    pts = rng.uniform(low=[sm32.x_domain[0], sm32.y_domain[0]], high=[sm32.x_domain[1], sm32.y_domain[1]],
                      size=(17, 2))
    nodes = np.array([(x, yv) for yv in sm32.points_r[1:-1] for x in sm32.points_c[1:-1]])
    save("g3_reduced_3x2_N4.npz", y=ys32, U=U32, Phi=Phi, Phi_snap=Phi_snap,
         fm=sm32.generate_fm_solutions(ys32, Phi), fm_snap=sm32.generate_fm_solutions(ys32, Phi_snap),
         proj=sm32.project_solutions(U32, Phi), proj_snap=sm32.project_solutions(U32, Phi_snap),
         fm_empty=sm32.generate_fm_solutions(ys32, np.empty((0, 0))),
         pts=pts, ev=sm32.evaluate_solutions(pts, U32), nodes=nodes,
         ev_nodes=sm32.evaluate_solutions(nodes, U32[:2]),
         riesz_l2=sm32.generate_riesz(pts[:3], norm="l2"))

    # ---- G4: greedy index sequences ----------------------------------------------------------
    ysg = 10 ** np.random.default_rng(42).uniform(0, 6, (100, 2, 2))
    for N in ([10] if args.skip_large else [10, 32]):
        smg = SM.SolutionsManagerFEM((2, 2), N=N, method="lsq")
        U = smg.generate_solutions(ysg)
        h1 = smg.H10norm(U)
        res = {}
        for tag, crit in (("gal", RB.GREEDY_FOR_GALERKIN), ("h10", RB.GREEDY_FOR_H10)):
            rb = RB.ReducedBasisGreedy(greedy_for=crit).build(n=10, sm=smg, solutions2train=U, a2train=ysg,
                                                             solutions2train_h1norm=h1)
            idx = [int(np.argmin(np.abs(U - b).sum(axis=1))) for b in rb.basis]
            res[f"idx_{tag}"] = np.array(idx)
            res[f"a_{tag}"] = np.array(rb.a)
            rb.orthonormalize()
            fm = rb.forward_modeling(smg, ysg)
            pj = rb.projection(smg, U)
            res[f"fm_err_{tag}"] = smg.H10norm(fm - U) / h1
            res[f"pj_err_{tag}"] = smg.H10norm(pj - U) / h1
            if N == 10:
                res[f"basis_orth_{tag}"] = rb.basis
        if N == 10:
            res["U"] = U
            # default solutions2train_h1norm=1 variant (largest-norm snapshot wins first)
            rb = RB.ReducedBasisGreedy().build(n=4, sm=smg, solutions2train=U, a2train=ysg)
            res["idx_gal_unnormalised"] = np.array([int(np.argmin(np.abs(U - b).sum(axis=1))) for b in rb.basis])
        save(f"g4_greedy_2x2_N{N}.npz", y=ysg, h1=h1, **res)
        if N == 10:
            U10, sm10, h110 = U, smg, h1

    # ---- G5/G6: PCA and random builders, state estimation, estimators ------------------------------------------
    from sklearn.decomposition import PCA
    full = PCA(n_components=10, svd_solver="full").fit(U10)
    ref_pca = RB.ReducedBasisPCA().build(n=10, sm=sm10, solutions2train=U10, a2train=ysg)
    ref_rand = RB.ReducedBasisRandom().build(n=10, sm=sm10, solutions2train=U10, a2train=ysg, seed=42)
    rbg = RB.ReducedBasisGreedy().build(n=6, sm=sm10, solutions2train=U10, a2train=ysg,
                                        solutions2train_h1norm=h110)
    mp_pts = np.random.default_rng(1).uniform(low=[-1, -1], high=[1, 1], size=(30, 2))
    meas = sm10.evaluate_solutions(mp_pts, U10[:20])
    c, est = rbg.state_estimation(sm10, mp_pts, meas, return_coefs=True)
    save("g5_builders_2x2_N10.npz", y=ysg, U=U10, pca_full_components=full.components_,
         pca_full_singular_values=full.singular_values_, pca_mean=full.mean_,
         pca_ref_components=ref_pca.basis, random_basis=ref_rand.basis, random_a=ref_rand.a,
         greedy6_basis=rbg.basis, greedy6_a=np.array(rbg.a), points=mp_pts, measurements=meas,
         se_c=c, se_est=est, inv=rbg.parameter_estimation_inverse(c), lin=rbg.parameter_estimation_linear(c),
         sliced_dim=rbg[:3].dim)

    # ---- G7: INFINIT_A training set (inf split) on (2,2) N=6 -----------------------------------------------------
    sm6 = SM.SolutionsManagerFEM((2, 2), N=6, method="lsq")
    hc = np.array([[1e10, 1e10], [1e10, 1.0], [1.0, 1e10], [1.0, 1.0], [3.0, 50.0], [700.0, 2.0], [9.0, 9e3],
                   [1e4, 1e2], [5.0, 1.5]])
    a6 = np.ones((len(hc), 2, 2))
    a6[:, 0, 0] = hc[:, 0]
    a6[:, 1, 1] = hc[:, 1]
    U6 = sm6.generate_solutions(a6)
    r6 = RB.ReducedBasisRandom(True).build(n=5, sm=sm6, solutions2train=U6, a2train=a6)
    r6n = RB.ReducedBasisRandom(False).build(n=3, sm=sm6, solutions2train=U6, a2train=a6)
    g6 = RB.ReducedBasisGreedy().build(n=5, sm=sm6, solutions2train=U6, a2train=a6,
                                       solutions2train_h1norm=sm6.H10norm(U6))
    save("g7_inf_2x2_N6.npz", a=a6, U=U6, rand_inf_basis=r6.basis, rand_inf_a=r6.a, rand_noinf_basis=r6n.basis,
         rand_noinf_a=r6n.a, greedy_idx=np.array([int(np.argmin(np.abs(U6 - b).sum(axis=1))) for b in g6.basis]))

    # ---- G8: floating 1e10 inclusion: the reference is only ~1e-5 accurate; store an mpmath truth --------------------
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle.fem import FEMOracle
    sm48 = SM.SolutionsManagerFEM((4, 4), N=8, method="lsq")
    a48 = np.ones((1, 4, 4))
    a48[0, 1:3, 1:3] = 1e10
    u_lsq = sm48.generate_solutions(a48)[0]
    sm48.method = "lsqsparse"
    u_sp = sm48.generate_solutions(a48)[0]
    orc = FEMOracle((4, 4), 8)
    truth = mp_truth(orc.matrix(a48[0]), orc.B_total, u_sp)
    rel = lambda u: np.linalg.norm(u - truth) / np.linalg.norm(truth)
    print(f"floating inclusion: reference lsq err {rel(u_lsq):.2e}, lsqsparse err {rel(u_sp):.2e}")
    save("g8_floating_4x4_N8.npz", a=a48, U_lsq=u_lsq, U_lsqsparse=u_sp, U_truth=truth)


if __name__ == "__main__":
    main()
