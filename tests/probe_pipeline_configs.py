"""Probe (not a test): wall times of the src.lib-level pipeline on the GPU for BASELINE configs[0] and configs[1], with the
CPU oracle running the same builders on the same snapshots (greedy index parity + timing)."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import torch
from lib.SolutionsManagers import SolutionsManagerFEM
from lib.ReducedBasis import ReducedBasisGreedy, ReducedBasisPCA, GREEDY_FOR_GALERKIN, GREEDY_FOR_H10
from oracle import FEMOracle
from oracle.rb import greedy_build, pca_components


def tm(f, *a, **k):
    torch.cuda.synchronize(); t = time.perf_counter(); r = f(*a, **k); torch.cuda.synchronize()
    return r, time.perf_counter() - t


for name, geo, N, K, n, do_oracle in (("configs[0]", (2, 2), 32, 100, 10, True), ("configs[1]", (3, 3), 43, 1000, 20, True)):
    y = 10 ** np.random.default_rng(42).uniform(0, 6, (K,) + geo)
    sm = SolutionsManagerFEM(geo, N, method="lsqsparse")
    sm.generate_solutions(y[:4])                                     # context + workspace warm-up
    U, t_first = tm(sm.generate_solutions, y)                       # pays the workspace / pinned staging allocations at this K
    U, t_snap = tm(sm.generate_solutions, y)
    h1, t_h1 = tm(sm.H10norm, U)
    out = {"snapshots_first_call_s": t_first, "snapshots_s": t_snap, "solves_per_s": K / t_snap, "H10norm_s": t_h1,
           "pcg_iterations_mean": float(np.mean(sm.last_solver_report["iterations"]))}
    picks = {}
    for crit in (GREEDY_FOR_GALERKIN, GREEDY_FOR_H10):
        b = ReducedBasisGreedy(greedy_for=crit)
        b.build(n=n, sm=sm, solutions2train=U, a2train=y, solutions2train_h1norm=h1)   # warm-up
        b = ReducedBasisGreedy(greedy_for=crit)
        rb, t_g = tm(b.build, n=n, sm=sm, solutions2train=U, a2train=y, solutions2train_h1norm=h1)
        out[f"greedy_{crit}_s"] = t_g
        picks[crit] = list(rb.selected_indices)
    _, t_p = tm(ReducedBasisPCA().build, n=n, sm=sm, solutions2train=U, a2train=y)
    rbp, t_p = tm(ReducedBasisPCA().build, n=n, sm=sm, solutions2train=U, a2train=y)
    out["pca_s"] = t_p
    rbo = rb[:n]; rbo.orthonormalize()
    _, t_fm = tm(rbo.forward_modeling, sm=sm, a=y)
    out["forward_modeling_s"] = t_fm
    print(name, geo, N, "K", K, "n", n, {k: (round(v, 4) if isinstance(v, float) else v) for k, v in out.items()}, flush=True)
    if do_oracle:
        o = FEMOracle(geo, N)
        t = time.perf_counter(); Uo = o.generate_solutions(y[:20]); t_o = (time.perf_counter() - t) / 20
        err = np.max(np.linalg.norm(U[:20] - Uo, axis=1) / np.linalg.norm(Uo, axis=1))
        res = {"oracle_solve_s_per_system_1core": round(t_o, 4), "snapshot_rel_err_vs_oracle": float(err)}
        for crit in (GREEDY_FOR_GALERKIN, GREEDY_FOR_H10):
            t = time.perf_counter()
            _, _, pk, trace = greedy_build(o, n, U, y, o.H10norm(U), greedy_for=crit, return_trace=True)
            res[f"oracle_greedy_{crit}_s"] = round(time.perf_counter() - t, 2)
            same = pk == picks[crit]
            gaps = [float(np.sort(e)[-1] - np.sort(e)[-2]) / float(np.sort(e)[-1]) for e in trace]
            res[f"indices_identical_{crit}"] = same
            if not same:
                first = next(i for i, (a_, b_) in enumerate(zip(pk, picks[crit])) if a_ != b_)
                res[f"first_diff_{crit}"] = (first, pk[first], picks[crit][first], gaps[first])
        t = time.perf_counter(); co, so, _ = pca_components(U, n); res["oracle_pca_s"] = round(time.perf_counter() - t, 2)
        res["pca_sv_rel_err"] = float(np.max(np.abs(np.asarray(rbp.singular_values_) - so) / so))
        print("   oracle:", res, flush=True)
