"""Probe (not a test): BASELINE configs[3] at full scale -- state and parameter estimation from m = 50 point measurements
over a 100 000-observation batch on the (4,4), N = 64 model; stage times on the GPU and parity of a sub-sample against
the CPU oracle's restatement of ReducedBasis.py:65-86 / Estimators.py:24-37."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from lib.SolutionsManagers import SolutionsManagerFEM
from lib.ReducedBasis import ReducedBasisGreedy
from oracle import FEMOracle
from oracle.rb import state_estimation, estimator_inv, estimator_linear

geo, N, n, m, Kobs, Ktrain = (4, 4), 64, 20, 50, 100000, 1000
sm = SolutionsManagerFEM(geo, N)
eng = sm._engine_()
def tm(f, *a, **k):
    torch.cuda.synchronize(); t = time.perf_counter(); r = f(*a, **k); torch.cuda.synchronize(); return r, time.perf_counter() - t
ytr = 10 ** np.random.default_rng(42).uniform(0, 6, (Ktrain,) + geo)
Utr = sm.generate_solutions(ytr)
rb, t_rb = tm(ReducedBasisGreedy().build, n=n, sm=sm, solutions2train=Utr, a2train=ytr, solutions2train_h1norm=sm.H10norm(Utr))
pts = np.random.default_rng(1).uniform(low=[sm.x_domain[0], sm.y_domain[0]], high=[sm.x_domain[1], sm.y_domain[1]], size=(m, 2))
yobs = 10 ** np.random.default_rng(44).uniform(0, 6, (Kobs,) + geo)
# observations: snapshot solves + point evaluation, chunked on the device (the (K, D) fields are never materialised)
def observe():
    Z = torch.empty((Kobs, m), dtype=torch.float64, device=eng.device)
    x = eng.empty(10000, eng.Dp)
    for k0 in range(0, Kobs, 10000):
        eng.solve(eng.params(yobs[k0:k0 + 10000]), out=x)
        Z[k0:k0 + 10000] = eng.evaluate(pts, x)
    return Z
Zd, t_obs = tm(observe)
Z = Zd.cpu().numpy()
c, t_se = tm(rb.state_estimation, sm, pts, Z, reconstruct=False)
inv, t_inv = tm(rb.parameter_estimation_inverse, c)
lin, t_lin = tm(rb.parameter_estimation_linear, c)
print(f"greedy basis n={n} on {Ktrain} snapshots: {t_rb:.3f} s; {Kobs} observations (solve + evaluate at {m} points): {t_obs:.3f} s "
      f"({Kobs / t_obs:.0f} /s); LS state estimation coefficients: {t_se * 1e3:.1f} ms; parameter estimation inverse {t_inv * 1e3:.1f} ms, "
      f"linear {t_lin * 1e3:.1f} ms", flush=True)
# parity on a sub-sample
o = FEMOracle(geo, N)
sel = np.arange(0, Kobs, Kobs // 64)[:64]
Uo = o.generate_solutions(yobs[sel[:4]])
Zo = o.evaluate_solutions(pts, Uo)
print("measurements rel err vs oracle (4 systems):", float(np.abs(Z[sel[:4]] - Zo).max() / np.abs(Zo).max()))
co, _ = state_estimation(o, np.asarray(rb.basis), pts, Z[sel])
print("coefficients rel err vs oracle lstsq:", float(np.linalg.norm(c[:, sel] - co) / np.linalg.norm(co)), "cond(E) =",
      float(np.linalg.cond(o.evaluate_solutions(pts, np.asarray(rb.basis)))))
print("estimator_inv rel err:", float(np.abs(inv[sel] - estimator_inv(c[:, sel], np.asarray(rb.a))).max() / np.abs(inv[sel]).max()),
      "estimator_linear rel err:", float(np.abs(lin[sel] - estimator_linear(c[:, sel], np.asarray(rb.a))).max() / np.abs(lin[sel]).max()))
err = np.abs(1 - inv / yobs)
print("median |1 - a_hat / a| (inverse estimator):", float(np.median(err)))
