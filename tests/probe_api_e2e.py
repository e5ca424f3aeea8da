"""Probe (not a test): end-to-end rate of the reference-facing call sm.generate_solutions(a) (numpy in, fresh numpy out)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from lib.SolutionsManagers import SolutionsManagerFEM
geo, N, K = (4, 4), 64, 10000
sm = SolutionsManagerFEM(geo, N, method="lsqsparse")
y = 10 ** np.random.default_rng(42).uniform(0, 6, (K,) + geo)
sm.generate_solutions(y[:64])
for rep in range(3):
    t = time.perf_counter(); U = sm.generate_solutions(y); dt = time.perf_counter() - t
    print(f"sm.generate_solutions: K={K} in {dt * 1e3:.1f} ms = {K / dt:.0f} solves/s (fresh (K, D) float64 output, {U.nbytes / 1e9:.1f} GB)", flush=True)
    del U
