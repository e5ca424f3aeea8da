"""Probe: smoothing schedule on the 512-column mesh of BASELINE configs[4] ((8,8), N = 64), where the tile kernels own only
6 of the 16 region rows at V(2,2) (halo 4 nu + 2): solves/s and PCG iterations per (nu, nu_mid, nu_tail)."""
import sys, itertools
import numpy as np, torch
sys.path.insert(0, ".")
from romhighcontrast_b200.engine import Engine
geo, N, K = (8, 8), 64, 3072
y_host = 10 ** np.random.default_rng(4000).uniform(0, 6, size=(K,) + geo)
for nu, nu_mid, nu_tail in ((2, 3, 4), (1, 3, 4), (1, 2, 4), (1, 4, 4), (2, 2, 4), (1, 3, 6)):
    eng = Engine(geo, N)
    eng.set_option("nu", nu); eng.set_option("nu_mid", nu_mid); eng.set_option("nu_tail", nu_tail)
    y = eng.params(y_host)
    x = eng.empty(K, eng.Dp)
    eng.solve(y, out=x); torch.cuda.synchronize()          # warm-up at full size: workspace allocation, kernel attributes
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); _, it, rel = eng.solve(y, out=x); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"nu={nu} nu_mid={nu_mid} nu_tail={nu_tail}: {K / ms * 1e3:8.0f} solves/s, iterations mean {float(it.double().mean()):.2f} max {int(it.max())}, "
          f"relres max {float(rel.max()):.1e}", flush=True)
    del eng, x, y
    torch.cuda.empty_cache()
