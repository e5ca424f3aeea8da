"""Probe: smoothing schedule at BASELINE configs[2] ((4,4), N = 64, K = 10 000) with the final kernels of round 2 (the
schedule V(2,2)/(3,3)/(4,4) was chosen in round 1 with slower level >= 1 / tail kernels): solves/s and PCG iterations per
(nu, nu_mid, nu_tail)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from romhighcontrast_b200.engine import Engine
geo, N, K = (4, 4), 64, 10000
y_host = bench.sample_params(K, 42)
for nu, nu_mid, nu_tail in ((2, 3, 4), (2, 2, 4), (2, 2, 3), (2, 3, 3), (2, 3, 6), (2, 4, 4), (2, 2, 6), (2, 3, 4)):
    eng = Engine(geo, N)
    eng.set_option("nu", nu); eng.set_option("nu_mid", nu_mid); eng.set_option("nu_tail", nu_tail)
    y = eng.params(y_host)
    x = eng.empty(K, eng.Dp)
    eng.solve(y, out=x); torch.cuda.synchronize()          # warm-up at full size: workspace allocation, kernel attributes
    ts = []
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); _, it, rel = eng.solve(y, out=x); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = min(ts)
    print(f"nu={nu} nu_mid={nu_mid} nu_tail={nu_tail}: {K / ms * 1e3:8.0f} solves/s ({ms:.1f} ms), iterations mean {float(it.double().mean()):.2f} max {int(it.max())}, "
          f"relres max {float(rel.max()):.1e}", flush=True)
    del eng, x, y
    torch.cuda.empty_cache()
