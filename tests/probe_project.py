"""Timing probe: reduced operators Phi A_q Phi^T, edge-difference kernel vs stencil apply + split-K DMMA product."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from romhighcontrast_b200.engine import Engine
for geo, N, n in (((4, 4), 64, 20), ((4, 4), 64, 8), ((8, 8), 64, 20), ((4, 4), 64, 64)):
    eng = Engine(geo, N)
    Phi = torch.randn(n, eng.Dp, dtype=torch.float64, device="cuda")
    for v in (0, 1):
        eng.set_option("proj_variant", v)
        eng.project_operators(Phi); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            eng.project_operators(Phi)
        e1.record(); torch.cuda.synchronize()
        print(geo, N, n, "dmma" if v else "edge", "%.3f ms" % (e0.elapsed_time(e1) / 10), flush=True)
