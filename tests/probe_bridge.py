"""Non-nested ("bridge") coarsening: iteration counts and time per solve with and without it.
Usage: python tests/probe_bridge.py            (prints one line per geometry)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import torch
from romhighcontrast_b200.engine import Engine

def run(geo, N, K, bridge):
    eng = Engine(geo, N)
    eng.set_option("bridge", bridge)
    y = eng.params(10 ** np.random.default_rng(42).uniform(0, 6, (K,) + geo))
    x, it, rel = eng.solve(y)
    dt = 1e9
    for _ in range(3):                       # best of three: the first geometry runs while the clocks still ramp up
        torch.cuda.synchronize()
        t = time.perf_counter()
        x, it, rel = eng.solve(y)
        torch.cuda.synchronize()
        dt = min(dt, time.perf_counter() - t)
    info = (C.c_int64 * 16)()
    eng.lib.romhc_get_info(eng.handle, info)
    return x, it.double().mean().item(), int(it.max()), K / dt, int(info[5]), int(info[14]), int(info[15])

import ctypes as C
for geo, N, K in [((3, 3), 43, 1000), ((3, 3), 43, 4000), ((4, 4), 20, 2000), ((3, 3), 44, 1000), ((2, 3), 27, 2000), ((4, 4), 63, 4000), ((2, 2), 45, 2000)]:
    x1, m1, mx1, s1, nl1, bl1, bn1 = run(geo, N, K, 1)
    x0, m0, mx0, s0, nl0, bl0, bn0 = run(geo, N, K, 0)
    d = (torch.linalg.vector_norm(x1 - x0, dim=1) / torch.linalg.vector_norm(x0, dim=1)).max().item()
    print(f"{geo} N={N} K={K}: bridge level {bl1} -> N'={bn1} ({nl1} levels): {m1:.2f} it (max {mx1}), {s1:.0f} solves/s | "
          f"without ({nl0} levels): {m0:.2f} it (max {mx0}), {s0:.0f} solves/s | max rel diff {d:.2e}", flush=True)
