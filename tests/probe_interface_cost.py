"""Probe (not a test): cost of the general (interface-row) path of the tile kernels: same 256^2 grid as (1,1) x N=256 (no
interfaces), (4,4) x N=64, (16,16) x N=16 (an interface row in every tile row group)."""
import sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, '.')
from romhighcontrast_b200 import _lib
from romhighcontrast_b200.engine import Engine
import bench
K = 2000
for geo, N in (((1, 1), 256), ((4, 4), 64), ((16, 16), 16)):
    eng = Engine(geo, N)
    y = eng.params(10 ** np.random.default_rng(42).uniform(0, 6, (K,) + geo))
    x = eng.empty(K, eng.Dp)
    eng.solve(y, out=x)
    eng.set_option("profile", 1)
    _, it, _ = eng.solve(y, out=x)
    pms, pn = (C.c_double * 8)(), (C.c_int64 * 8)()
    _lib.check(eng.lib.romhc_get_profile(eng.handle, pms, pn))
    print(geo, N, "iterations", round(it.double().mean().item(), 2),
          {nm: round(pms[i] / pn[i], 3) for i, nm in enumerate(bench.KIND_NAMES) if pn[i] > 0}, flush=True)
    del eng
