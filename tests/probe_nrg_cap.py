"""Probe: most row groups per tile-kernel region (option tile_nrg_cap, default 16).  With 32 a 64-row level (level 2 of the
256^2 mesh, 16 column groups) becomes ONE region per system instead of two overlapping ones."""
import ctypes as C, sys
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from romhighcontrast_b200 import _lib
from romhighcontrast_b200.engine import Engine
K = 10000
y_host = bench.sample_params(K, 42)
x_ref = None
for cap in (16, 20, 32, 16, 20):
    eng = Engine((4, 4), 64)
    eng.set_option("tile_nrg_cap", cap)
    y = eng.params(y_host); x = eng.empty(K, eng.Dp)
    eng.solve(y, out=x); eng.solve(y, out=x)
    eng.set_option("profile", 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.solve(y, out=x); eng.solve(y, out=x); e1.record(); torch.cuda.synchronize()
    pms, pn = (C.c_double * 8)(), (C.c_int64 * 8)()
    _lib.check(eng.lib.romhc_get_profile(eng.handle, pms, pn))
    _, it, rr = eng.solve(y, out=x)
    xs = x[:256].clone()
    if x_ref is None:
        x_ref = xs
    dev = float(((xs - x_ref).norm(dim=1) / x_ref.norm(dim=1)).max())
    print(f"tile_nrg_cap={cap}: {2 * K / e0.elapsed_time(e1) * 1e3:8.0f} solves/s, per-kind ms {[round(pms[i] / max(pn[i], 1), 3) for i in range(7)]}, "
          f"iterations mean {float(it.double().mean()):.3f}, max rel. difference to the first variant {dev:.2e}", flush=True)
    del eng, x, y
    torch.cuda.empty_cache()
