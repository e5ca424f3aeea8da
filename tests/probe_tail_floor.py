"""Probe: the per-iteration floor of the batched PCG (what a late iteration with few active systems costs), the whole
step at K = 10 000 and the per-kernel times of the all-active iterations.  ROMHC_LIB_PATH selects the build (A/B)."""
import ctypes as C
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from romhighcontrast_b200 import _lib
from romhighcontrast_b200.engine import Engine
eng = Engine((4, 4), 64)
yall = eng.params(bench.sample_params(10000, 42))
x = eng.empty(10000, eng.Dp)
for _ in range(2):
    eng.solve(yall, out=x)
torch.cuda.synchronize()
ev = lambda: torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(3):
    e0, e1 = ev(), ev()
    e0.record(); _, it, _ = eng.solve(yall, out=x); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print("K=10000 step ms:", ["%.2f" % t for t in ts], "launched iterations", eng.last_solve_stats["launched_iterations"],
      "mean", float(it.double().mean()), flush=True)
eng.set_option("profile", 1)
eng.solve(yall, out=x)
pms, pn = (C.c_double * 8)(), (C.c_int64 * 8)()
_lib.check(eng.lib.romhc_get_profile(eng.handle, pms, pn))
eng.set_option("profile", 0)
print("per-kind avg ms (all-active iterations):", ["%.3f" % (pms[i] / max(pn[i], 1)) for i in range(8)], list(pn), flush=True)
for K in (8, 64, 512, 2500):
    y = yall[:K].contiguous()
    eng.solve(y, out=x[:K]); torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record(); _, it, _ = eng.solve(y, out=x[:K]); e1.record(); torch.cuda.synchronize()
    n_launched = eng.last_solve_stats["launched_iterations"]
    print(f"K={K}: {e0.elapsed_time(e1):.2f} ms, launched iterations {n_launched}, max {int(it.max())} mean {float(it.double().mean()):.2f} -> {e0.elapsed_time(e1) / max(n_launched, 1):.3f} ms per launched iteration", flush=True)
