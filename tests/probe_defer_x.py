"""Probe: deferred update of the iterate (option defer_x; x touched every second PCG iteration) against the plain update at
configs[2]: solves/s, time of the fused update + going-down kernel, and bit-equality of the solutions."""
import ctypes as C, sys
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from romhighcontrast_b200 import _lib
from romhighcontrast_b200.engine import Engine
K = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
y_host = bench.sample_params(K, 42)
eng = Engine((4, 4), 64)
y = eng.params(y_host); x = eng.empty(K, eng.Dp)
x_ref = None
for mode, pp in ((0, 1), (1, 1), (0, 1), (1, 1), (1, 2)):
    eng.set_option("defer_x", mode); eng.set_option("papply_pers", pp)
    eng.solve(y, out=x); eng.solve(y, out=x)
    eng.set_option("profile", 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.solve(y, out=x); eng.solve(y, out=x); e1.record(); torch.cuda.synchronize()
    pms, pn = (C.c_double * 8)(), (C.c_int64 * 8)()
    _lib.check(eng.lib.romhc_get_profile(eng.handle, pms, pn))
    eng.set_option("profile", 0)
    _, it, rr = eng.solve(y, out=x)
    if x_ref is None:
        x_ref = x.clone()
    same = bool(torch.equal(x, x_ref))
    dev = float(((x - x_ref).norm(dim=1) / x_ref.norm(dim=1)).max())
    print(f"defer_x={mode} papply_pers={pp}: {2 * K / e0.elapsed_time(e1) * 1e3:8.0f} solves/s, update+down {pms[2] / max(pn[2], 1):.3f} ms, "
          f"p_apply {pms[0] / max(pn[0], 1):.3f} ms, iterations mean {float(it.double().mean()):.3f} max {int(it.max())} "
          f"(odd last iteration: {int((it % 2 == 1).sum())}), bit-identical to the first variant: {same}, max rel. difference {dev:.2e}", flush=True)
