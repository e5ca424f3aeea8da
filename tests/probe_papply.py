"""Probe: variant of k_pcg_p_apply at configs[2] (papply_pers 0: one CTA per strip, 1: persistent with the fp64 stencil form,
2: persistent with the fp32 combination + edge form): solves/s, the kernel's time, iteration counts, agreement of the solutions."""
import ctypes as C, sys
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from romhighcontrast_b200 import _lib
from romhighcontrast_b200.engine import Engine
K = 10000
y_host = bench.sample_params(K, 42)
x_ref = None
for pers, kb in ((1, 113), (2, 113), (1, 113), (2, 113)):
    eng = Engine((4, 4), 64)
    eng.set_option("papply_pers", pers); eng.set_option("strip_kb", kb)
    y = eng.params(y_host); x = eng.empty(K, eng.Dp)
    eng.solve(y, out=x); eng.solve(y, out=x)
    eng.set_option("profile", 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.solve(y, out=x); eng.solve(y, out=x); e1.record(); torch.cuda.synchronize()
    pms, pn = (C.c_double * 8)(), (C.c_int64 * 8)()
    _lib.check(eng.lib.romhc_get_profile(eng.handle, pms, pn))
    _, it, rr = eng.solve(y, out=x)
    xs = x[:64].clone()
    if x_ref is None:
        x_ref = xs
    dev = float(((xs - x_ref).norm(dim=1) / x_ref.norm(dim=1)).max())
    print(f"papply_pers={pers} strip_kb={kb}: {2 * K / e0.elapsed_time(e1) * 1e3:8.0f} solves/s, k_pcg_p_apply {pms[0] / max(pn[0], 1):.3f} ms, "
          f"iterations mean {float(it.double().mean()):.3f} max {int(it.max())}, max relres {float(rr.max()):.2e}, "
          f"max rel. difference to the first variant (64 systems) {dev:.2e}", flush=True)
    del eng, x, y
    torch.cuda.empty_cache()
