"""Probe: strip height (strip_kb) and variant (papply_pers) of k_pcg_p_apply at configs[2]: solves/s and the kernel's time."""
import ctypes as C, sys
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from romhighcontrast_b200 import _lib
from romhighcontrast_b200.engine import Engine
K = 10000
y_host = bench.sample_params(K, 42)
for pers, kb in ((0, 113), (1, 113), (1, 227)):
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
    print(f"papply_pers={pers} strip_kb={kb}: {2 * K / e0.elapsed_time(e1) * 1e3:8.0f} solves/s, k_pcg_p_apply {pms[0] / max(pn[0], 1):.3f} ms", flush=True)
    del eng, x, y
    torch.cuda.empty_cache()
