"""Probe: strip height / CTA size of the persistent search-direction kernel after its clean-up (options strip_kb, threads):
the kernel's time and solves/s at configs[2]."""
import ctypes as C, sys
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from romhighcontrast_b200 import _lib
from romhighcontrast_b200.engine import Engine
K = 10000
y_host = bench.sample_params(K, 42)
for kb, th in ((113, 512), (75, 512), (56, 512), (150, 512), (227, 512), (113, 256), (75, 256), (56, 256), (113, 512)):
    eng = Engine((4, 4), 64)
    eng.set_option("strip_kb", kb); eng.set_option("threads", th)
    y = eng.params(y_host); x = eng.empty(K, eng.Dp)
    try:
        eng.solve(y, out=x); eng.solve(y, out=x)
        eng.set_option("profile", 1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.solve(y, out=x); eng.solve(y, out=x); e1.record(); torch.cuda.synchronize()
        pms, pn = (C.c_double * 8)(), (C.c_int64 * 8)()
        _lib.check(eng.lib.romhc_get_profile(eng.handle, pms, pn))
        print(f"strip_kb={kb} threads={th}: {2 * K / e0.elapsed_time(e1) * 1e3:8.0f} solves/s, k_pcg_p_apply {pms[0] / max(pn[0], 1):.3f} ms", flush=True)
    except Exception as exc:
        print(f"strip_kb={kb} threads={th}: {exc!r}"[:200], flush=True)
    del eng, x, y
    torch.cuda.empty_cache()
