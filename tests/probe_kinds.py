"""Per-kernel-kind CUDA-event times of the first PCG iterations (option "profile") for any geometry.
Usage: python tests/probe_kinds.py nrb ncb N K [option=value ...]"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from romhighcontrast_b200.engine import Engine
from romhighcontrast_b200 import _lib
nrb, ncb, N, K = [int(v) for v in sys.argv[1:5]]
geo = (nrb, ncb)
eng = Engine(geo, N)
for kv in sys.argv[5:]:
    k, v = kv.split("=")
    eng.set_option(k, float(v))
y = eng.params(10 ** np.random.default_rng(42).uniform(0, 6, (K,) + geo))
x = eng.empty(K, eng.Dp)
eng.solve(y, out=x)
eng.set_option("profile", 1)
torch.cuda.synchronize()
import time
t = time.perf_counter(); _, it, _ = eng.solve(y, out=x); torch.cuda.synchronize(); dt = time.perf_counter() - t
ms = (C.c_double * 8)(); n = (C.c_int64 * 8)()
_lib.check(eng.lib.romhc_get_profile(eng.handle, ms, n))
names = ["p_apply", "update", "down(l0)", "down(l>=1)", "tail", "up(l0)", "up(l>=1)", "bridge"]
print(geo, N, K, sys.argv[5:], f"{K/dt:.0f} solves/s, {it.double().mean().item():.2f} iterations, Dp={eng.Dp}, levels={eng.nlevels}, tail={eng.tail_level}")
GB = K * eng.Dp * 8 / 1e9
tot = 0.0
for i in range(8):
    if n[i]:
        per_it = ms[i] / n[i] * (n[i] / max(n[0], 1))
        tot += per_it
        print(f"  {names[i]:11s} {ms[i]/n[i]:8.3f} ms/launch x {n[i]/max(n[0],1):.1f} per iteration = {per_it:7.3f} ms")
print(f"  per iteration {tot:.3f} ms; one fine-level vector = {GB:.2f} GB -> {tot and GB/tot*1e3:.0f} GB/s per stream-equivalent")
