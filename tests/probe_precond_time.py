"""Probe (not a test): time of one V-cycle (romhc_precond) on K systems of the 256^2 mesh."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch, bench
from romhighcontrast_b200.engine import Engine
eng = Engine((4, 4), 64); K = 4000
y = eng.params(bench.sample_params(K, 42)); r = torch.randn(K, eng.Dp, dtype=torch.float64, device="cuda")
r = eng.pad(eng.unpad(r))
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); z = eng.precond(y, r); e1.record(); torch.cuda.synchronize()
    print("precond ms", round(e0.elapsed_time(e1), 3), flush=True)
