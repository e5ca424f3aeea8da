"""Timing probe: greedy error sweep at configs[2] / configs[4] shapes, DMMA kernel vs strip kernel."""
import sys, json
import numpy as np, torch
sys.path.insert(0, ".")
from romhighcontrast_b200.engine import Engine
out = {}
for geo, N, K, n in (((4, 4), 64, 10000, 20), ((4, 4), 64, 10000, 8), ((8, 8), 64, 12500, 20), ((3, 3), 43, 4000, 20)):
    eng = Engine(geo, N)
    g = torch.Generator(device="cuda").manual_seed(0)
    U = torch.randn(K, eng.Dp, dtype=torch.float64, device="cuda", generator=g)
    Phi = torch.randn(n, eng.Dp, dtype=torch.float64, device="cuda", generator=g)
    C = torch.randn(K, n, dtype=torch.float64, device="cuda", generator=g)
    for sw in (2, 1, 0):
        eng.set_option("sweep", sw)
        eng.error_norm(U, C, Phi); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            eng.error_norm(U, C, Phi)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        gb = K * eng.D * 8 / 1e9
        out[f"{geo}_N{N}_K{K}_n{n}_sweep{sw}"] = {"ms": ms, "GBps_algorithmic": gb / ms * 1e3, "TFLOPs": 2.0 * K * eng.D * n / ms / 1e9}
        print(geo, N, K, n, {2: "dmma-v2", 1: "dmma-v1", 0: "strip"}[sw], "%.3f ms  %.0f GB/s  %.1f TF" % (ms, gb / ms * 1e3, 2.0 * K * eng.D * n / ms / 1e9), flush=True)
    eng.set_option("sweep", 1)
    del U, Phi, C, eng
json.dump(out, open("gpurun_out/r2_sweep_probe.json", "w"), indent=1)
