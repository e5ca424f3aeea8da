"""Probe (not a test): BASELINE configs[4] geometry ((8,8) subdomains, N=64: 512^2 cells, D = 261 121) on one GPU, both kernel families."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from romhighcontrast_b200.engine import Engine
from oracle import FEMOracle
geo, N, K = (8, 8), 64, 1024
y = 10 ** np.random.default_rng(7).uniform(0, 6, (K,) + geo)
for tile in (1, 0):
    eng = Engine(geo, N)
    eng.set_option("tile", tile)
    yd = eng.params(y); x = eng.empty(K, eng.Dp)
    eng.solve(yd, out=x); torch.cuda.synchronize()
    t = time.perf_counter(); _, it, rel = eng.solve(yd, out=x); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"tile={tile}: {K / dt:.0f} solves/s, iterations mean {it.double().mean().item():.2f} max {it.max().item()}", flush=True)
    if tile == 1:
        U = eng.unpad(x[:2]).cpu().numpy()
        Uo = FEMOracle(geo, N).generate_solutions(y[:2])
        print("rel l2 err vs oracle", np.linalg.norm(U - Uo, axis=1) / np.linalg.norm(Uo, axis=1), flush=True)
    del eng, x
