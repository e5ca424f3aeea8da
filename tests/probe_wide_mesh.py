"""Probe (not a test): BASELINE configs[4] geometry ((8,8) subdomains, N=64: 512^2 cells, D = 261 121) on one GPU:
kernel families and smoothing schedules (nu on the finest level / nu_mid / nu_tail).
Usage: python tests/probe_wide_mesh.py [K]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from romhighcontrast_b200.engine import Engine
from oracle import FEMOracle
geo, N = (8, 8), 64
K = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
y = 10 ** np.random.default_rng(7).uniform(0, 6, (K,) + geo)
for tile, nu, nu_mid, nu_tail in ((1, 2, 3, 4), (1, 1, 3, 4), (1, 1, 2, 4), (1, 1, 2, 3), (1, 3, 3, 4), (0, 2, 3, 4), (0, 1, 3, 4)):
    eng = Engine(geo, N)
    eng.set_option("tile", tile)
    eng.set_option("nu", nu); eng.set_option("nu_mid", nu_mid); eng.set_option("nu_tail", nu_tail)
    yd = eng.params(y); x = eng.empty(K, eng.Dp)
    eng.solve(yd, out=x); torch.cuda.synchronize()
    t = time.perf_counter(); _, it, rel = eng.solve(yd, out=x); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"tile={tile} nu={nu}/{nu_mid}/{nu_tail}: {K / dt:.0f} solves/s, iterations mean {it.double().mean().item():.2f} max {it.max().item()}", flush=True)
    if tile == 1 and nu == 2:
        U = eng.unpad(x[:2]).cpu().numpy()
        Uo = FEMOracle(geo, N).generate_solutions(y[:2])
        print("rel l2 err vs oracle", np.linalg.norm(U - Uo, axis=1) / np.linalg.norm(Uo, axis=1), flush=True)
    del eng, x
