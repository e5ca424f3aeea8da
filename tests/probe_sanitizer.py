import sys
sys.path.insert(0, ".")
import numpy as np, torch
from romhighcontrast_b200.engine import Engine
for geo, N, K in (((2, 2), 16, 8), ((4, 4), 64, 4)):
    eng = Engine(geo, N)
    y = eng.params(10 ** np.random.default_rng(1).uniform(0, 6, (K,) + geo))
    x, it, rel = eng.solve(y)
    torch.cuda.synchronize()
    print(geo, N, K, "iterations", it.tolist(), "max relres", float(rel.max()))
