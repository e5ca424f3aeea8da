"""Probe: end-to-end rate of romhc_generate_solutions_host (pinned buffers) against the chunk count of its pipeline."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from romhighcontrast_b200.engine import Engine
K = 10000
eng = Engine((4, 4), 64)
y_host = bench.sample_params(K, 42)
U = torch.empty((K, eng.D), dtype=torch.float64, pin_memory=True).numpy()
y = torch.from_numpy(np.ascontiguousarray(y_host.reshape(K, -1))).pin_memory().numpy()
for hc in (4, 2, 3, 4, 5, 6, 8, 12):
    eng.set_option("host_chunks", hc)
    eng.generate_solutions_host(y, out=U)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(2):
        eng.generate_solutions_host(y, out=U)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 2
    print(f"host_chunks={hc}: {K / dt:8.0f} solves/s ({dt * 1e3:.1f} ms)", flush=True)
