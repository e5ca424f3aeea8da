"""Probe (torchrun, N GPUs): per-stage times of dist.distributed_pca on a 10k-snapshot union, 4 repetitions."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, ".")
import bench
local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from romhighcontrast_b200 import dist as rd
from romhighcontrast_b200.engine import Engine
w, r = rd.world(), rd.rank()
eng = Engine((4, 4), 64)
K = 10000
b = rd.shard_bounds(K, w); Kl = b[r + 1] - b[r]
x, _, _ = eng.solve(eng.params(bench.sample_params(Kl, 42 + r)))
counts = [b[i + 1] - b[i] for i in range(w)]
dims = []
_eigh = np.linalg.eigh
def eigh_logged(a):
    t = time.perf_counter(); out = _eigh(a); dims.append((a.shape[0], round(1e3 * (time.perf_counter() - t), 1))); return out
np.linalg.eigh = eigh_logged
for rep in range(4):
    tm = {}
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    comps, sig = rd.distributed_pca(eng, x.clone(), 20, counts=counts, timings=tm)
    torch.cuda.synchronize()
    if r == 0:
        print("eigh dims/ms", dims); dims.clear()
        print(rep, "%.1f ms" % (1e3 * (time.perf_counter() - t0)), {k: round(v, 2) for k, v in tm.items()}, flush=True)
dist.destroy_process_group()
