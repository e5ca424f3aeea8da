"""Probe (not a test): where the time of the POD eigensolve goes at K = 10 000 (block Lanczos on the Gram matrix)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import bench
from romhighcontrast_b200.engine import Engine
from romhighcontrast_b200 import pod

eng = Engine((4, 4), 64)
K, n = 10000, 20
x, _, _ = eng.solve(eng.params(bench.sample_params(K, 42)))
mean = eng.column_mean(x); eng.center_rows_(x, mean)
G = eng.gemm_nt(x, x, symmetric=True)
torch.cuda.synchronize()
T = {"gq": 0.0, "eigh": 0.0, "qr": 0.0, "n_gq": 0, "n_eigh": 0}
orig_gemm = eng.gemm_nt
def timed_gemm(A, B, symmetric=False):
    torch.cuda.synchronize(); t = time.perf_counter(); r = orig_gemm(A, B, symmetric); torch.cuda.synchronize()
    T["gq"] += time.perf_counter() - t; T["n_gq"] += 1; return r
eng.gemm_nt = timed_gemm
orig_eigh = np.linalg.eigh
def timed_eigh(a):
    t = time.perf_counter(); r = orig_eigh(a); T["eigh"] += time.perf_counter() - t; T["n_eigh"] += 1; T.setdefault("dims", []).append(a.shape[0]); return r
np.linalg.eigh = timed_eigh
orig_qr = torch.linalg.qr
def timed_qr(a):
    torch.cuda.synchronize(); t = time.perf_counter(); r = orig_qr(a); torch.cuda.synchronize(); T["qr"] += time.perf_counter() - t; return r
torch.linalg.qr = timed_qr
for rep in range(2):
    for k in ("gq", "eigh", "qr"): T[k] = 0.0
    T["n_gq"] = T["n_eigh"] = 0; T["dims"] = []
    torch.cuda.synchronize(); t0 = time.perf_counter()
    lam, V = pod.top_eigenpairs(eng, G, n)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"total {dt*1e3:.1f} ms: G@Q {T['gq']*1e3:.1f} ms in {T['n_gq']} products, host eigh {T['eigh']*1e3:.1f} ms dims {T['dims']}, qr {T['qr']*1e3:.1f} ms", flush=True)
print("sigma head", torch.sqrt(lam[:5]).cpu().numpy())
