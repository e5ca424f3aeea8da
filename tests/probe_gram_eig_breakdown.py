"""Probe: where the time of pod.top_eigenpairs (block Lanczos on the Gram matrix, K = 10 000) goes, per primitive."""
import sys, time, collections
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from romhighcontrast_b200.engine import Engine
from romhighcontrast_b200 import pod
eng = Engine((4, 4), 64)
K, n = 10000, 20
x, _, _ = eng.solve(eng.params(bench.sample_params(K, 42)))
mean = eng.column_mean(x); eng.center_rows_(x, mean)
G = eng.gemm_nt(x, x, symmetric=True)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    lam, V = pod.top_eigenpairs(eng, G, n); torch.cuda.synchronize()
    print("plain total %.1f ms" % (1e3 * (time.perf_counter() - t0)), flush=True)
T = collections.defaultdict(float); N = collections.defaultdict(int)
def wrap(obj, name, label=None):
    f = getattr(obj, name)
    def g(*a, **k):
        torch.cuda.synchronize(); t = time.perf_counter(); r = f(*a, **k); torch.cuda.synchronize()
        T[label or name] += time.perf_counter() - t; N[label or name] += 1; return r
    setattr(obj, name, g)
for nm in ("gemm_nt", "gemm_nn", "gemm_tn", "tsqr_r", "row_norms"):
    wrap(eng, nm)
wrap(np.linalg, "eigh", "host_eigh"); wrap(np.linalg, "svd", "host_svd")
torch.cuda.synchronize(); t0 = time.perf_counter()
lam, V = pod.top_eigenpairs(eng, G, n); torch.cuda.synchronize()
tot = time.perf_counter() - t0
print("instrumented total %.1f ms" % (1e3 * tot))
for k in sorted(T, key=T.get, reverse=True):
    print("  %-12s %7.2f ms in %3d calls" % (k, 1e3 * T[k], N[k]))
print("  other %.2f ms" % (1e3 * (tot - sum(T.values()))))
