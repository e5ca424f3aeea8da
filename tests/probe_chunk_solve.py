"""Probe: GPU time of resident solves in the chunk sizes of the host pipeline (is chunking itself the e2e gap?)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from romhighcontrast_b200.engine import Engine
K = 10000
eng = Engine((4, 4), 64)
y = eng.params(bench.sample_params(K, 42)); x = eng.empty(K, eng.Dp)
def run(sizes):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); k0 = 0
    for s in sizes:
        eng.solve(y[k0:k0 + s], out=x[k0:k0 + s]); k0 += s
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)
for sizes in ([10000], [5000, 5000], [2500] * 4, [2500, 2500, 2500, 1250, 625, 625], [1250] * 8, [625] * 16):
    run(sizes); print(len(sizes), "chunks:", "%.1f ms" % run(sizes), flush=True)
u = eng.empty(K, eng.D)
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng.unpad(x); e1.record(); torch.cuda.synchronize(); print("unpack 10000: %.2f ms" % e0.elapsed_time(e1))
