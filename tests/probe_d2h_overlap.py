"""Probe (not a test): D2H copy bandwidth while the batched solver is running on the same GPU."""
import sys, threading, time
import numpy as np, torch
sys.path.insert(0, '.')
import bench
from romhighcontrast_b200.engine import Engine

eng = Engine((4, 4), 64)
K = 6000
yd = eng.params(bench.sample_params(K, 42)); x = eng.empty(K, eng.Dp)
src = torch.empty(1300 * 1024 * 1024 // 8, dtype=torch.float64, device='cuda')
dst = torch.empty_like(src, device='cpu').pin_memory()
copy_stream = torch.cuda.Stream()
eng.solve(yd, out=x); torch.cuda.synchronize()

def copier(tag, reps=3):
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(copy_stream):
            e0.record(); dst.copy_(src, non_blocking=True); e1.record()
        e1.synchronize()
        print(tag, 'D2H GB/s', src.numel() * 8 / (e0.elapsed_time(e1) * 1e-3) / 1e9, flush=True)

copier('idle')
t = threading.Thread(target=copier, args=('busy', 4))
t0 = time.perf_counter()
t.start()
for _ in range(3):
    eng.solve(yd, out=x)
torch.cuda.synchronize()
print('3 solves of', K, 'took ms', (time.perf_counter() - t0) * 1e3, flush=True)
t.join()
