"""Probe: (K, D) numpy arrays in and out of the device at configs[2] size (K = 10 000, D = 65 025: 5.2 GB) --
torch's plain pageable copy + layout kernel against the pipelined romhc_pack_host / romhc_unpack_host."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from romhighcontrast_b200.engine import Engine
from romhighcontrast_b200 import _lib

K = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
eng = Engine((4, 4), 64)
U = np.random.default_rng(0).standard_normal((K, eng.D))
GB = U.nbytes / 1e9
print("THP:", open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip(), "| cores", os.cpu_count())

def tm(f, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t = time.perf_counter(); r = f(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t)
    return r, best

def plain_pad():
    c = torch.as_tensor(U).to(eng.device)
    out = eng.empty(K, eng.Dp)
    _lib.check(eng.lib.romhc_pack(eng.handle, c.data_ptr(), out.data_ptr(), K, eng.stream()))
    return out

ref, t0 = tm(plain_pad)
got, t1 = tm(lambda: eng.pad(U))
assert torch.equal(ref, got)
print(f"host -> padded device, {GB:.2f} GB pageable: plain {t0*1e3:.0f} ms ({GB/t0:.1f} GB/s), pipelined {t1*1e3:.0f} ms ({GB/t1:.1f} GB/s)")
_, t2 = tm(lambda: eng.unpad(ref).cpu().numpy())
back, t3 = tm(lambda: eng.unpad_host(ref))
np.testing.assert_array_equal(back, U)
print(f"padded device -> fresh numpy: plain {t2*1e3:.0f} ms ({GB/t2:.1f} GB/s), pipelined {t3*1e3:.0f} ms ({GB/t3:.1f} GB/s)")
out = np.empty_like(U)
_, t4 = tm(lambda: eng.unpad_host(ref, out=out))
print(f"padded device -> existing numpy: pipelined {t4*1e3:.0f} ms ({GB/t4:.1f} GB/s)")
