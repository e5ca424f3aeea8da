import sys, torch
sys.path.insert(0, ".")
from romhighcontrast_b200.engine import Engine
import ctypes as C
from romhighcontrast_b200 import _lib
for geo, N, K in (((4, 4), 64, 10000), ((8, 8), 64, 2000), ((3, 3), 43, 4000)):
    eng = Engine(geo, N)
    x = torch.randn(K, eng.Dp, dtype=torch.float64, device="cuda"); u = eng.empty(K, eng.D)
    def f(): _lib.check(eng.lib.romhc_unpack(eng.handle, C.c_void_p(x.data_ptr()), C.c_void_p(u.data_ptr()), K, eng.stream()))
    def g(): _lib.check(eng.lib.romhc_pack(eng.handle, C.c_void_p(u.data_ptr()), C.c_void_p(x.data_ptr()), K, eng.stream()))
    for nm, fn in (("unpack", f), ("pack", g)):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); [fn() for _ in range(5)]; e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(geo, N, K, nm, "%.3f ms" % ms, "%.0f GB/s" % (2 * K * eng.D * 8 / ms / 1e6), flush=True)
    del eng, x, u
