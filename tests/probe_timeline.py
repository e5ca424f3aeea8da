"""Probe (not a test): per-phase clock64 timeline of k_mgp_up<64> items on one SM (needs the TILE_TIMELINE build of the library
copied over romhighcontrast_b200/libromhc.so)."""
import sys, ctypes as C
sys.path.insert(0, ".")
import numpy as np, torch, bench
from romhighcontrast_b200.engine import Engine
eng = Engine((4, 4), 64); K = 4000
y = eng.params(bench.sample_params(K, 42)); x = eng.empty(K, eng.Dp)
eng.solve(y, out=x); torch.cuda.synchronize()
buf = (C.c_longlong * (2 * 24 * 16))()
eng.lib.romhc_debug_timeline.argtypes = [C.c_void_p]
assert eng.lib.romhc_debug_timeline(buf) == 0
t = np.array(buf[:]).reshape(2, 24, 16)
names = ["top", "staged", "copied+published", "S1", "before sweeps", "black1 pub", "S2", "red1 pub", "S3", "black2 pub", "S4", "-", "-", "red2 done", "stored"]
for w in range(2):
    print("warp", w * 8)
    for it in range(0, 23):
        row = t[w, it]
        d = [int(row[i] - row[0]) for i in (1, 11, 2, 3, 12, 4, 5, 6, 7, 8, 9, 10, 13, 14)]
        nxt = int(t[w, it + 1, 0] - row[0])
        print("  item %2d staged %5d regs %5d init+pub %5d S1 %5d settle %5d pre %5d | b1 %5d S2 %5d | r1 %5d S3 %5d | b2 %5d S4 %5d | r2 %5d stored %5d | next %5d" % tuple([it] + d + [nxt]))
