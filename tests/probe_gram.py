"""Probe (not a test): the fp64 DMMA Gram kernel alone (K = 10 000 snapshots of the 256^2 mesh), for ncu."""
import sys, torch
sys.path.insert(0, '.')
from romhighcontrast_b200.engine import Engine
eng = Engine((4, 4), 64)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
X = torch.randn(K, eng.Dp, dtype=torch.float64, device='cuda')
Gref = None
for variant in (0, 1):
    eng.set_option("gram_variant", variant)
    for _ in range(2):
        G = eng.gemm_nt(X, X, symmetric=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); G = eng.gemm_nt(X, X, symmetric=True); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if Gref is None:
        Gref = G.clone()
    print(f"variant {variant} K={K}: {ms:.2f} ms, {K * (K + 1) * eng.D / ms / 1e9:.2f} TFLOP/s (triangle, algorithmic D), "
          f"max |G - G0| / max |G0| = {float((G - Gref).abs().max() / Gref.abs().max()):.2e}", flush=True)
