"""Measurement script (not a test): POD of K snapshots on one GPU, Gram route against the Gram-free Krylov route.

    python tests/probe_krylov_pod.py [K] [n] [blocks]     -> one JSON line (also written to gpurun_out/krylov_pod_probe.json)

Default geometry: BASELINE configs[2], (4,4) blocks, N = 64 (D = 65 025); blocks = 8 gives configs[4]'s (8,8) blocks
(D = 261 121; K = 12 500 is the per-GPU share of its 100 000 snapshots on 8 GPUs).  Snapshots are real solves
(contrast 10^U(0,6))."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from romhighcontrast_b200.engine import Engine            # noqa: E402
from romhighcontrast_b200.pod import krylov_pca, pca_components      # noqa: E402


def timed(fn, reps=2):
    out = fn()                                               # warm-up (lazy kernel attributes, scratch growth)
    torch.cuda.synchronize()
    best = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return out, best


def main():
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    nb = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    eng = Engine((nb, nb), 64)
    y = 10 ** np.random.default_rng(42).uniform(0, 6, (K, nb, nb))
    X, _, _ = eng.solve(eng.params(y))
    stats = {}
    (ck, sk, _), ms_k = timed(lambda: krylov_pca(eng, X, n, stats=stats))
    (cg, sg, _), ms_g = timed(lambda: pca_components(eng, X, n))
    W = torch.randn(32, eng.Dp, dtype=torch.float64, device=eng.device)
    _, ms_nt_plain = timed(lambda: eng.gemm_nt(X, W), 3)
    _, ms_nt = timed(lambda: eng.gemm_nt(X, W, splitk=True), 3)
    _, ms_qr = timed(lambda: torch.linalg.qr(W.T, mode="r"), 3)
    Y = eng.gemm_nt(X, W)
    _, ms_tn = timed(lambda: eng.gemm_tn(Y, X), 3)
    eng.set_option("tn_variant", 0)
    _, ms_tn_fma = timed(lambda: eng.gemm_tn(Y, X), 3)
    eng.set_option("tn_variant", 1)
    V = torch.randn(480, eng.Dp, dtype=torch.float64, device=eng.device)
    _, ms_gs = timed(lambda: eng.gemm_nt(V, W, splitk=True), 3)
    _, ms_gs_plain = timed(lambda: eng.gemm_nt(V, W), 3)
    C = torch.randn(32, 480, dtype=torch.float64, device=eng.device)
    _, ms_nn = timed(lambda: eng.gemm_nn(C, V), 3)
    _, ms_T = timed(lambda: eng.gemm_nt(V, V, splitk=True), 3)
    _, ms_T_plain = timed(lambda: eng.gemm_nt(V, V), 3)
    bytes_X = X.numel() * 8
    line = {"K": K, "n": n, "D": eng.D, "krylov_ms": ms_k, "gram_route_ms": ms_g, "krylov": stats,
            "sv_rel_diff": float(((sk - sg).abs() / sg).max()), "comp_abs_diff": float((ck - cg).abs().max()),
            "pad_slots_exact_zero": float((eng.pad(eng.unpad(ck)) - ck).abs().max()) == 0.0,
            "qr_r_only_ms": ms_qr,
            "apply": {"X_Wt_ms": ms_nt, "X_Wt_plain_ms": ms_nt_plain, "Yt_X_ms": ms_tn, "Yt_X_fma_kernel_ms": ms_tn_fma, "X_GBps_nt": bytes_X / ms_nt / 1e6,
                      "X_GBps_tn": bytes_X / ms_tn / 1e6, "TFLOPs_nt": 2e-9 * K * eng.Dp * 32 / ms_nt,
                      "TFLOPs_tn": 2e-9 * K * eng.Dp * 32 / ms_tn},
            "basis_products_dim480": {"V_Wt_splitk_ms": ms_gs, "V_Wt_plain_ms": ms_gs_plain, "C_V_ms": ms_nn,
                                      "V_Vt_splitk_ms": ms_T, "V_Vt_plain_ms": ms_T_plain}}
    print(json.dumps(line))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(line, open("gpurun_out/krylov_pod_probe.json", "w"))


if __name__ == "__main__":
    main()
