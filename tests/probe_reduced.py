"""Timing probe of the online stage: 1M reduced Galerkin solves (n = 20, nb = 16) per kernel variant; run on the GPU box."""
import ctypes as C
import json
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from romhighcontrast_b200 import _lib
from romhighcontrast_b200.engine import Engine

eng = Engine((4, 4), 8)
dev = torch.device("cuda")
out = {}
for n, nb, K in ((20, 16, 1000000), (10, 16, 1000000), (16, 16, 1000000), (24, 16, 1000000), (20, 64, 500000), (32, 16, 200000)):
    rng = np.random.default_rng(n)
    Bm = rng.standard_normal((nb, n, n))
    Ahat = torch.as_tensor(np.einsum("qij,qkj->qik", Bm, Bm) + 0.1 * np.eye(n), device=dev)
    y = torch.as_tensor(10 ** rng.uniform(0, 6, (K, nb)), device=dev)
    rhs = torch.as_tensor(rng.standard_normal(n), device=dev)
    Cc = torch.empty(K, n, dtype=torch.float64, device=dev)
    info = torch.empty(K, dtype=torch.int32, device=dev)
    for occ in (0,):
        call = lambda: _lib.call("romhc_reduced_solve", C.c_void_p(y.data_ptr()), nb, C.c_void_p(Ahat.data_ptr()),
                                 C.c_void_p(rhs.data_ptr()), 0, n, K, C.c_void_p(Cc.data_ptr()), C.c_void_p(info.data_ptr()), None)
        call(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            call()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        flop = K * (2 * nb * n * (n + 1) / 2 + n ** 3 / 3 + 2 * n * n)
        sel = slice(0, 2000)
        Ak = np.einsum("kq,qij->kij", y[sel].cpu().numpy(), Ahat.cpu().numpy())
        ref = np.linalg.solve(Ak, np.broadcast_to(rhs.cpu().numpy(), (2000, n))[..., None])[..., 0]
        err = float(np.linalg.norm(Cc[sel].cpu().numpy() - ref) / np.linalg.norm(ref))
        out[f"n{n}_nb{nb}_K{K}_occ{occ}"] = {"ms": ms, "TFLOPs": flop / ms / 1e9, "relerr": err, "bad": int(info.sum())}
        print(n, nb, K, occ, "%.3f ms" % ms, "%.2f TF" % (flop / ms / 1e9), "err %.1e" % err, flush=True)
json.dump(out, open("gpurun_out/r2_reduced_probe.json", "w"), indent=1)
