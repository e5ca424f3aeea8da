"""Measurement script (not a test): the two tall-skinny products of one S-apply of pod.krylov_pca on random data.

    python tests/probe_pod_kernels.py [K] [Dp]       (default 10000 x 65792, the configs[2] snapshot matrix)
Used under ncu (profiles/): k_gemm_nt<128,32,...,SPLITK> = X W^T and k_gemm_tn_mma<4> = Y^T X, plus their reductions."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from romhighcontrast_b200.engine import Engine            # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
Dp = int(sys.argv[2]) if len(sys.argv) > 2 else 65792
eng = Engine((2, 2), 4)
X = torch.randn(K, Dp, dtype=torch.float64, device=eng.device)
W = torch.randn(32, Dp, dtype=torch.float64, device=eng.device)
for _ in range(2):
    Y = eng.gemm_nt(X, W, splitk=True)
    Z = eng.gemm_tn(Y, X)
torch.cuda.synchronize()
print("ok", float(Z[0, 0]))
