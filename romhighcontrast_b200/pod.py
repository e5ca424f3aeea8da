"""POD by the method of snapshots on the device (replaces sklearn PCA, /root/reference/src/lib/ReducedBasis.py:196).

    mean -> centre -> G = Xc Xc^T (fp64 DMMA SYRK, libromhc gemm_nt) -> top-n eigenpairs -> components = V^T Xc / sigma

The eigensolve is the "small" part: host LAPACK for K <= 1024, otherwise blocked subspace iteration whose only
large operation, G @ Q, is again the DMMA kernel (G is read once per iteration).  Signs follow sklearn's
svd_flip(u_based_decision=False): the entry of largest magnitude of every component is positive.
"""
from __future__ import annotations

import numpy as np
import torch


def top_eigenpairs(eng, G, n, oversample=10, tol=1e-14, maxit=300, seed=0):
    """Largest n eigenpairs of the symmetric PSD device matrix G (K, K): (lam (n,), V (K, n)) device tensors."""
    K = G.shape[0]
    n = min(n, K)
    if K <= 1024:
        lam, V = np.linalg.eigh(G.cpu().numpy())
        order = np.argsort(lam)[::-1][:n]
        return (torch.as_tensor(np.ascontiguousarray(lam[order]), device=G.device),
                torch.as_tensor(np.ascontiguousarray(V[:, order]), device=G.device))
    b = int(min(32, K, n + oversample))
    gen = torch.Generator(device=G.device).manual_seed(seed)
    Q = torch.linalg.qr(torch.randn(K, b, dtype=torch.float64, device=G.device, generator=gen))[0]
    prev = None
    S = None
    lam = None
    stable = 0
    for it in range(maxit):
        Z = eng.gemm_nt(G, Q.T.contiguous())              # (K, b) = G Q   (G symmetric)
        if it % 2 == 1 or it == maxit - 1:
            T = eng.gemm_tn(Q.contiguous(), Z).cpu().numpy()   # (b, b) = Q^T G Q
            w, S_h = np.linalg.eigh(0.5 * (T + T.T))
            order = np.argsort(w)[::-1]
            lam, S = w[order], S_h[:, order]
            if prev is not None:
                rel = np.max(np.abs(lam[:n] - prev[:n]) / np.maximum(np.abs(lam[:n]), 1e-300))
                stable = stable + 1 if rel < tol else 0
                if stable >= 2:
                    break
            prev = lam
        Q = torch.linalg.qr(Z)[0]
    # Rayleigh-Ritz vectors of the last tested subspace (Q before the final re-orthonormalisation spans the same space)
    T = eng.gemm_tn(Q.contiguous(), eng.gemm_nt(G, Q.T.contiguous())).cpu().numpy()
    w, S_h = np.linalg.eigh(0.5 * (T + T.T))
    order = np.argsort(w)[::-1][:n]
    Sd = torch.as_tensor(np.ascontiguousarray(S_h[:, order]), device=G.device)
    V = eng.gemm_nn(Q.contiguous(), Sd)
    return torch.as_tensor(np.ascontiguousarray(w[order]), device=G.device), V


def pca_components(eng, X_pad, n, center_in_place=False):
    """PCA(n_components=n, svd_solver='full') equivalent on padded device snapshots X_pad (K, Dp).

    Returns (components (n, Dp) device, singular_values (n,) device, mean (Dp,) device)."""
    X = X_pad if center_in_place else X_pad.clone()
    mean = eng.column_mean(X)
    eng.center_rows_(X, mean)
    G = eng.gemm_nt(X, X, symmetric=True)
    lam, V = top_eigenpairs(eng, G, n)
    lam = torch.clamp(lam, min=0.0)
    sig = torch.sqrt(lam)
    comps = eng.gemm_tn(V.contiguous(), X)
    comps = comps / torch.where(sig > 0, sig, torch.ones_like(sig))[:, None]
    # svd_flip(u_based_decision=False)
    idx = comps.abs().argmax(dim=1)
    sign = torch.sign(comps[torch.arange(comps.shape[0], device=comps.device), idx])
    sign = torch.where(sign == 0, torch.ones_like(sign), sign)
    return comps * sign[:, None], sig, mean
