"""POD by the method of snapshots on the device (replaces sklearn PCA, /root/reference/src/lib/ReducedBasis.py:196).

    mean -> centre -> G = Xc Xc^T (fp64 DMMA SYRK, libromhc gemm_nt) -> top-n eigenpairs -> components = V^T Xc / sigma

The eigensolve is the "small" part: host LAPACK for K <= 1024, otherwise blocked subspace iteration whose only
large operation, G @ Q, is again the DMMA kernel (G is read once per iteration).  Signs follow sklearn's
svd_flip(u_based_decision=False): the entry of largest magnitude of every component is positive.
"""
from __future__ import annotations

import numpy as np
import torch


def top_eigenpairs(eng, G, n, extra=12, tol=1e-13, max_dim=960, seed=0, host_max=1024):
    """Largest n eigenpairs of the symmetric PSD device matrix G (K, K): (lam (n,), V (K, n)) device tensors.

    K <= host_max: LAPACK on the host.  Otherwise block Lanczos with full (twice-applied) reorthogonalisation and
    block size b = n + extra <= 32: every step costs one G @ block product -- the DMMA kernel, G is streamed once --
    plus O(K * dim * b) orthogonalisation work; the Rayleigh-Ritz problem on the accumulated Krylov basis (dimension
    <= max_dim) is solved on the host.  Stops when ||G v - lam v|| <= tol * lam_1 for the n wanted pairs.  (POD
    spectra of high-dimensional parameter sets decay slowly: plain subspace iteration stalls and an aggressive
    polynomial filter wipes out the smaller wanted directions in fp64; a Krylov basis does neither.)"""
    K = G.shape[0]
    n = min(n, K)
    if K <= host_max:
        lam, V = np.linalg.eigh(G.cpu().numpy())
        order = np.argsort(lam)[::-1][:n]
        return (torch.as_tensor(np.ascontiguousarray(lam[order]), device=G.device),
                torch.as_tensor(np.ascontiguousarray(V[:, order]), device=G.device))
    b = int(min(32, K, n + extra))
    gen = torch.Generator(device=G.device).manual_seed(seed)
    GQ = lambda X: eng.gemm_nt(G, X.T.contiguous())          # (K, b) = G X  (G symmetric)
    Q = torch.linalg.qr(torch.randn(K, b, dtype=torch.float64, device=G.device, generator=gen))[0]
    V = torch.empty(K, max_dim, dtype=torch.float64, device=G.device)      # Krylov basis
    Z = torch.empty(K, max_dim, dtype=torch.float64, device=G.device)      # G @ basis
    dim = 0
    lam = vec = None
    while True:
        V[:, dim:dim + b] = Q
        Zj = GQ(Q)
        Z[:, dim:dim + b] = Zj
        dim += b
        Vd, Zd = V[:, :dim], Z[:, :dim]
        if dim >= 2 * b and (dim // b) % 2 == 0 or dim + b > max_dim:
            T = (Vd.T @ Zd).cpu().numpy()                     # small (dim x dim) projected matrix
            w, S = np.linalg.eigh(0.5 * (T + T.T))
            order = np.argsort(w)[::-1][:n]
            lam = torch.as_tensor(np.ascontiguousarray(w[order]), device=G.device)
            Sd = torch.as_tensor(np.ascontiguousarray(S[:, order]), device=G.device)
            vec = Vd @ Sd
            res = torch.linalg.vector_norm(Zd @ Sd - vec * lam[None, :], dim=0)
            if float(res.max()) <= tol * max(float(lam[0]), 1e-300) or dim + b > max_dim:
                break
        W = Zj
        for _ in range(2):                                    # block Gram-Schmidt against the whole basis, twice
            W = W - Vd @ (Vd.T @ W)
        Q, R = torch.linalg.qr(W)
        if float(R.diagonal().abs().min()) < 1e-300:          # invariant subspace found
            continue
    return lam, vec.contiguous()


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
