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

    K <= host_max: LAPACK on the host (the "small eigensolve").  Otherwise the same block Lanczos as the Gram-free route
    (`_block_lanczos`: full reorthogonalisation through split-K DMMA products, rank-revealing orthonormalisation by the
    device TSQR, Rayleigh-Ritz on the host for a basis of <= max_dim rows) with S = G: every step costs one G @ block
    product -- the DMMA kernel, G is streamed once.  Stops when every wanted pair has ||G v - lam v|| <= max(1e-10 lam_i,
    tol lam_1) or the residuals stall at the rounding level of G (a Gram matrix summed over ranks carries a slightly
    higher floor than one computed in one piece).  No library GEMM / QR call anywhere on this path."""
    K = G.shape[0]
    n = min(n, K)
    if K <= host_max:
        lam, V = np.linalg.eigh(G.cpu().numpy())
        order = np.argsort(lam)[::-1][:n]
        return (torch.as_tensor(np.ascontiguousarray(lam[order]), device=G.device),
                torch.as_tensor(np.ascontiguousarray(V[:, order]), device=G.device))
    b = int(min(32, K, n + extra))
    gen = torch.Generator(device=G.device).manual_seed(seed)
    R0 = torch.randn(b, K, dtype=torch.float64, device=G.device, generator=gen)
    apply_G = lambda Q: eng.gemm_nt(G, Q).T.contiguous()            # rows: (G Q^T)^T = Q G, G read once by the skinny DMMA kernel
    lam, comps, _ = _block_lanczos(eng, apply_G, R0, n, rtol=1e-10, floor=tol, max_dim=int(min(max_dim, K)))
    if comps.shape[0] < n:                                        # rank of G below n: zero eigenvalues, zero vectors
        pad = n - comps.shape[0]
        lam = torch.cat((lam, torch.zeros(pad, dtype=torch.float64, device=G.device)))
        comps = torch.cat((comps, torch.zeros(pad, K, dtype=torch.float64, device=G.device)))
    return lam, comps.T.contiguous()


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


def _orthonormal_rows(eng, W, thr, passes=2, conditioned=False):
    """Orthonormal rows spanning the directions of W (b, Dp) whose singular value exceeds thr (None if there are none).

    Rank revealing and built from row combinations of W only, so slots that are zero in every row stay exactly zero.
    First pass: R of a Householder QR of W^T (device TSQR, `romhc_tsqr_r`; backward stable: singular values resolved down to eps * sigma_max, where a
    b x b Gram matrix would lose everything below sqrt(eps)), SVD R = U S V^T on the host, rows (1 / s_i) v_i^T W for
    s_i > thr; exhausted Krylov directions (rounding noise) are dropped instead of being normalised into vectors that
    are no longer orthogonal to the basis.  The division amplifies rounding by sigma_max / s_i, so a second pass restores
    orthonormality to machine precision; its input is nearly orthonormal already (condition number O(1)), which is
    exactly where the b x b Gram matrix W W^T -- one split-K DMMA product instead of a library QR -- is accurate:
    rows lam_i^-1/2 e_i^T W.  conditioned=True: the caller knows W is nearly orthonormal (re-normalisation after a
    Gram-Schmidt pass), every pass takes the Gram form."""
    for p in range(passes):
        if p == 0 and not conditioned:
            R = eng.tsqr_r(W)                                   # (b, b): Householder TSQR on the device (csrc/dense.cu)
            _, sv, Vt = np.linalg.svd(R.cpu().numpy())
            keep = sv > thr
            if not keep.any():
                return None
            Tm = Vt[keep] / sv[keep, None]
        else:
            Gs = eng.gemm_nt(W, W, splitk=True).cpu().numpy()
            ev, E = np.linalg.eigh(0.5 * (Gs + Gs.T))
            keep = ev > (thr * thr if p == 0 else 0.25)
            if not keep.any():
                return None
            Tm = (E[:, keep] / np.sqrt(ev[keep])).T[::-1]
        W = eng.gemm_nn(torch.as_tensor(np.array(Tm, order="C", copy=True), device=W.device), W)
    return W


def _block_lanczos(eng, apply_S, R0, n, rtol=1e-10, floor=2e-14, max_dim=960, w=1):
    """Block Lanczos with full reorthogonalisation for the n leading eigenpairs of a symmetric PSD operator S that is
    only APPLIED: apply_S maps a row block (b', L) to (b', L).  R0 (b, L): random start block.  Rows are the basis
    vectors throughout (Krylov basis V, S V = Z); all large products are libromhc kernels.  w > 1: the caller's apply_S
    all-reduces over ranks, every rank runs the same arithmetic and rank 0's block / stop decisions are broadcast.
    Returns (lam (n',), components (n', L) rows, info dict)."""
    dev = R0.device
    b, Dp = R0.shape
    def agree(Qb):
        """Sharded runs: every rank continues with rank 0's block.  The replicated arithmetic is deterministic, so this
        changes nothing in practice; it turns 'all ranks hold the same basis' from an expectation into a guarantee -- a
        rank that kept a different number of rows would otherwise hang the next all_reduce."""
        if w == 1:
            return Qb
        cnt = torch.tensor([0 if Qb is None else Qb.shape[0]], dtype=torch.int64, device=dev)
        torch.distributed.broadcast(cnt, src=0)
        c = int(cnt.item())
        if c == 0:
            return None
        if Qb is None or Qb.shape[0] != c:
            Qb = torch.empty(c, Dp, dtype=torch.float64, device=dev)
        Qb = Qb.contiguous()
        torch.distributed.broadcast(Qb, src=0)
        return Qb

    def ritz(Vd, Zd):
        T = eng.gemm_nt(Vd, Zd, splitk=True).cpu().numpy()    # (dim, dim) projected operator
        wv, S = np.linalg.eigh(0.5 * (T + T.T))
        order = np.argsort(wv)[::-1][:n]
        lam = torch.as_tensor(np.ascontiguousarray(wv[order]), device=dev)
        St = torch.as_tensor(np.ascontiguousarray(S[:, order].T), device=dev)      # (n, dim)
        comps = eng.gemm_nn(St, Vd)
        res = eng.row_norms((eng.gemm_nn(St, Zd) - comps * lam[:, None]).contiguous())
        lam1 = max(float(lam[0]), 1e-300)
        excess = float((res / torch.clamp(torch.maximum(rtol * lam, torch.full_like(lam, floor * lam1)), min=1e-300)).max())
        return lam, comps, float(res.max()), excess, lam1

    # start block: S applied to a random block.  It lies in the row space of Xc, so slots that are zero in every
    # snapshot (the padded grid's Dirichlet / alignment slots) are exactly zero in every basis row and component.
    S0 = apply_S(R0)
    Q = _orthonormal_rows(eng, S0, 1e-12 * float(eng.row_norms(S0).max()))
    if Q is None:                                             # Xc == 0: every singular value is zero
        Q = _orthonormal_rows(eng, R0, 0.0)
    Q = agree(Q)
    V = torch.empty(max_dim, Dp, dtype=torch.float64, device=dev)      # Krylov basis, rows
    Z = torch.empty(max_dim, Dp, dtype=torch.float64, device=dev)      # S applied to the basis rows
    dim = steps = 0
    lam = comps = None
    scale = 0.0                                               # running estimate of lambda_1
    prev_res = float("inf")
    flag = torch.zeros(1, dtype=torch.int64, device=dev)
    while True:
        bq = Q.shape[0]
        V[dim:dim + bq] = Q
        Zj = apply_S(Q)
        Z[dim:dim + bq] = Zj
        dim += bq
        steps += 1
        Vd, Zd = V[:dim], Z[:dim]
        scale = max(scale, float(eng.row_norms(Zj).max()))
        # next block: S Q orthogonalised against the whole basis (block Gram-Schmidt, repeated), directions below the
        # requested accuracy dropped
        Qn = None
        if dim + 1 <= max_dim and scale > 0.0:
            Wn = Zj
            for sweep in range(2):
                for _ in range(2):
                    Wn = Wn - eng.gemm_nn(eng.gemm_nt(Vd, Wn, splitk=True).T.contiguous(), Vd)
                Wn = _orthonormal_rows(eng, Wn, 1e-15 * scale if sweep == 0 else 0.5, passes=2 if sweep == 0 else 1,
                                       conditioned=sweep > 0)
                if Wn is None:
                    break
            if Wn is not None:
                Qn = Wn[:max_dim - dim].contiguous()
        Qn = agree(Qn)
        last = Qn is None
        if (steps >= 2 and (dim <= 256 or steps % 2 == 0)) or last:
            lam, comps, res, excess, lam1 = ritz(Vd, Zd)
            stalled = res <= 1e-11 * lam1 and res > 0.5 * prev_res
            prev_res = res
            flag[0] = int(last or excess <= 1.0 or stalled)
            if w > 1:
                torch.distributed.broadcast(flag, src=0)
            if int(flag.item()):
                break
        Q = Qn
    info = {"steps": steps, "krylov_dim": dim, "block": b, "residual": res}
    return lam, comps, info


def krylov_pca(eng, X_local_pad, n, K_total=None, center_in_place=False, extra=12, rtol=1e-10, floor=2e-14, max_dim=960,
               seed=0, stats=None, distributed=True):
    """PCA(n) without the K x K Gram matrix: block Lanczos on S = Xc^T Xc (D x D), applied as two tall-skinny products.

    For K >> 10^4 (BASELINE configs[4]: K = 100 000, D = 261 121) the Gram route costs K^2 D = 2.6e15 flop, an
    all_to_all of the whole snapshot set (209 GB) and an 80 GB matrix; the n = 20 leading pairs need none of it.
    Every Lanczos step applies S to a block W (b, Dp), b = n + extra <= 32, as

        Y_r = X_r W^T          (K_r, b)   gemm_nt, fp64 DMMA, X_r streamed once
        Z   = sum_r Y_r^T X_r  (b, Dp)    gemm_tn, X_r streamed once more; ONE all_reduce of b * Dp doubles (67 MB)

    on the K-sharded snapshots exactly as the solver left them (no transpose, no Gram): 4 K D b flop and two passes
    over X per step, a few dozen steps.  The Krylov basis (rows of length Dp) and all orthogonalisation work are
    replicated -- every rank performs the same arithmetic on the same all-reduced data, and rank 0's stop decision is
    broadcast -- so the Ritz vectors are the principal components themselves on every rank: no back-projection, no
    division by sigma.  The products against the basis have a tiny output and a contraction of length Dp: they run
    on the split-K form of the DMMA kernel.  Same conventions as pca_components (sklearn PCA, ReducedBasis.py:196):
    Euclidean, mean-centred with the GLOBAL column mean, singular values sqrt(lambda),
    svd_flip(u_based_decision=False) signs.

    Stops when every wanted Ritz pair has ||S v_i - lam_i v_i|| <= max(rtol * lam_i, floor * lam_1): the residual bounds
    the eigenvalue error, so sigma_i is accurate to rtol / 2 relative (1e-9 is the parity bar) down to the modes whose
    lam_i / lam_1 reaches the rounding level of S itself (floor); also when the residuals stop improving below
    1e-11 lam_1, or the basis reaches max_dim rows.

    X_local_pad (K_r, Dp): this rank's rows (K_r may be 0 on some ranks as long as K_total > 0); distributed=False
    treats them as the whole set even inside an initialised process group (rank-local POD, no collective).
    Returns (components (n, Dp), singular_values (n,), mean (Dp,)) on every rank."""
    from . import dist as rd
    w = rd.world() if distributed else 1
    Kr, Dp = X_local_pad.shape
    dev = X_local_pad.device
    if K_total is None:
        kt = torch.tensor([Kr], dtype=torch.int64, device=dev)
        if w > 1:
            torch.distributed.all_reduce(kt)
        K_total = int(kt.item())
    if K_total <= 0:
        raise ValueError("krylov_pca: empty snapshot set")
    X = X_local_pad if center_in_place else X_local_pad.clone()
    colsum = eng.column_mean(X) * float(Kr) if Kr else torch.zeros(Dp, dtype=torch.float64, device=dev)
    if w > 1:
        torch.distributed.all_reduce(colsum)
    mean = colsum / float(K_total)
    if Kr:
        eng.center_rows_(X, mean)
    n = int(min(n, K_total, Dp))
    if n <= 0:
        return (torch.empty(0, Dp, dtype=torch.float64, device=dev), torch.empty(0, dtype=torch.float64, device=dev), mean)
    b = int(min(32, Dp, n + extra))
    max_dim = int(min(max_dim, Dp))

    ar_events = []                                            # (start, stop) CUDA events around every block all_reduce

    def apply_S(W):                                           # (b', Dp) -> (b', Dp), identical on every rank
        if Kr:
            Zw = eng.gemm_tn(eng.gemm_nt(X, W, splitk=True), X)
        else:
            Zw = torch.zeros_like(W)
        if w > 1:
            if stats is not None and Zw.is_cuda:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); torch.distributed.all_reduce(Zw); e1.record()
                ar_events.append((e0, e1, Zw.numel() * 8))
            else:
                torch.distributed.all_reduce(Zw)
        return Zw

    lam, comps, info = _block_lanczos(eng, apply_S, torch.randn(b, Dp, dtype=torch.float64, device=dev,
                                                                generator=torch.Generator(device=dev).manual_seed(seed)),
                                      n, rtol=rtol, floor=floor, max_dim=max_dim, w=w)
    steps, dim, res = info["steps"], info["krylov_dim"], info["residual"]
    if stats is not None:
        stats.update(steps=steps, krylov_dim=dim, block=b, residual=res)
        if ar_events:
            torch.cuda.synchronize()
            ms = [e0.elapsed_time(e1) for e0, e1, _ in ar_events]
            stats.update(block_allreduce_calls=len(ms), block_allreduce_ms_min=min(ms), block_allreduce_ms_median=float(np.median(ms)),
                         block_allreduce_ms_total=float(sum(ms)), block_allreduce_bytes=int(max(nb for _, _, nb in ar_events)))
    sig = torch.sqrt(torch.clamp(lam, min=0.0))
    idx = comps.abs().argmax(dim=1)
    sign = torch.sign(comps[torch.arange(comps.shape[0], device=dev), idx])
    sign = torch.where(sign == 0, torch.ones_like(sign), sign)
    comps = comps * sign[:, None]
    if comps.shape[0] < n:        # rank of Xc below n: the remaining singular values are zero; zero rows, as the Gram
        pad = n - comps.shape[0]  # route's V^T Xc gives for sigma = 0
        comps = torch.cat((comps, torch.zeros(pad, Dp, dtype=torch.float64, device=dev)))
        sig = torch.cat((sig, torch.zeros(pad, dtype=torch.float64, device=dev)))
    return comps.contiguous(), sig, mean
