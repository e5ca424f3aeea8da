"""Device-resident core of the weak greedy (reference `ReducedBasisGreedy.build`, /root/reference/src/lib/ReducedBasis.py:112-139).

One round = reduced operators of the current orthonormal basis -> K reduced solves (Galerkin) or K projections (H10)
-> fused error sweep || c Phi - u ||_{A_1} over the resident snapshots -> argmax with np.argmax semantics -> the winner
joins the basis.  Nothing in a round touches the host: the selected index stays a device scalar (row gather by
index_select), and the orthonormal basis grows by ONE classical Gram-Schmidt step with re-orthogonalisation (two passes,
split-K DMMA products) instead of the reference's QR of all rows.  The reference re-sorts the rows by contrast before its
QR (`sort_orthogonalize_base`, :24-29); that ordering only changes WHICH orthonormal basis of span{selected snapshots}
comes out, and both greedy criteria depend on the span alone (the reduced Galerkin solution and the H10 projection are
basis independent), so the selection is the same up to rounding-level ties; what the reference stores (`basis`, `a`,
`selected_indices`) are the raw snapshots, which are copied bit for bit.

Sharded training sets: every rank runs the same loop on its rows; per round ONE all_gather of a
(value, global index) pair per rank, merged on the device with np.argmax semantics (first maximum wins, NaN is maximal),
and ONE all_reduce that delivers the winning row and its parameter (the owner contributes the row, everybody else zeros:
x + 0 is exact).  No host synchronisation inside the loop either way; indices and errors come back once at the end.
"""
from __future__ import annotations

import time

import numpy as np
import torch

GREEDY_FOR_H10 = r"$H^1_0$"
GREEDY_FOR_GALERKIN = "galerkin"


def _merge_pairs_device(allp):
    """allp (w, 2): rows (local max value, its GLOBAL index or -1) -> (value, index) of np.argmax over the concatenation."""
    vals, idxs = allp[:, 0], allp[:, 1]
    valid = idxs >= 0
    isn = torch.isnan(vals) & valid
    key = torch.where(isn, torch.full_like(vals, float("inf")), vals)
    key = torch.where(valid, key, torch.full_like(vals, -float("inf")))
    best = key.max()
    cand = torch.where(isn.any(), isn, (key == best) & valid)
    win = torch.where(cand, idxs, torch.full_like(idxs, float("inf"))).min()
    val = torch.where(isn.any(), torch.full_like(best, float("nan")), best)
    return val, win


def greedy_select(sm, eng, n, U, y, norm, greedy_for, offset=0, distributed=False, timings=None, progress=None):
    """U (K_r, Dp) padded device snapshots of this rank (None or 0 rows: empty shard), y (K_r, nb), norm (K_r,).

    Returns (picked: list of n global indices, max_errors: list of n floats, rows (n, Dp) device, params (n, nb) device,
    Q (n, Dp): the Euclidean-orthonormal basis of the span)."""
    import torch.distributed as dist
    w = dist.get_world_size() if distributed and dist.is_available() and dist.is_initialized() else 1
    dev = eng.device
    Kr = 0 if U is None else int(U.shape[0])
    Dp, nb = eng.Dp, eng.nb
    f64 = torch.float64
    Q = torch.zeros(n, Dp, dtype=f64, device=dev)
    rows = torch.zeros(n, Dp, dtype=f64, device=dev)
    params = torch.zeros(n, nb, dtype=f64, device=dev)
    picked = torch.zeros(n, dtype=f64, device=dev)
    maxerr = torch.zeros(n, dtype=f64, device=dev)
    bad = torch.zeros((), dtype=torch.bool, device=dev)
    # incremental right-hand sides of the H10 criterion (device engines only; the CPU test double projects from scratch)
    Bh = ones = None
    if greedy_for == GREEDY_FOR_H10 and Kr and hasattr(eng, "apply"):
        Bh = torch.zeros(Kr, n, dtype=f64, device=dev)
        ones = torch.ones(Kr, nb, dtype=f64, device=dev)
    sync = (lambda: torch.cuda.synchronize()) if (timings is not None and dev.type == "cuda") else (lambda: None)
    tacc = {"sweep_ms": 0.0, "exchange_ms": 0.0, "orthonormalise_ms": 0.0}
    it = range(n) if progress is None else progress(range(n))
    for k in it:
        sync(); t0 = time.perf_counter()
        if Kr:
            if k == 0:
                err = eng.error_norm(U, None, None)                       # approximation == 0 (reference :89-91, :109-111)
            else:
                Phi = Q[:k]
                if greedy_for == GREEDY_FOR_H10 and Bh is not None:
                    # H10 projection (:122, SolutionsManagers.py:108-139): right-hand sides B[:, j] = U A_1 q_j.  The basis grows
                    # by one Gram-Schmidt vector per round and the old vectors never change, so only the NEW column is a
                    # pass over U (one GEMV) instead of a (K, Dp) x (Dp, k) product
                    w_new = eng.apply(None, Q[k - 1:k])
                    Bh[:, k - 1] = eng.row_dots(U, w_new)                  # GEMV: U streamed once at HBM speed
                    Ahat, _ = eng.project_operators(Phi)
                    Cc = eng.reduced_solve(ones, Ahat, Bh[:, :k].contiguous())
                elif greedy_for == GREEDY_FOR_H10:
                    Cc = sm._projection_coefficients_dev(eng, U, Phi)     # :122
                elif greedy_for == GREEDY_FOR_GALERKIN:
                    Ahat, bhat = eng.project_operators(Phi)               # :124
                    Cc, info = eng.reduced_solve(y, Ahat, bhat, check=False, return_info=True)
                    bad = bad | info.any()
                else:
                    raise Exception(f"Not implemented greedy for {greedy_for}")
                err = eng.error_norm(U, Cc, Phi)
            li, lv = eng.argmax_dev(err / norm)                            # :129 (true division: the round-1 tie is exactly 1.0)
            gi = (li + offset).to(f64)
        else:
            li = torch.zeros(1, dtype=torch.int64, device=dev)
            lv = torch.zeros(1, dtype=f64, device=dev)
            gi = torch.full((1,), -1.0, dtype=f64, device=dev)
        sync(); t1 = time.perf_counter()
        if w > 1:
            pair = torch.cat((lv.reshape(1), gi.reshape(1)))
            allp = torch.empty(w, 2, dtype=f64, device=dev)
            if dev.type == "cuda":
                dist.all_gather_into_tensor(allp, pair.reshape(1, 2))
            else:
                parts = [torch.empty(2, dtype=f64) for _ in range(w)]
                dist.all_gather(parts, pair)
                allp = torch.stack(parts)
            val, win = _merge_pairs_device(allp)
            mine = (gi.reshape(()) == win)
            buf = torch.zeros(Dp + nb, dtype=f64, device=dev)
            if Kr:
                buf[:Dp] = torch.where(mine, U.index_select(0, li.reshape(1)).reshape(-1), buf[:Dp])
                buf[Dp:] = torch.where(mine, y.index_select(0, li.reshape(1)).reshape(-1), buf[Dp:])
            dist.all_reduce(buf)
            row, par = buf[:Dp], buf[Dp:]
        else:
            val, win = lv.reshape(()), gi.reshape(())
            row = U.index_select(0, li.reshape(1)).reshape(-1)
            par = y.index_select(0, li.reshape(1)).reshape(-1)
        picked[k] = win
        maxerr[k] = val
        rows[k] = row
        params[k] = par
        sync(); t2 = time.perf_counter()
        # one Gram-Schmidt step, twice (Euclidean, as np.linalg.qr in the reference's orthonormalize_base, :18-21)
        v = row.reshape(1, Dp).clone()
        if k:
            Qk = Q[:k]
            for _ in range(2):
                c = eng.gemm_nt(Qk, v, splitk=True)                        # (k, 1) = Q v
                v = v - eng.gemm_nn(c.T.contiguous(), Qk)                  # (1, Dp)
        Q[k] = (v / eng.l2_norm(v)).reshape(-1)
        sync(); t3 = time.perf_counter()
        tacc["sweep_ms"] += 1e3 * (t1 - t0); tacc["exchange_ms"] += 1e3 * (t2 - t1); tacc["orthonormalise_ms"] += 1e3 * (t3 - t2)
    if timings is not None:
        timings.update(tacc)
    if bool(bad.item()):
        raise np.linalg.LinAlgError("reduced Galerkin matrix is not positive definite")
    return [int(i) for i in picked.cpu().tolist()], [float(v) for v in maxerr.cpu().tolist()], rows, params, Q
