"""Multi-GPU layer: one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch) for the plumbing.

The parameter training set shards naturally: snapshot solves, reduced solves, point evaluation and
state / parameter estimation are independent per parameter vector, so every rank works on a contiguous slice and
no data-path collective is needed.  Only two stages exchange data:

  * greedy:  all_gather of one (error, global index) pair per rank and step, merged with np.argmax semantics
             (first maximum wins, NaN is maximal), then a broadcast of the selected snapshot (8 D bytes) and its
             parameter from the owning rank;
  * POD:     one all_to_all that turns the K-sharded snapshot matrix into a D-sharded one (full columns local, so
             the column mean needs no further exchange), a local centred SYRK (fp64 DMMA) over the D slice, one
             all_reduce of the K x K partial Gram matrices, a replicated small eigensolve and an all_gather of the
             (n, D / G) back-projected component slices.  For K ~ 10^5 (configs[4]) the Gram-free route
             (method="krylov") replaces all of that by one (b, Dp) all_reduce per block-Lanczos step.

All collective helpers work on CPU tensors with the gloo backend as well; tests/test_dist_cpu.py runs them with
world_size 2 on the host.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.distributed as dist


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard_bounds(K: int, nshards: int):
    """Contiguous, balanced slices: shard r owns [b[r], b[r+1])."""
    return [(K * r) // nshards for r in range(nshards + 1)]


def local_slice(K: int, r=None, w=None):
    r = rank() if r is None else r
    w = world() if w is None else w
    b = shard_bounds(K, w)
    return slice(b[r], b[r + 1])


def column_bounds(Dp: int, nshards: int, align: int = 8):
    """Column slices of the padded layout, aligned to `align` doubles (64 bytes) so TMA / cp.async stay aligned."""
    units = Dp // align
    b = [((units * r) // nshards) * align for r in range(nshards + 1)]
    b[-1] = Dp
    return b


def merge_argmax(vals, idxs):
    """np.argmax over the concatenation, given each shard's (local max value, its GLOBAL index); idx < 0 = empty shard."""
    best_v, best_i, best_r = None, -1, -1
    for r, (v, i) in enumerate(zip(vals, idxs)):
        if i < 0:
            continue
        v = float(v)
        if best_i < 0:
            better = True
        else:
            vn, bn = math.isnan(v), math.isnan(best_v)
            if vn != bn:
                better = vn
            elif vn and bn:
                better = i < best_i
            else:
                better = v > best_v or (v == best_v and i < best_i)
        if better:
            best_v, best_i, best_r = v, int(i), r
    return best_v, best_i, best_r


def global_argmax(local_val: float, local_global_idx: int, device=None):
    """all_gather one (value, global index) pair per rank -> (value, global index, owner rank), identical on all ranks."""
    if world() == 1:
        return float(local_val), int(local_global_idx), 0
    dev = device if device is not None else torch.device("cpu")
    v = torch.tensor([float(local_val)], dtype=torch.float64, device=dev)
    i = torch.tensor([int(local_global_idx)], dtype=torch.int64, device=dev)
    vs = [torch.empty_like(v) for _ in range(world())]
    is_ = [torch.empty_like(i) for _ in range(world())]
    dist.all_gather(vs, v)
    dist.all_gather(is_, i)
    return merge_argmax([float(t.item()) for t in vs], [int(t.item()) for t in is_])


def broadcast_from(t: torch.Tensor, owner: int):
    if world() > 1:
        dist.broadcast(t, src=owner)
    return t


def k_to_d_shards(X_local: torch.Tensor, counts=None):
    """(K_r, Dp) K-sharded -> (K, Dp_r) D-sharded with one all_to_all; rows keep their global order.

    counts[r] = number of rows held by rank r (default: all equal to X_local.shape[0])."""
    w, me = world(), rank()
    if w == 1:
        return X_local
    Kl, Dp = X_local.shape
    counts = [Kl] * w if counts is None else list(counts)
    cb = column_bounds(Dp, w)
    wme = cb[me + 1] - cb[me]
    if X_local.is_cuda:
        # one flat exchange: the send buffer holds this rank's rows column block by column block, the receive buffer IS the
        # (K, Dp_r) result (rank j's rows are a contiguous run of counts[j] * Dp_r doubles): no per-peer tensors, no cat
        send = torch.empty(Kl * Dp, dtype=X_local.dtype, device=X_local.device)
        o = 0
        for j in range(w):
            wj = cb[j + 1] - cb[j]
            send[o:o + Kl * wj].view(Kl, wj).copy_(X_local[:, cb[j]:cb[j + 1]])
            o += Kl * wj
        out = torch.empty(int(sum(counts)), wme, dtype=X_local.dtype, device=X_local.device)
        dist.all_to_all_single(out.view(-1), send, output_split_sizes=[c * wme for c in counts],
                               input_split_sizes=[Kl * (cb[j + 1] - cb[j]) for j in range(w)])
        return out
    send = [X_local[:, cb[j]:cb[j + 1]].contiguous() for j in range(w)]
    recv = [torch.empty(counts[j], wme, dtype=X_local.dtype, device=X_local.device) for j in range(w)]
    _all_to_all_fallback(recv, send)
    return torch.cat(recv, dim=0)


def _all_to_all_fallback(recv, send):
    """gloo has no all_to_all: w broadcasts-worth of point-to-point traffic via all_gather of padded blocks (tests only)."""
    w, me = world(), rank()
    for src in range(w):
        for dst in range(w):
            if src == dst:
                if me == src:
                    recv[src].copy_(send[dst])
                continue
            if me == src:
                dist.send(send[dst], dst=dst)
            elif me == dst:
                dist.recv(recv[src], src=src)


def all_gather_cols(t_local: torch.Tensor, Dp: int):
    """(n, Dp_r) column slices -> (n, Dp) on every rank."""
    w = world()
    if w == 1:
        return t_local
    cb = column_bounds(Dp, w)
    parts = [torch.empty(t_local.shape[0], cb[j + 1] - cb[j], dtype=t_local.dtype, device=t_local.device)
             for j in range(w)]
    dist.all_gather(parts, t_local.contiguous()) if len({p.shape for p in parts}) == 1 else _all_gather_uneven(parts, t_local)
    return torch.cat(parts, dim=1)


def _all_gather_uneven(parts, t_local):
    for src in range(world()):
        buf = t_local.contiguous() if src == rank() else parts[src]
        dist.broadcast(buf, src=src)
        if src == rank():
            parts[src].copy_(buf)


# --------------------------------------------------------------------------------------------------------------
# sharded POD and greedy on top of the device engine
# --------------------------------------------------------------------------------------------------------------
def distributed_pca(eng, X_local_pad: torch.Tensor, n: int, counts=None, timings=None, method="gram"):
    """PCA(n) of the K-sharded padded snapshots; returns (components (n, Dp), singular values (n,)) on every rank.

    method="gram": all_to_all + partial SYRK + all_reduce of the K x K Gram (the north-star's fp64 Gram GEMM; right for
    K ~ 10^4).  method="krylov": Gram-free block Lanczos on the K-sharded rows as they are (pod.krylov_pca): one
    all_reduce of (b, Dp) doubles per step, no all_to_all, no K x K matrix -- the route for K ~ 10^5 (configs[4]: the
    Gram alone would be 80 GB and 2.6e15 flop)."""
    if method == "krylov":
        from .pod import krylov_pca
        K_total = None if counts is None else int(sum(counts))
        comps, sig, _ = krylov_pca(eng, X_local_pad, n, K_total=K_total, stats=timings)
        return comps, sig
    if method != "gram":
        raise ValueError(f"distributed_pca: unknown method {method!r}")
    from .pod import top_eigenpairs
    Dp = X_local_pad.shape[1]
    ev = (lambda: torch.cuda.Event(enable_timing=True)) if X_local_pad.is_cuda else None
    marks = []

    def mark(name):
        if timings is not None and ev is not None:
            e = ev(); e.record(); marks.append((name, e))

    mark("start")
    Xs = k_to_d_shards(X_local_pad, counts).contiguous()            # (K, Dp_r)
    mark("all_to_all")
    mean = eng.column_mean(Xs)
    eng.center_rows_(Xs, mean)
    G = eng.gemm_nt(Xs, Xs, symmetric=True)                         # partial Gram over this rank's columns
    mark("gram_partial")
    if world() > 1:
        dist.all_reduce(G)
    mark("gram_allreduce")
    lam, V = top_eigenpairs(eng, G, n)
    mark("eigensolve")
    lam = torch.clamp(lam, min=0.0)
    sig = torch.sqrt(lam)
    comps_local = eng.gemm_tn(V.contiguous(), Xs) / torch.where(sig > 0, sig, torch.ones_like(sig))[:, None]
    mark("backproject")
    comps = all_gather_cols(comps_local, Dp)
    mark("components_allgather")
    idx = comps.abs().argmax(dim=1)
    sign = torch.sign(comps[torch.arange(comps.shape[0], device=comps.device), idx])
    sign = torch.where(sign == 0, torch.ones_like(sign), sign)
    if timings is not None and marks:
        torch.cuda.synchronize()
        for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
            timings[n1] = e0.elapsed_time(e1)
    return comps * sign[:, None], sig


def greedy_build_sharded(sm, n, U_local, a_local, h1_local, K_total, greedy_for="galerkin", timings=None):
    """Weak greedy (reference ReducedBasis.py:112-139) on a K-sharded training set.

    Every rank passes its contiguous slice (local_slice(K_total)) -- numpy arrays, or the padded device tensors the
    solver left behind ((K_r, Dp) snapshots, (K_r, nb) parameters, (K_r,) norms) -- and all ranks return the same
    (basis (n, D), a list, global indices).  The loop is greedy_core.greedy_select: errors never leave the device; per
    step one all_gather of a (value, global index) pair per rank, merged on the device, and one all_reduce that
    delivers the winning snapshot and its parameter.  timings (dict, optional): wall milliseconds of the device sweep,
    the exchange and the Gram-Schmidt step, summed over the n steps (synchronised stage by stage when given)."""
    from .greedy_core import greedy_select
    eng = sm._engine_()
    off = shard_bounds(K_total, world())[rank()]
    D, geo = sm.vspace_dim, tuple(sm.blocks_geometry)
    if isinstance(U_local, torch.Tensor):
        U = U_local if U_local.shape[0] else None
        y = a_local.reshape(U_local.shape[0], -1).contiguous() if U is not None else None
        norm = h1_local.expand(U_local.shape[0]).contiguous() if U is not None else None
    else:
        U_local = np.asarray(U_local, dtype=np.float64)
        a_local = np.asarray(a_local, dtype=np.float64)
        U = eng.pad(U_local) if len(U_local) else None
        y = eng.params(a_local) if len(a_local) else None
        norm = eng.dev(np.broadcast_to(np.asarray(h1_local, dtype=np.float64), (len(U_local),)).copy()) if len(U_local) else None
    picked, _, rows, params, _ = greedy_select(sm, eng, n, U, y, norm, greedy_for, offset=off, distributed=True, timings=timings)
    basis = eng.unpad(rows).cpu().numpy().reshape(n, D)
    pa = params.cpu().numpy().reshape((n,) + geo)
    return basis, [pa[i] for i in range(n)], picked
