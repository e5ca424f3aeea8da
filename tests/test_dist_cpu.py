"""Host-side logic of the multi-GPU path on CPU: gloo backend, world_size 2 (rendezvous on 127.0.0.1)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from romhighcontrast_b200 import dist as rd


def test_shard_and_column_bounds():
    for K, w in [(10, 1), (10, 3), (100000, 8), (7, 8)]:
        b = rd.shard_bounds(K, w)
        assert b[0] == 0 and b[-1] == K and all(x <= y for x, y in zip(b, b[1:]))
        assert max(np.diff(b)) - min(np.diff(b)) <= 1
    for Dp, w in [(65792, 8), (65792, 3), (272, 2), (16770 // 8 * 8, 5)]:
        cb = rd.column_bounds(Dp, w)
        assert cb[0] == 0 and cb[-1] == Dp and all(c % 8 == 0 for c in cb)


def test_merge_argmax_is_np_argmax():
    rng = np.random.default_rng(0)
    for trial in range(200):
        K = int(rng.integers(1, 40))
        v = rng.integers(0, 4, K).astype(float)          # many ties
        if trial % 5 == 0:
            v[rng.integers(0, K)] = np.nan
        w = int(rng.integers(1, 6))
        b = rd.shard_bounds(K, w)
        vals, idxs = [], []
        for r in range(w):
            seg = v[b[r]:b[r + 1]]
            if len(seg) == 0:
                vals.append(0.0); idxs.append(-1)
            else:
                j = int(np.argmax(seg))
                vals.append(seg[j]); idxs.append(b[r] + j)
        _, gi, owner = rd.merge_argmax(vals, idxs)
        assert gi == int(np.argmax(v))
        assert b[owner] <= gi < b[owner + 1]


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # greedy argmax exchange: a tie across ranks must resolve to the lowest global index
        K = 11
        v = np.array([1.0, 3.0, 2.0, 3.0, 0.0, 3.0, 1.0, 3.0, 2.0, 0.5, 3.0])
        sl = rd.local_slice(K)
        seg = v[sl]
        j = int(np.argmax(seg))
        val, gi, owner = rd.global_argmax(seg[j], sl.start + j)
        assert (val, gi, owner) == (3.0, 1, 0), (val, gi, owner)
        row = torch.full((5,), float(rank))
        rd.broadcast_from(row, 1)
        assert torch.equal(row, torch.ones(5))
        # K-sharded -> D-sharded transpose, then column all_gather must reproduce the full matrix
        Kt, Dp = 9, 48
        full = torch.arange(Kt * Dp, dtype=torch.float64).reshape(Kt, Dp)
        b = rd.shard_bounds(Kt, world)
        counts = [b[r + 1] - b[r] for r in range(world)]
        Xs = rd.k_to_d_shards(full[b[rank]:b[rank + 1]].clone(), counts)
        cb = rd.column_bounds(Dp, world)
        assert torch.equal(Xs, full[:, cb[rank]:cb[rank + 1]])
        # partial Gram + all_reduce == full Gram (the POD exchange), small dense check on CPU tensors
        G = Xs @ Xs.T
        dist.all_reduce(G)
        assert torch.allclose(G, full @ full.T)
        back = rd.all_gather_cols(Xs[:3].contiguous(), Dp)
        assert torch.equal(back, full[:3])
        # Gram-free POD on K-sharded rows (uneven shards): both ranks get the SVD of the full centred matrix
        import sys
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from host_engine import HostEngine
        rng = np.random.default_rng(5)
        Kt, Dp, n = 130, 96, 6
        Xf = (rng.standard_normal((Kt, 40)) * 0.7 ** np.arange(40)) @ rng.standard_normal((40, Dp)) + rng.standard_normal(Dp)
        cut = 37
        mine = torch.as_tensor(Xf[:cut].copy() if rank == 0 else Xf[cut:].copy())
        comps, sig = rd.distributed_pca(HostEngine(), mine, n, counts=[cut, Kt - cut], method="krylov")
        Xc = Xf - Xf.mean(axis=0)
        _, s_ref, vt = np.linalg.svd(Xc, full_matrices=False)
        np.testing.assert_allclose(sig.numpy(), s_ref[:n], rtol=1e-9)
        vt = vt[:n] * np.sign(vt[np.arange(n), np.abs(vt[:n]).argmax(axis=1)])[:, None]
        assert np.abs(comps.numpy() - vt).max() < 1e-7
        both = [torch.empty_like(comps) for _ in range(world)]
        dist.all_gather(both, comps)
        assert torch.equal(both[0], both[1])                      # replicated Lanczos: bit-identical on all ranks
        # a rank with no rows takes part in the collectives only
        empty = torch.as_tensor(Xf if rank == 0 else np.empty((0, Dp)))
        comps2, sig2 = rd.distributed_pca(HostEngine(), empty, n, counts=[Kt, 0], method="krylov")
        np.testing.assert_allclose(sig2.numpy(), s_ref[:n], rtol=1e-9)
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def _greedy_worker(rank, world, port, tmp):
    """dist.greedy_build_sharded over 2 ranks (argmax all_gather + owner broadcast per step) == the oracle's greedy on
    the whole training set: same index sequence, same raw basis rows, same parameters, for both criteria."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from host_engine import HostSolutionsManager
        from oracle import FEMOracle
        from oracle.rb import greedy_build
        o = FEMOracle((2, 2), 6)
        K, n = 23, 5                                              # odd K: shards of 11 and 12
        y = 10 ** np.random.default_rng(42).uniform(0, 6, (K, 2, 2))
        U = o.generate_solutions(y)
        h1 = o.H10norm(U)
        sl = rd.local_slice(K)
        sm = HostSolutionsManager(o)
        for crit in ("galerkin", "$H^1_0$"):
            basis, a, picked = rd.greedy_build_sharded(sm, n, U[sl], y[sl], h1[sl], K, greedy_for=crit)
            b_ref, a_ref, p_ref = greedy_build(o, n, U, y, h1, greedy_for=crit)
            assert picked == p_ref, (crit, picked, p_ref)
            assert picked[0] == 0                                   # the exact 1.0-vs-1.0 tie of round 1 goes to index 0
            np.testing.assert_array_equal(basis, b_ref)
            np.testing.assert_array_equal(np.asarray(a), np.asarray(a_ref))
        # a rank that owns nothing (K < world) still takes part in every exchange
        basis, a, picked = rd.greedy_build_sharded(sm, 1, U[:1] if rank == 1 else U[:0], y[:1] if rank == 1 else y[:0],
                                                   h1[:1] if rank == 1 else h1[:0], 1)
        assert picked == [0] and np.array_equal(basis[0], U[0])
        open(os.path.join(tmp, f"greedy_ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_greedy_gloo(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_greedy_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "greedy_ok0").exists() and (tmp_path / "greedy_ok1").exists()


def test_two_rank_collectives_gloo(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
