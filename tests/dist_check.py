#!/usr/bin/env python
"""Multi-GPU check of the two communicating stages (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py

Every rank builds the same seeded training set, keeps only its contiguous shard, and the sharded greedy / POD must
reproduce the single-GPU results computed redundantly on each rank (indices identical, singular values to 1e-9,
components to 1e-7)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from romhighcontrast_b200 import dist as rd
    from romhighcontrast_b200.lib.ReducedBasis import ReducedBasisGreedy, GREEDY_FOR_GALERKIN, GREEDY_FOR_H10
    from romhighcontrast_b200.lib.SolutionsManagers import SolutionsManagerFEM
    from romhighcontrast_b200.pod import pca_components
    # default: a small ragged case; DIST_CHECK_CASE=config4 runs the geometry of BASELINE configs[4] ((8,8) subdomains,
    # 512 x 512 cells) at a reduced training-set size
    geo, N, K, n = ((8, 8), 64, 1024, 10) if os.environ.get("DIST_CHECK_CASE") == "config4" else ((3, 3), 16, 203, 8)
    y = 10 ** np.random.default_rng(5).uniform(0, 6, (K,) + geo)
    sm = SolutionsManagerFEM(geo, N)
    eng = sm._engine_()
    sl = rd.local_slice(K)
    U_loc = sm.generate_solutions(y[sl])                       # sharded snapshot solves: no collective
    # gather to compare against the single-GPU path (test only)
    parts = [None] * rd.world()
    dist.all_gather_object(parts, U_loc)
    U = np.vstack(parts)
    h1 = sm.H10norm(U)
    ok = True
    for crit in (GREEDY_FOR_GALERKIN, GREEDY_FOR_H10):
        ref = ReducedBasisGreedy(greedy_for=crit).build(n=n, sm=sm, solutions2train=U, a2train=y, solutions2train_h1norm=h1)
        basis, a, picked = rd.greedy_build_sharded(sm, n, U[sl], y[sl], h1[sl], K, greedy_for=crit)
        same = picked == ref.selected_indices and np.array_equal(basis, ref.basis)
        ok &= same
        if rd.rank() == 0:
            print(f"greedy {crit}: sharded {picked} single {ref.selected_indices} -> {'OK' if same else 'MISMATCH'}")
    Xp = eng.pad(U)
    comps_ref, sig_ref, _ = pca_components(eng, Xp, n)
    counts = [s.stop - s.start for s in (rd.local_slice(K, r, rd.world()) for r in range(rd.world()))]
    timings = {}
    comps, sig = rd.distributed_pca(eng, eng.pad(U[sl]), n, counts=counts, timings=timings)
    e_s = float(((sig - sig_ref).abs() / sig_ref).max())
    e_c = float((comps - comps_ref).abs().max())
    good = e_s < 1e-9 and e_c < 1e-7
    ok &= good
    if rd.rank() == 0:
        print(f"distributed POD: singular values rel err {e_s:.2e}, components max abs err {e_c:.2e} -> {'OK' if good else 'MISMATCH'}; timings(ms) {timings}")
    # Gram-free route on the K-sharded rows (one (b, Dp) all_reduce per block-Lanczos step); every rank must also hold
    # bit-identical components (replicated Lanczos on all-reduced data)
    kst = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    comps_k, sig_k = rd.distributed_pca(eng, eng.pad(U[sl]), n, counts=counts, timings=kst, method="krylov")
    e1.record(); torch.cuda.synchronize()
    e_s = float(((sig_k - sig_ref).abs() / sig_ref).max())
    e_c = float((comps_k - comps_ref).abs().max())
    ref0 = comps_k.clone()
    dist.broadcast(ref0, src=0)
    identical = bool(torch.equal(ref0, comps_k))
    good = e_s < 1e-9 and e_c < 1e-7 and identical
    ok &= good
    if rd.rank() == 0:
        print(f"distributed POD (krylov): singular values rel err {e_s:.2e}, components max abs err {e_c:.2e}, "
              f"bit-identical across ranks {identical} -> {'OK' if good else 'MISMATCH'}; {e0.elapsed_time(e1):.1f} ms incl. "
              f"first-call setup, {kst}")
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rd.rank() == 0:
        print("DIST_CHECK", "PASS" if int(flag.item()) else "FAIL")
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
