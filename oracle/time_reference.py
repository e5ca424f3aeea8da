#!/usr/bin/env python
"""Time the UNMODIFIED reference on BASELINE configs[0] (TEST / MEASUREMENT INFRASTRUCTURE ONLY).

    python oracle/time_reference.py [--ref /root/reference] [--out tests/golden/reference_timing_config0.json]

configs[0]: (2,2) subdomains, N = 32 (64 x 64 cells, D = 3969), 100 random-contrast snapshots (10^U(0,6)),
greedy / POD n = 10 + Galerkin -- the reference's own CPU-runnable case, with the `experiment()` defaults
method="lsqsparse", num_cores=1 (src/experiments/HighContrast.py:125,496).  The Python reference cannot travel to
the GPU box, so this runs where /root/reference exists (the build container) and the result is committed; bench.py
carries it as `cpu_baseline.true_reference_config0` next to the GPU figures for the same workload
(`secondary.config0`).  The host it ran on is recorded in the file.
"""
from __future__ import annotations

import argparse
import json
import os
import platform
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_golden import import_reference  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                                                  "reference_timing_config0.json"))
    args = ap.parse_args()
    SM, RB, ES = import_reference(args.ref)
    geo, N, K, n = (2, 2), 32, 100, 10
    y = 10 ** np.random.default_rng(42).uniform(0, 6, (K,) + geo)
    T = {}

    def tm(name, f, *a, **kw):
        t0 = time.perf_counter()
        r = f(*a, **kw)
        T[name] = time.perf_counter() - t0
        return r

    sm = tm("assembly_s", SM.SolutionsManagerFEM, geo, N, num_cores=1, method="lsqsparse")
    U = tm("snapshots_s", sm.generate_solutions, y)
    h1 = tm("h10norm_s", sm.H10norm, U)
    rbg = tm("greedy_galerkin_s", RB.ReducedBasisGreedy(greedy_for=RB.GREEDY_FOR_GALERKIN).build, n=n, sm=sm, solutions2train=U,
             a2train=y, solutions2train_h1norm=h1)
    rbh = tm("greedy_h10_s", RB.ReducedBasisGreedy(greedy_for=RB.GREEDY_FOR_H10).build, n=n, sm=sm, solutions2train=U, a2train=y,
             solutions2train_h1norm=h1)
    rbp = tm("pca_s", RB.ReducedBasisPCA().build, n=n, sm=sm, solutions2train=U, a2train=y)
    rbg.orthonormalize()
    fm = tm("forward_modeling_s", rbg.forward_modeling, sm, y)
    pj = tm("projection_s", rbg.projection, sm, U)
    cpu = ""
    try:
        cpu = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:
        pass
    out = {
        "config": "configs[0]: (2,2) subdomains, N=32 (64x64 cells, D=3969), 100 snapshots, contrast 10^U(0,6) seed 42, n=10",
        "reference": "unmodified /root/reference/src/lib, method='lsqsparse', num_cores=1 (pathos shim only)",
        "seconds": T,
        "snapshot_solves_per_s": K / T["snapshots_s"],
        "reduced_galerkin_solves_per_s": K / T["forward_modeling_s"],
        "host": {"cpu": cpu, "logical_cores": os.cpu_count(), "python": platform.python_version(),
                 "numpy": np.__version__, "threads": os.environ.get("OMP_NUM_THREADS", "default")},
        "where": "build container (the reference is Python and does not travel to the GPU box)",
        "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
        "checks": {"greedy_galerkin_indices": None, "fm_rel_err_max": float(np.max(sm.H10norm(fm - U) / h1))},
    }
    json.dump(out, open(args.out, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
