#!/usr/bin/env python
"""Golden record of the UNMODIFIED reference driver `experiment()` (TEST INFRASTRUCTURE ONLY).

Run in the build container, where /root/reference exists:
    python oracle/gen_golden_experiment.py [--ref /root/reference] [--out tests/golden]

Runs /root/reference/src/experiments/HighContrast.py::experiment (:118-215) as it is -- the training-set sampler
`get_a2test_and_train` (:99-115), the four builders of its `__main__` block (:33-38, :512), the per-n statistics loop, the
joblib checkpoint -- on a geometry the dense reference handles in seconds, and records everything the GPU classes have
to reproduce when the same driver runs on them: the training set, the measurement points, the keys of the `data`
dictionary (timing entries included), every error curve, the selected snapshots, and what the cached path does on a
second call (`reduced_basis_builder.marker`, :172, is an attribute nobody defines).

Stubs (nothing is written under /root/reference): `pathos` (2-line shim, as in gen_golden.py), matplotlib / seaborn
(MagicMock: plotting is out of scope), `src.config` (its import would mkdir under the reference root; replaced by a
module whose `results_path` points into a temp dir).  `tests/test_gpu5_experiment_flow.py` reads the .npz.
"""
from __future__ import annotations

import argparse
import os
import sys
import tempfile
import types
import warnings
from pathlib import Path
from unittest.mock import MagicMock

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_golden import import_reference   # noqa: E402

CONFIG = dict(mesh_discretization_per_dim=6, diff_coef_refinement=10, vn_max_dim=8, num_measurements=30,
              blocks_geometry=(4, 4), max_num_samples_offline=120, seed=42, num_cores=1, method="lsqsparse", verbose=False)
HIGH_CONTRAST_BLOCKS = [[(0, 1)], [(1, 3)], [(2, 1), (2, 2), (2, 3)]]          # HighContrast.py:512


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
    args = ap.parse_args()
    warnings.filterwarnings("ignore")
    import_reference(args.ref)
    results = Path(tempfile.mkdtemp(prefix="romhc_results_"))
    cfg = types.ModuleType("src.config")
    cfg.results_path = results
    cfg.data_path = results
    sys.modules["src.config"] = cfg
    for m in ("matplotlib", "matplotlib.pylab", "matplotlib.pyplot", "matplotlib.ticker", "matplotlib.cm", "seaborn"):
        sys.modules.setdefault(m, MagicMock())
    import src.experiments.HighContrast as HC
    assert HC.__file__.startswith(args.ref), HC.__file__

    builders = HC.reduced_basis_builders                       # the module's own list (:33-38)
    names = [b.name for b in builders]
    # measurement points are drawn inside experiment() from the legacy global RNG right after the sampler: record them by
    # replaying the same RNG calls (the driver itself stays untouched)
    sm0, a0, ahc0 = HC.get_a2test_and_train(CONFIG["blocks_geometry"], HIGH_CONTRAST_BLOCKS, CONFIG["mesh_discretization_per_dim"],
                                            CONFIG["diff_coef_refinement"], CONFIG["max_num_samples_offline"], CONFIG["seed"],
                                            CONFIG["num_cores"], CONFIG["method"])
    points = np.random.uniform(size=(CONFIG["num_measurements"], 2))

    sm, data, a, ahc = HC.experiment(name="golden", reduced_basis_builders=builders, high_contrast_blocks=HIGH_CONTRAST_BLOCKS,
                                     recalculate=True, recalculate_basis=True, **CONFIG)
    np.testing.assert_array_equal(a, a0)
    # (that `points` are the driver's own measurement points is checked by tests/test_experiment_driver_cpu.py: the restated
    # driver draws them the same way and reproduces every state-estimation error of this record bit for bit)
    out = dict(a=a, a_high_contrast=ahc, points=points, solutions_H1norm=data["solutions_H1norm"],
               solutions=data["solutions"], names=np.array(names), data_keys=np.array(sorted(data.keys())),
               fields=np.array(HC.TypeOfProblems._fields))
    U = data["solutions"]
    for i, name in enumerate(names):
        d = data[name]
        out[f"b{i}_keys"] = np.array(sorted(d.keys()))
        rb = d["basis"]
        out[f"b{i}_basis"] = np.asarray(rb.basis)
        out[f"b{i}_a"] = np.asarray(rb.a)
        out[f"b{i}_idx"] = np.array([int(np.argmin(np.abs(U - b).sum(axis=1))) for b in np.asarray(rb.basis)])
        out[f"b{i}_ns"] = np.array(sorted(d["errors"].keys()))
        assert sorted(d["times"].keys()) == sorted(d["errors"].keys())
        assert type(d["times"][1]).__name__ == "TypeOfProblems" and all(isinstance(t, float) for t in d["times"][1])
        for n in sorted(d["errors"].keys()):
            for f, v in zip(HC.TypeOfProblems._fields, d["errors"][n]):
                out[f"b{i}_n{n}_{f}"] = np.asarray(v)
    # second call: everything is cached -> the driver reaches `reduced_basis_builder.marker` (:172)
    try:
        HC.experiment(name="golden", reduced_basis_builders=builders, high_contrast_blocks=HIGH_CONTRAST_BLOCKS, **CONFIG)
        out["cached_call_exception"] = np.array("")
    except Exception as e:                                      # noqa: BLE001
        out["cached_call_exception"] = np.array(f"{type(e).__name__}: {e}")
    print("cached call:", out["cached_call_exception"])
    os.makedirs(args.out, exist_ok=True)
    np.savez_compressed(os.path.join(args.out, "g9_experiment_4x4_N6.npz"), **out)
    print("wrote g9_experiment_4x4_N6.npz:", {k: np.shape(v) for k, v in list(out.items())[:12]})
    for i, name in enumerate(names):
        print(name, out[f"b{i}_idx"], "max fm err n=8: %.3e" % out[f"b{i}_n8_forward_modeling"].max())


if __name__ == "__main__":
    main()
