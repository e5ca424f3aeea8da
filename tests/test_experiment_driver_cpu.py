"""The restated experiment driver (tests/experiment_driver.py) on the UNMODIFIED reference classes reproduces, bit for bit,
what the reference's own driver recorded (tests/golden/g9_experiment_4x4_N6.npz, oracle/gen_golden_experiment.py).

This pins the test infrastructure the GPU flow test stands on.  It needs /root/reference (build container only) and runs in
a subprocess because importing the reference re-points `src` / `lib` on sys.path.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

SCRIPT = r"""
import sys, tempfile, types, warnings
import numpy as np
warnings.filterwarnings("ignore")
sys.path.insert(0, ROOT + "/oracle"); sys.path.insert(0, ROOT + "/tests")
from gen_golden import import_reference
from gen_golden_experiment import CONFIG, HIGH_CONTRAST_BLOCKS
SM, RB, ES = import_reference(REF)
sys.path.insert(0, ROOT + "/tests")
import experiment_driver as drv
ns = types.SimpleNamespace(SolutionsManagerFEM=SM.SolutionsManagerFEM, ReducedBasisGreedy=RB.ReducedBasisGreedy,
                           ReducedBasisRandom=RB.ReducedBasisRandom, INFINIT_A=RB.INFINIT_A,
                           GREEDY_FOR_H10=RB.GREEDY_FOR_H10, GREEDY_FOR_GALERKIN=RB.GREEDY_FOR_GALERKIN)
g = np.load(ROOT + "/tests/golden/g9_experiment_4x4_N6.npz", allow_pickle=False)
path = tempfile.mkdtemp(prefix="romhc_drv_")
kw = dict(N=CONFIG["mesh_discretization_per_dim"], refinement=CONFIG["diff_coef_refinement"], vn_max_dim=CONFIG["vn_max_dim"],
          num_measurements=CONFIG["num_measurements"], blocks_geometry=CONFIG["blocks_geometry"],
          high_contrast_blocks=HIGH_CONTRAST_BLOCKS, max_samples=CONFIG["max_num_samples_offline"], seed=CONFIG["seed"],
          method=CONFIG["method"])
builders = drv.default_builders(ns)
sm, data, a, a_hc, points = drv.experiment(ns, path, builders, recalculate=True, recalculate_basis=True, **kw)
assert [b.name for b in builders] == list(g["names"])
assert sorted(data.keys()) == list(g["data_keys"]), sorted(data.keys())
for k in ("a", "a_high_contrast", "points", "solutions", "solutions_H1norm"):
    np.testing.assert_array_equal({"a": a, "a_high_contrast": a_hc, "points": points}.get(k, data.get(k)), g[k], err_msg=k)
for i, name in enumerate(g["names"]):
    d = data[str(name)]
    assert sorted(d.keys()) == list(g[f"b{i}_keys"])
    np.testing.assert_array_equal(np.asarray(d["basis"].basis), g[f"b{i}_basis"])
    np.testing.assert_array_equal(np.asarray(d["basis"].a), g[f"b{i}_a"])
    assert sorted(d["errors"].keys()) == list(g[f"b{i}_ns"]) == sorted(d["times"].keys())
    for n in g[f"b{i}_ns"]:
        for f, v in zip(g["fields"], d["errors"][n]):
            np.testing.assert_array_equal(np.asarray(v), g[f"b{i}_n{n}_{f}"], err_msg=f"{name} n={n} {f}")
        assert all(isinstance(t, float) for t in d["times"][n])
try:
    drv.experiment(ns, path, builders, **kw)
    msg = ""
except Exception as e:
    msg = f"{type(e).__name__}: {e}"
assert msg == str(g["cached_call_exception"]), msg
print("DRIVER-OK")
"""


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "lib")), reason="needs /root/reference (build container)")
def test_restated_driver_reproduces_the_reference_driver_record():
    code = f"ROOT = {ROOT!r}\nREF = {REF!r}\n" + SCRIPT
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env, cwd="/tmp")
    assert r.returncode == 0 and "DRIVER-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
