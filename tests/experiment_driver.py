"""Test infrastructure: the reference's experiment driver, restated so it can run where /root/reference does not exist.

Follows /root/reference/src/experiments/HighContrast.py: `get_full_a` :58-64, `calculate_time` :67-78, `get_data` :93-96,
`get_a2test_and_train` :99-115 and `experiment` :118-215 -- same call sequence on the manager / builder classes it is
handed, same legacy-RNG draws, same `data` dictionary (keys, timing entries, `TypeOfProblems` tuples), same joblib
checkpoint file, same behaviour on the cached path (`:172` reads `reduced_basis_builder.marker`, which no class defines).
The classes come in through `ns` (anything with SolutionsManagerFEM, ReducedBasisGreedy, ReducedBasisRandom, INFINIT_A,
GREEDY_FOR_H10, GREEDY_FOR_GALERKIN), so the same driver runs on the unmodified reference (CPU, build container:
`tests/test_experiment_driver_cpu.py` checks it reproduces the golden record bit for bit) and on the GPU mirror
(`tests/test_gpu5_experiment_flow.py`).  The golden record itself comes from the reference's OWN driver
(`oracle/gen_golden_experiment.py`).
"""
import os
from collections import namedtuple
from time import time

import joblib
import numpy as np

TypeOfProblems = namedtuple("TypeOfProblems", "forward_modeling projection state_estimation parameter_estimation_inverse "
                                              "parameter_estimation_linear")


def timed(func):
    """(seconds, result) of a keyword-only call (HighContrast.py:67-78 without the prints)."""
    def call(**kwargs):
        t0 = time()
        res = func(**kwargs)
        return time() - t0, res
    return call


def default_builders(ns):
    """The module-level builder list of the driver (HighContrast.py:33-38)."""
    return [ns.ReducedBasisRandom(), ns.ReducedBasisRandom(False), ns.ReducedBasisGreedy(greedy_for=ns.GREEDY_FOR_H10),
            ns.ReducedBasisGreedy(greedy_for=ns.GREEDY_FOR_GALERKIN)]


def training_set(ns, blocks_geometry, high_contrast_blocks, N, refinement, max_samples, seed, num_cores=1, method="lsq"):
    """Manager + training parameters: the {1e10, 1}^d corners first, then a (sub-sampled) 1 / linspace grid (:99-115)."""
    sm = ns.SolutionsManagerFEM(blocks_geometry, N=N, num_cores=num_cores, method=method)
    d = len(high_contrast_blocks)
    per_dim = min(refinement * int(np.log2(ns.INFINIT_A)), int(np.ceil(max_samples ** (1 / d))))
    axis = 1 / np.linspace(1 / ns.INFINIT_A, 1, num=per_dim, endpoint=False)
    grid = np.transpose([g.ravel() for g in np.meshgrid(*[axis] * d)])
    np.random.seed(seed)
    corners = np.transpose([g.ravel() for g in np.meshgrid(*[[ns.INFINIT_A, 1]] * d)])
    if len(grid) > max_samples - len(corners):
        grid = grid[np.random.choice(len(grid), size=max(0, max_samples - len(corners)), replace=False)]
    a_hc = np.vstack((corners, grid))
    a = np.ones((len(a_hc),) + tuple(sm.blocks_geometry))
    for column, blocks in zip(a_hc.T, high_contrast_blocks):
        for p, q in blocks:
            a[:, p, q] = column
    return sm, a, a_hc


def experiment(ns, experiment_path, builders, N=6, refinement=30, vn_max_dim=20, num_measurements=50, blocks_geometry=(4, 4),
               high_contrast_blocks=(((1, 1), (1, 2), (2, 1), (2, 2)),), vn_max_dim2do_stats=None, recalculate=False,
               num_cores=1, max_samples=10000, seed=42, recalculate_basis=False, method="lsqsparse"):
    """HighContrast.py:118-215 with `get_folder_from_params(name)` replaced by an explicit directory."""
    stats_up_to = vn_max_dim if vn_max_dim2do_stats is None else vn_max_dim2do_stats
    os.makedirs(experiment_path, exist_ok=True)
    data_path = os.path.join(str(experiment_path), "data.compressed")
    data = joblib.load(data_path) if os.path.exists(data_path) else dict()

    sm, a, a_hc = training_set(ns, blocks_geometry, high_contrast_blocks, N, refinement, max_samples, seed, num_cores, method)
    if recalculate or "solutions" not in data:
        data["time2calculate_solutions"], data["solutions"] = timed(sm.generate_solutions)(a2try=a)
        data["time2calculate_h1norm"], data["solutions_H1norm"] = timed(sm.H10norm)(solutions=data["solutions"])
        joblib.dump(data, data_path)
    points = np.random.uniform(size=(num_measurements, 2))
    measurements = sm.evaluate_solutions(points, data["solutions"])

    for builder in builders:
        entry = data.get(builder.name)
        if entry is None or entry["basis"].dim < vn_max_dim or recalculate_basis:
            data[builder.name] = {"errors": {}, "times": {}}
            data[builder.name]["time2build"], data[builder.name]["basis"] = timed(builder.build)(
                n=vn_max_dim, sm=sm, solutions2train=data["solutions"], a2train=a, optim_method="lsq",
                solutions2train_h1norm=data["solutions_H1norm"])
            joblib.dump(data, data_path)
        else:
            data[builder.name]["basis"].marker = builder.marker       # :172 -- AttributeError in the reference, kept

    U, h1 = data["solutions"], data["solutions_H1norm"]
    for n in np.arange(1, vn_max_dim + 1):
        for builder in builders:
            entry = data[builder.name]
            if n > stats_up_to or not (recalculate or n not in entry["errors"]):
                continue
            rb = entry["basis"][:n]
            t_se, (c, u_se) = timed(rb.state_estimation)(sm=sm, measurement_points=points, measurements=measurements,
                                                        return_coefs=True)
            t_inv, _ = timed(rb.parameter_estimation_inverse)(c=c)
            t_lin, _ = timed(rb.parameter_estimation_linear)(c=c)
            rb.orthonormalize()
            t_fm, u_fm = timed(rb.forward_modeling)(sm=sm, a=a)
            t_pj, u_pj = timed(rb.projection)(sm=sm, true_solutions=U)
            e_fm = timed(sm.H10norm)(solutions=u_fm - U)[1]
            e_pj = timed(sm.H10norm)(solutions=u_pj - U)[1]
            e_se = timed(sm.H10norm)(solutions=u_se - U)[1]
            entry["errors"][n] = TypeOfProblems(
                forward_modeling=e_fm / h1, projection=e_pj / h1, state_estimation=e_se / h1,
                parameter_estimation_inverse=np.abs(1 - np.array(rb.parameter_estimation_inverse(c)) / a),
                parameter_estimation_linear=np.abs(1 - np.array(rb.parameter_estimation_linear(c)) / a))
            entry["times"][n] = TypeOfProblems(forward_modeling=t_fm, projection=t_pj, state_estimation=t_se,
                                               parameter_estimation_inverse=t_inv, parameter_estimation_linear=t_lin)
            joblib.dump(data, data_path)
    return sm, data, a, a_hc, points
