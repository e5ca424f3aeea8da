"""CPU oracle for the ROMHighContrast hot path -- TEST INFRASTRUCTURE ONLY.

A sparse (scipy) restatement of the reference's `src/lib` numerics, used as the
checker for the CUDA path.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this package;
the product (`romhighcontrast_b200`) never does.

Parity status: PINNED.  `oracle/gen_golden.py` imports the unmodified reference
from /root/reference (with a 2-line `pathos` shim) in the build container, runs
it on seeded inputs and stores input/output vectors under `tests/golden/`;
`tests/test_oracle_golden.py` checks this restatement against those vectors.
"""
from .fem import (FEMOracle, stiffness_csr, load_vector, block_stiffness_csr)  # noqa: F401
from .rb import (greedy_build, pca_build, random_build, state_estimation,  # noqa: F401
                 estimator_inv, estimator_linear, sort_orthogonalize_base,
                 high_contrast_coefficient, pbdw_correction, INFINIT_A)
