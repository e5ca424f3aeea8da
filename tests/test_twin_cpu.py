"""CPU check that the numpy twin of the GPU algorithm agrees with the oracle (algorithm validation)."""
import numpy as np
import pytest

from conftest import golden, relerr
from gmg_twin import pcg
from oracle import FEMOracle


@pytest.mark.parametrize("geo,N,cmax", [((2, 2), 8, 1e6), ((3, 2), 4, 1e6), ((4, 4), 8, 1e6), ((2, 3), 6, 1e4),
                                        ((3, 3), 5, 1e6), ((1, 3), 8, 1e3)])
def test_twin_matches_oracle(geo, N, cmax):
    rng = np.random.default_rng(5)
    a = 10 ** rng.uniform(0, np.log10(cmax), geo)
    o = FEMOracle(geo, N)
    u, it = pcg(a, N)
    assert relerr(u, o.generate_solutions(a[None])[0]) < 1e-10
    assert it < 400


def test_twin_beats_reference_on_floating_inclusion():
    g = golden("g8_floating_4x4_N8.npz")
    u, it = pcg(g["a"][0], 8)
    assert relerr(u, g["U_truth"]) < 1e-11        # reference's own solvers: ~1e-5
    assert it < 40
