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


def test_coarsening_chain_rule():
    """the hierarchy rule shared with Context::build_levels: nested halving; when it gets stuck on an odd count whose
    grid the dense coarsest solve cannot take, one non-nested transfer to a power of two from the deepest batched level"""
    from gmg_twin import coarsening_chain
    assert coarsening_chain(4, 4, 64) == ([64, 32, 16, 8, 4, 2, 1], None)          # BASELINE configs[2]: untouched
    assert coarsening_chain(2, 2, 32) == ([32, 16, 8, 4, 2, 1], None)
    assert coarsening_chain(3, 3, 43) == ([43, 16, 8, 4, 2, 1], 0)                 # BASELINE configs[1]
    assert coarsening_chain(3, 3, 44) == ([44, 22, 8, 4, 2, 1], 1)                 # deepest batched level bridges
    assert coarsening_chain(4, 4, 63) == ([63, 32, 16, 8, 4, 2, 1], 0)
    assert coarsening_chain(2, 2, 10) == ([10, 5], None)                           # 9 x 9 coarsest grid: dense solve
    assert coarsening_chain(2, 2, 27) == ([27], None)                              # whole hierarchy inside the tail kernel
    assert coarsening_chain(3, 3, 43, bridge=False) == ([43], None)
    for n in range(3, 200):
        chain, j = coarsening_chain(3, 3, n)
        if j is not None:
            ratio = chain[j] / chain[j + 1]
            assert 1.5 <= ratio < 3.0 and chain[-1] == 1, (n, chain)


def test_bridge_matrix_is_partition_of_unity_inside_blocks():
    from gmg_twin import bridge_matrix
    P = bridge_matrix(3, 43, 16)
    assert P.shape == (130, 49)
    np.testing.assert_allclose(P.sum(axis=1), 1.0, rtol=0, atol=1e-15)
    for b in range(4):                                   # subdomain edges are vertices of both meshes
        row = P[b * 43]
        assert row[b * 16] == 1.0 and np.count_nonzero(row) == 1
    for i in range(130):                                 # no interpolation across a subdomain edge
        cols = np.nonzero(P[i])[0]
        blk = min(i // 43, 2)
        assert cols.min() >= blk * 16 and cols.max() <= (blk + 1) * 16


def test_twin_with_bridge_matches_oracle_and_converges_fast():
    """(3,3), N = 43 is BASELINE configs[1] (129 x 129 cells, prime cells per subdomain)"""
    geo, N = (3, 3), 43
    a = 10 ** np.random.default_rng(5).uniform(0, 6, geo)
    u, it = pcg(a, N)
    o = FEMOracle(geo, N)
    assert relerr(u, o.generate_solutions(a[None])[0]) < 1e-10
    assert it <= 14                                      # single-level smoothing needs ~170 iterations here
