"""Host logic of the Gram-free POD (pod.krylov_pca) on CPU tensors: control flow, conventions, rank-deficient input.

The dense products go through tests/host_engine.HostEngine (torch matmul); the GPU parity of the same function against
the Gram route and the oracle is in tests/test_zz_krylov_pod_gpu.py."""
import numpy as np
import pytest
import torch

from host_engine import HostEngine
from romhighcontrast_b200.pod import krylov_pca


def _svd_reference(X, n):
    Xc = X - X.mean(axis=0)
    _, s, vt = np.linalg.svd(Xc, full_matrices=False)
    vt = vt[:n]
    sign = np.sign(vt[np.arange(len(vt)), np.abs(vt).argmax(axis=1)])
    return vt * sign[:, None], s[:n]


def _snapshots(K, D, decay, seed):
    rng = np.random.default_rng(seed)
    r = min(K, D)
    U = np.linalg.qr(rng.standard_normal((K, r)))[0]
    V = np.linalg.qr(rng.standard_normal((D, r)))[0]
    s = decay ** np.arange(r)
    return (U * s) @ V.T + 0.3 * rng.standard_normal(D)       # non-zero column mean


@pytest.mark.parametrize("K,D,decay,n", [(300, 200, 0.8, 20), (150, 400, 0.9, 10), (2000, 96, 0.7, 20)])
def test_krylov_pca_matches_svd(K, D, decay, n):
    X = _snapshots(K, D, decay, 0)
    stats = {}
    comps, sig, mean = krylov_pca(HostEngine(), torch.as_tensor(X), n, stats=stats)
    vt, s = _svd_reference(X, n)
    np.testing.assert_allclose(mean.numpy(), X.mean(axis=0), rtol=0, atol=1e-14)
    np.testing.assert_allclose(sig.numpy(), s, rtol=1e-9)
    c = comps.numpy()
    np.testing.assert_allclose(c @ c.T, np.eye(n), atol=1e-12)
    for i in range(n):
        if s[i] / s[0] > 1e-6:
            assert np.abs(c[i] - vt[i]).max() < 1e-7, i
    assert stats["krylov_dim"] <= 960 and stats["steps"] >= 2
    assert torch.equal(torch.as_tensor(X), torch.as_tensor(_snapshots(K, D, decay, 0)))     # input not modified


def test_krylov_pca_rank_deficient_and_small():
    rng = np.random.default_rng(1)
    X = rng.standard_normal((40, 5)) @ rng.standard_normal((5, 64))           # centred rank <= 5 < n
    comps, sig, _ = krylov_pca(HostEngine(), torch.as_tensor(X), 8)
    _, s = _svd_reference(X, 8)
    np.testing.assert_allclose(sig.numpy()[:5], s[:5], rtol=1e-9)
    assert np.all(sig.numpy()[5:] <= 1e-6 * s[0])
    # fewer snapshots than requested components: n is clipped like sklearn's min(K, D) bound
    comps, sig, _ = krylov_pca(HostEngine(), torch.as_tensor(X[:3].copy()), 8)
    assert comps.shape == (3, 64) and sig.shape == (3,)
    # centring in place is honoured
    Xt = torch.as_tensor(X.copy())
    krylov_pca(HostEngine(), Xt, 4, center_in_place=True)
    np.testing.assert_allclose(Xt.numpy().mean(axis=0), 0, atol=1e-13)


def test_krylov_pca_degenerate_arguments():
    X = torch.as_tensor(np.random.default_rng(2).standard_normal((10, 16)))
    comps, sig, mean = krylov_pca(HostEngine(), X, 0)
    assert comps.shape == (0, 16) and sig.shape == (0,) and mean.shape == (16,)
    with pytest.raises(ValueError):
        krylov_pca(HostEngine(), X[:0], 3)
    # identical snapshots: the centred matrix is zero, every singular value is 0 and the call still returns
    same = torch.ones(7, 16, dtype=torch.float64) * 0.25
    comps, sig, _ = krylov_pca(HostEngine(), same, 3)
    assert comps.shape[1] == 16 and float(sig.abs().max()) == 0.0
