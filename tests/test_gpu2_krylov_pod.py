"""Gram-free POD (pod.krylov_pca) and the kernels it adds, on the GPU through the C ABI.

Named to sort last: this path is newer than the rest of the suite and a failure here must not hide the others under -x.
Checks: split-K form of the fp64 DMMA product against a plain fp64 matmul (and its run-to-run determinism), the
reconstruction product with a contraction longer than 768 (opt-in shared-memory window), and the Krylov PCA against
both the Gram route (pca_components) and the oracle's PCA(svd_solver="full") on real snapshots.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_mod():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def make_engine(geo, N):
    from romhighcontrast_b200.engine import Engine
    return Engine(geo, N)


@pytest.mark.parametrize("M,N,Kd", [(192, 192, 65792), (960, 32, 66048), (32, 32, 263168), (20, 7, 4104), (130, 70, 9001 * 2),
                                    (1, 1, 8192), (300, 40, 5000)])
def test_gemm_nt_splitk(torch_mod, M, N, Kd):
    """tolerance: fp64 products of O(1) entries summed over Kd terms in a different order -> 1e-13 relative"""
    torch = torch_mod
    eng = make_engine((2, 2), 4)
    gen = torch.Generator(device="cpu").manual_seed(M * 7 + N)
    A = torch.randn(M, Kd, dtype=torch.float64, generator=gen).cuda()
    B = torch.randn(N, Kd, dtype=torch.float64, generator=gen).cuda()
    ref = A @ B.T
    out = eng.gemm_nt(A, B, splitk=True)
    assert float((out - ref).abs().max() / ref.abs().max()) < 1e-13
    assert torch.equal(out, eng.gemm_nt(A, B, splitk=True))                  # fixed-order reduction
    plain = eng.gemm_nt(A, B)
    assert float((out - plain).abs().max() / ref.abs().max()) < 1e-13
    # operands that are not 16-byte aligned take the plain kernel and stay correct
    A1 = torch.randn(M * Kd + 1, dtype=torch.float64, generator=gen).cuda()[1:].view(M, Kd)      # contiguous, 8-byte offset
    out1 = eng.gemm_nt(A1, B, splitk=True)
    ref1 = A1 @ B.T
    assert float((out1 - ref1).abs().max() / ref1.abs().max()) < 1e-13


@pytest.mark.parametrize("M,Kd,N", [(20, 960, 4104), (32, 769, 1000), (9, 2048, 333)])
def test_gemm_nn_long_contraction(torch_mod, M, Kd, N):
    torch = torch_mod
    eng = make_engine((2, 2), 4)
    gen = torch.Generator(device="cpu").manual_seed(Kd)
    A = torch.randn(M, Kd, dtype=torch.float64, generator=gen).cuda()
    B = torch.randn(Kd, N, dtype=torch.float64, generator=gen).cuda()
    ref = A @ B
    assert float((eng.gemm_nn(A, B) - ref).abs().max() / ref.abs().max()) < 1e-13


@pytest.mark.parametrize("Kd,M,N", [(700, 12, 1032), (10000, 32, 4104), (513, 20, 250), (1500, 7, 1088), (16, 1, 2),
                                    (3000, 25, 66048), (5, 32, 512), (2049, 9, 258)])
def test_gemm_tn_tensor_core_kernel(torch_mod, Kd, M, N):
    """A^T B with the contraction over rows: DMMA kernel (default) against fp64 matmul and against the plain-FMA kernel."""
    torch = torch_mod
    eng = make_engine((2, 2), 4)
    gen = torch.Generator(device="cpu").manual_seed(Kd + M)
    A = torch.randn(Kd, M, dtype=torch.float64, generator=gen).cuda()
    B = torch.randn(Kd, N, dtype=torch.float64, generator=gen).cuda()
    ref = A.T @ B
    scale = float(ref.abs().max())
    out = eng.gemm_tn(A, B)
    assert float((out - ref).abs().max()) / scale < 1e-13
    assert torch.equal(out, eng.gemm_tn(A, B))
    eng.set_option("tn_variant", 0)
    try:
        old = eng.gemm_tn(A, B)
    finally:
        eng.set_option("tn_variant", 1)
    assert float((out - old).abs().max()) / scale < 1e-13
    # rows of B that are not 16-byte aligned fall back to the plain kernel
    B1 = torch.randn(Kd * N + 1, dtype=torch.float64, generator=gen).cuda()[1:].view(Kd, N)
    assert float((eng.gemm_tn(A, B1) - A.T @ B1).abs().max()) / scale < 1e-13


def test_krylov_pca_matches_gram_route_and_oracle(torch_mod):
    """(4,4) blocks, N = 8 (D = 961), K = 1500 snapshots at contrast up to 1e6, n = 20.
    Bars: singular values <= 1e-9 relative, sign-fixed components <= 1e-7 for sigma_i / sigma_1 > 1e-6."""
    torch = torch_mod
    from oracle.rb import pca_components as oracle_pca
    from romhighcontrast_b200.pod import krylov_pca, pca_components
    eng = make_engine((4, 4), 8)
    y = 10 ** np.random.default_rng(42).uniform(0, 6, (1500, 4, 4))
    X, _, _ = eng.solve(eng.params(y))
    n = 20
    stats = {}
    comps, sig, mean = krylov_pca(eng, X, n, stats=stats)
    comps_g, sig_g, mean_g = pca_components(eng, X, n)
    assert float((mean - mean_g).abs().max()) < 1e-15
    np.testing.assert_allclose(sig.cpu().numpy(), sig_g.cpu().numpy(), rtol=1e-9)
    assert float((comps - comps_g).abs().max()) < 1e-7
    U = eng.unpad(X).cpu().numpy()
    co, so, _ = oracle_pca(U, n)
    np.testing.assert_allclose(sig.cpu().numpy(), so, rtol=1e-9)
    c = eng.unpad(comps).cpu().numpy()
    for i in range(n):
        if so[i] / so[0] > 1e-6:
            assert np.abs(c[i] - co[i]).max() < 1e-7, i
    np.testing.assert_allclose(c @ c.T, np.eye(n), atol=1e-12)
    assert stats["krylov_dim"] <= 960 and stats["residual"] <= 1e-10 * float(sig[0]) ** 2
    # the padded slots of every component stay exactly zero (the layout's Dirichlet convention)
    assert float((eng.pad(eng.unpad(comps)) - comps).abs().max()) == 0.0
    # a K-split of the same rows through the K_total argument (what each rank of a sharded run sees, minus the
    # all_reduce) is covered on CPU by tests/test_dist_cpu.py; here: K < n and in-place centring
    c3, s3, _ = krylov_pca(eng, X[:3].contiguous(), 8)
    assert c3.shape == (3, eng.Dp) and float(s3[2]) <= 1e-6 * float(s3[0])       # 3 centred rows have rank 2
    Xc = X.clone()
    krylov_pca(eng, Xc, 4, center_in_place=True)
    assert float(eng.column_mean(Xc).abs().max()) < 1e-15


def test_reduced_basis_pca_krylov_option(torch_mod):
    """ReducedBasisPCA(...).build(pod_method="krylov") gives the basis of the default (Gram) build."""
    from lib.ReducedBasis import ReducedBasisPCA
    from lib.SolutionsManagers import SolutionsManagerFEM
    sm = SolutionsManagerFEM((2, 2), N=10, num_cores=1, method="lsq")
    y = 10 ** np.random.default_rng(42).uniform(0, 6, (100, 2, 2))
    U = sm.generate_solutions(a2try=y)
    ref = ReducedBasisPCA().build(n=10, sm=sm, solutions2train=U, a2train=y)
    kry = ReducedBasisPCA().build(n=10, sm=sm, solutions2train=U, a2train=y, pod_method="krylov")
    np.testing.assert_allclose(kry.singular_values_, ref.singular_values_, rtol=1e-9)
    assert np.abs(np.asarray(kry.basis) - np.asarray(ref.basis)).max() < 1e-7
    # "auto" switches on the number of snapshots
    small = ReducedBasisPCA()
    small.GRAM_MAX_SNAPSHOTS = 50
    auto = small.build(n=10, sm=sm, solutions2train=U, a2train=y)
    np.testing.assert_allclose(np.asarray(auto.basis), np.asarray(kry.basis), rtol=0, atol=1e-10)   # same route, same input
    with pytest.raises(ValueError):
        ReducedBasisPCA().build(n=10, sm=sm, solutions2train=U, a2train=y, pod_method="svd")


@pytest.mark.parametrize("b,Dp,decades", [(32, 65792, 0), (22, 262656, 0), (32, 66048, 12), (5, 100, 0), (32, 63, 3), (1, 7, 0),
                                          (32, 400, 8), (17, 4104, 14)])
def test_tsqr_r_matches_lapack(b, Dp, decades):
    """romhc_tsqr_r (Householder TSQR, csrc/dense.cu): R of the QR factorisation of W^T for a row block W (b, Dp), against
    numpy's Householder QR up to the signs of the rows; rows spanning `decades` decades of magnitude (the regime of the
    block Lanczos after a few steps, where a b x b Gram matrix loses everything below sqrt(eps))."""
    import torch
    from romhighcontrast_b200.engine import Engine
    eng = Engine((2, 2), 4)
    rng = np.random.default_rng(b * 1000 + Dp % 977)
    W = rng.standard_normal((b, Dp)) * (10.0 ** -np.linspace(0, decades, b))[:, None]
    R = eng.tsqr_r(torch.as_tensor(W, device="cuda")).cpu().numpy()
    assert R.shape == (b, b) and np.allclose(R, np.triu(R))
    Rn = np.linalg.qr(W.T, mode="r")
    k = min(b, Dp)
    # well-conditioned block (rows scaled, not dependent): entries agree with LAPACK's Householder R up to row signs,
    # relative to the magnitude of the column (column j carries the scale of row j of W)
    scale = np.abs(Rn).max(axis=0)
    assert np.max(np.abs(np.abs(R[:k]) - np.abs(Rn[:k])) / scale) < 1e-11
    # with a nearly dependent row the trailing rows of R are ill conditioned entry by entry, but the factorisation is still
    # backward stable: W W^T = R^T R relative to the row norms, singular values to eps * sigma_max
    if b > 3:
        W[3] = W[1] * 0.5 + 1e-9 * W[3]
        R = eng.tsqr_r(torch.as_tensor(W, device="cuda")).cpu().numpy()
        Rn = np.linalg.qr(W.T, mode="r")
    nr = np.linalg.norm(W, axis=1)
    G = W @ W.T
    assert np.max(np.abs(R.T @ R - G) / np.outer(nr, nr)) < 1e-13
    sv, svn = np.linalg.svd(R, compute_uv=False), np.linalg.svd(Rn, compute_uv=False)
    np.testing.assert_allclose(sv[:k], svn[:k], rtol=1e-9, atol=1e-14 * svn[0])
