"""Test infrastructure only: a torch-on-CPU object with the dense-helper signatures of romhighcontrast_b200.engine.Engine.

It lets the `-m "not gpu"` suite drive the HOST logic that sits on those helpers (Lanczos control flow, shard
bookkeeping, collectives under gloo) without a GPU.  The product never imports it: Engine itself raises without CUDA.
"""
import torch


class HostEngine:
    device = torch.device("cpu")

    def gemm_nt(self, A, B, symmetric=False, splitk=False):
        assert A.is_contiguous() and B.is_contiguous() and A.shape[1] == B.shape[1]
        return A @ B.T

    def gemm_nn(self, A, B):
        assert A.is_contiguous() and B.is_contiguous() and A.shape[1] == B.shape[0]
        return A @ B

    def gemm_tn(self, A, B):
        assert A.is_contiguous() and B.is_contiguous() and A.shape[0] == B.shape[0]
        return A.T @ B

    def tsqr_r(self, W):
        return torch.linalg.qr(W.T.contiguous(), mode="r")[1]

    def row_norms(self, X):
        return torch.linalg.vector_norm(X, dim=1)

    def column_mean(self, X):
        return X.mean(dim=0)

    def center_rows_(self, X, mean):
        X -= mean
        return X


class HostFEMEngine(HostEngine):
    """The FEM operations dist.greedy_build_sharded asks of an Engine, restated on the CPU oracle's sparse matrices
    (compact layout: Dp == D, pad / unpad are the identity).  Test infrastructure for the gloo world-size-2 run."""

    def __init__(self, oracle):
        import numpy as np
        self.o = oracle
        self.np = np
        self.nb = oracle.blocks_geometry[0] * oracle.blocks_geometry[1]
        self.D = self.Dp = oracle.vspace_dim

    def dev(self, a):
        return torch.as_tensor(self.np.ascontiguousarray(self.np.asarray(a, dtype=self.np.float64)))

    def pad(self, compact):
        return self.dev(compact).reshape(-1, self.D)

    def params(self, a):
        return self.dev(a).reshape(-1, self.nb)

    def error_norm(self, U, coef, basis):
        diff = U.numpy() if basis is None else coef.numpy() @ basis.numpy() - U.numpy()
        return torch.as_tensor(self.o.H10norm(diff))

    def project_operators(self, Phi):
        Ahat, bhat = self.o.reduced_operators(Phi.numpy())
        return torch.as_tensor(Ahat.reshape(self.nb, len(Phi), len(Phi))), torch.as_tensor(bhat)

    def reduced_solve(self, y, Ahat, rhs, check=True, return_info=False):
        Ak = self.np.einsum("qij,kq->kij", Ahat.numpy(), y.numpy())
        r = rhs.numpy()
        r = self.np.broadcast_to(r, (len(Ak), r.shape[-1])) if r.ndim == 1 else r
        Cc = torch.as_tensor(self.np.linalg.solve(Ak, r[..., None])[..., 0])
        return (Cc, torch.zeros(len(Ak), dtype=torch.int32)) if return_info else Cc

    def argmax_dev(self, v):
        i = int(self.np.argmax(v.numpy()))
        return torch.tensor([i], dtype=torch.int64), v[i].reshape(1).clone()

    def argmax(self, v):
        i = int(self.np.argmax(v.numpy()))
        return i, float(v[i])

    def l2_norm(self, X):
        return torch.linalg.vector_norm(X, dim=1)

    def unpad(self, X):
        return X


class HostSolutionsManager:
    """What greedy_build_sharded reads from a SolutionsManagerFEM, on top of HostFEMEngine."""

    def __init__(self, oracle):
        self.o = oracle
        self.vspace_dim = oracle.vspace_dim
        self.blocks_geometry = oracle.blocks_geometry
        self._eng = HostFEMEngine(oracle)

    def _engine_(self):
        return self._eng

    def _projection_coefficients_dev(self, eng, U, Phi):
        return torch.as_tensor(self.o.projection_coefficients(U.numpy(), Phi.numpy()))
