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
        assert A.is_contiguous() and B.is_contiguous() and A.shape[1] == B.shape[0] and A.shape[1] <= 2048
        return A @ B

    def gemm_tn(self, A, B):
        assert A.is_contiguous() and B.is_contiguous() and A.shape[0] == B.shape[0] and A.shape[1] <= 32
        return A.T @ B

    def column_mean(self, X):
        return X.mean(dim=0)

    def center_rows_(self, X, mean):
        X -= mean
        return X
