"""Device engine: one `Engine` per (blocks_geometry, N, device).

PyTorch is used only as plumbing (device memory, streams, torch.distributed); every numerical
operation on the hot path is a hand-written sm_100a kernel reached through the C ABI in
include/romhc.h.  No CPU fallback: constructing an Engine without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

F64 = torch.float64


def _ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "device tensors passed to libromhc must be contiguous CUDA tensors"
    return C.c_void_p(t.data_ptr())


class PinnedPool:
    """Pinned host blocks handed out as numpy arrays (results of the reference-shaped API: numpy in, numpy out).

    A fresh pageable destination costs a page fault per 4 KB on first touch plus a bounce-buffer memcpy (round 1: the
    class API ran 27 % behind the pinned C-ABI path); a block from this pool is the D2H destination itself.  A block
    returns to the pool when the array AND every view of it have been garbage collected (weakref finaliser on the buffer
    owner), so arrays the caller keeps stay valid for as long as they are referenced.  Free blocks beyond `max_free_bytes`
    are released; if pinning fails the caller falls back to a pageable array."""
    max_free_bytes = 24 << 30

    def __init__(self):
        self.free = []                           # (capacity, address)
        self.lib = None

    def take(self, shape):
        import weakref
        n = int(np.prod(shape))
        nbytes = n * 8
        if self.lib is None:
            self.lib = _lib.load()
        pick = None
        for i, (cap, addr) in enumerate(self.free):
            if nbytes <= cap <= 2 * nbytes + (1 << 20) and (pick is None or cap < self.free[pick][0]):
                pick = i
        if pick is not None:
            cap, addr = self.free.pop(pick)
        else:
            ptr = C.c_void_p()
            cap = (nbytes + (1 << 21) - 1) & ~((1 << 21) - 1)
            if self.lib.romhc_malloc_host(C.byref(ptr), cap) != 0 or not ptr.value:
                return None
            addr = ptr.value
        owner = (C.c_double * n).from_address(addr)
        weakref.finalize(owner, self._give_back, cap, addr)
        return np.ctypeslib.as_array(owner).reshape(shape)

    def _give_back(self, cap, addr):
        try:
            self.free.append((cap, addr))
            while sum(c for c, _ in self.free) > self.max_free_bytes:
                c, a = self.free.pop(0)
                self.lib.romhc_free_host(C.c_void_p(a))
        except Exception:                        # interpreter shutdown
            pass


PINNED_POOL = PinnedPool()


class Engine:
    def __init__(self, blocks_geometry, N, device=None):
        if not torch.cuda.is_available():
            raise _lib.RomhcError("no CUDA device: the ROMHighContrast B200 path has no CPU fallback")
        self.lib = _lib.load()
        nrb, ncb = int(blocks_geometry[0]), int(blocks_geometry[1])
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        self.handle = C.c_void_p()
        _lib.check(self.lib.romhc_create(nrb, ncb, int(N), self.device.index, C.byref(self.handle)))
        info = (C.c_int64 * 16)()
        _lib.check(self.lib.romhc_get_info(self.handle, info))
        (self.D, self.Dp, self.P, self.R, self.C, self.nlevels, self.tail_level, self.coarse_D, self.coarse_direct,
         self.nrb, self.ncb, self.N, self.solve_bytes_per_system, self.tail_smem, self.bridge_level,
         self.bridge_N) = [int(v) for v in info[:16]]
        self.nb = nrb * ncb
        self.last_solve_stats = None

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.romhc_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ---- plumbing ---------------------------------------------------------------------------------
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def set_option(self, name, value):
        _lib.check(self.lib.romhc_set_option(self.handle, name.encode(), float(value)))

    def dev(self, a, dtype=F64):
        """ndarray / list / tensor -> contiguous device tensor."""
        if isinstance(a, torch.Tensor):
            return a.to(device=self.device, dtype=dtype).contiguous()
        return torch.as_tensor(np.ascontiguousarray(np.asarray(a, dtype=np.float64)), dtype=dtype).to(self.device)

    def empty(self, *shape, dtype=F64):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    def params(self, a):
        """(K, nrb, ncb) array-like -> (K, nb) device tensor."""
        y = self.dev(a)
        return y.reshape(-1, self.nb).contiguous()

    # ---- layout -----------------------------------------------------------------------------------
    HOST_PIPELINE_MIN_BYTES = 8 << 20       # smaller host arrays take torch's plain copy

    def pad(self, compact):
        """(K, D) compact (reference layout) -> (K, Dp) padded-grid device tensor.  Large host arrays go through the
        pipelined host entry point (romhc_pack_host: pinned bounce buffers, several copy threads)."""
        if not isinstance(compact, torch.Tensor):
            a = np.ascontiguousarray(np.asarray(compact, dtype=np.float64)).reshape(-1, self.D)
            if a.nbytes >= self.HOST_PIPELINE_MIN_BYTES:
                out = self.empty(a.shape[0], self.Dp)
                _lib.check(self.lib.romhc_pack_host(self.handle, C.c_void_p(a.ctypes.data), _ptr(out), a.shape[0],
                                                    self.stream()))
                return out
            compact = a
        c = self.dev(compact).reshape(-1, self.D)
        out = self.empty(c.shape[0], self.Dp)
        _lib.check(self.lib.romhc_pack(self.handle, _ptr(c), _ptr(out), c.shape[0], self.stream()))
        return out

    def unpad(self, padded):
        K = padded.shape[0]
        out = self.empty(K, self.D)
        _lib.check(self.lib.romhc_unpack(self.handle, _ptr(padded), _ptr(out), K, self.stream()))
        return out

    def host_result(self, shape):
        """Destination of a (K, D) result that leaves through the numpy API: a pooled pinned block for large results (the
        D2H copies land in it directly), a plain numpy array for small ones or when pinning fails."""
        if int(np.prod(shape)) * 8 >= self.HOST_PIPELINE_MIN_BYTES:
            arr = PINNED_POOL.take(tuple(shape))
            if arr is not None:
                return arr
        return np.empty(shape)

    def unpad_host(self, padded, out=None):
        """(K, Dp) padded device tensor -> (K, D) numpy array (fresh unless `out` is given); the pipelined counterpart
        of unpad(...).cpu().numpy() for results that leave through the reference's API."""
        K = padded.shape[0]
        U = self.host_result((K, self.D)) if out is None else out
        if U.nbytes < self.HOST_PIPELINE_MIN_BYTES:
            U[...] = self.unpad(padded).cpu().numpy()
            return U
        _lib.check(self.lib.romhc_unpack_host(self.handle, _ptr(padded.contiguous()), C.c_void_p(U.ctypes.data), K,
                                              self.stream()))
        return U

    # ---- K1a / K2 -----------------------------------------------------------------------------------
    def apply(self, y, u_pad):
        out = torch.zeros_like(u_pad)
        _lib.check(self.lib.romhc_apply(self.handle, _ptr(y), _ptr(u_pad), _ptr(out), u_pad.shape[0], self.stream()))
        return out

    def energy_norm(self, y, u_pad):
        out = self.empty(u_pad.shape[0])
        _lib.check(self.lib.romhc_energy_norm(self.handle, _ptr(y), _ptr(u_pad), u_pad.shape[0], _ptr(out),
                                              self.stream()))
        return out

    def h10_norm(self, u_pad):
        return self.energy_norm(None, u_pad)

    def l2_norm(self, u_pad):
        out = self.empty(u_pad.shape[0])
        _lib.check(self.lib.romhc_l2_norm(self.handle, _ptr(u_pad), u_pad.shape[0], _ptr(out), self.stream()))
        return out

    def error_norm(self, U_pad, coef, basis_pad):
        """|| coef @ basis - U ||_{H10} per row (coef None / empty basis: ||U||)."""
        K = U_pad.shape[0]
        n = 0 if basis_pad is None else basis_pad.shape[0]
        out = self.empty(K)
        _lib.check(self.lib.romhc_error_norm(self.handle, _ptr(U_pad), _ptr(coef) if n else None,
                                             _ptr(basis_pad) if n else None, n, K, _ptr(out), self.stream()))
        return out

    # ---- K1 --------------------------------------------------------------------------------------------
    def solve(self, y, out=None, rhs=None):
        """Batched snapshot solves; y (K, nb) device -> (x_pad (K, Dp), iters (K,), relres (K,)).

        rhs (K, Dp) padded device tensor: caller-supplied right-hand sides (default: the reference's load vector
        b = 1 / N^2); with rhs given, y may be None (a == 1: the H10 operator A_1)."""
        K = y.shape[0] if y is not None else rhs.shape[0]
        x = self.empty(K, self.Dp) if out is None else out
        iters = self.empty(K, dtype=torch.int32)
        relres = self.empty(K)
        stats = (C.c_int64 * 4)()
        if rhs is None:
            _lib.check(self.lib.romhc_solve(self.handle, _ptr(y), K, _ptr(x), _ptr(iters), _ptr(relres), self.stream(),
                                            stats))
        else:
            assert rhs.shape == (K, self.Dp)
            _lib.check(self.lib.romhc_solve_rhs(self.handle, _ptr(y), _ptr(rhs), K, _ptr(x), _ptr(iters), _ptr(relres),
                                                self.stream(), stats))
        self.last_solve_stats = {"launched_iterations": int(stats[0]), "chunks": int(stats[1]),
                                 "status": int(stats[2]), "workspace_bytes": int(stats[3])}
        return x, iters, relres

    def check_guards(self):
        """Bytes of the workspace guard zones (option "ws_guard") that were overwritten since the workspace was laid out."""
        n_bad = C.c_int64(0)
        _lib.check(self.lib.romhc_check_guards(self.handle, C.byref(n_bad)))
        return int(n_bad.value)

    def precond(self, y, r_pad):
        z = torch.zeros_like(r_pad)
        _lib.check(self.lib.romhc_precond(self.handle, _ptr(y), _ptr(r_pad), _ptr(z), r_pad.shape[0], self.stream()))
        return z

    # ---- K4 / K5 ----------------------------------------------------------------------------------------
    def project_operators(self, basis_pad):
        n = basis_pad.shape[0]
        Ahat = self.empty(self.nb, n, n)
        bhat = self.empty(n)
        _lib.check(self.lib.romhc_project_operators(self.handle, _ptr(basis_pad), n, _ptr(Ahat), _ptr(bhat),
                                                    self.stream()))
        return Ahat, bhat

    def reduced_solve(self, y, Ahat, rhs, check=True, return_info=False):
        """(sum_q y[k,q] Ahat[q]) c_k = rhs ; rhs (n,) shared or (K, n).  return_info: also the per-system flag tensor
        (1: not positive definite) so that a device-resident loop can defer the check."""
        K, n = y.shape[0], Ahat.shape[-1]
        per = 1 if rhs.dim() == 2 else 0
        Cc = self.empty(K, n)
        info = self.empty(K, dtype=torch.int32)
        _lib.check(self.lib.romhc_reduced_solve(_ptr(y), self.nb, _ptr(Ahat), _ptr(rhs), per, n, K, _ptr(Cc),
                                                _ptr(info), self.stream()))
        if check and bool(info.any().item()):
            raise np.linalg.LinAlgError("reduced Galerkin matrix is not positive definite")
        return (Cc, info) if return_info else Cc

    # ---- dense helpers -------------------------------------------------------------------------------------
    def gemm_nt(self, A, B, symmetric=False, splitk=False):
        """A B^T on the fp64 tensor cores.  splitk: small output with a long contraction (Krylov-basis products):
        the contraction is split over enough CTAs to fill the GPU, partial sums reduced in a fixed order."""
        M, Kd = A.shape
        Nn = B.shape[0]
        out = self.empty(M, Nn)
        _lib.check(self.lib.romhc_gemm_nt(_ptr(A), A.stride(0), _ptr(B), B.stride(0), _ptr(out), Nn, M, Nn, Kd,
                                          1 if symmetric else (2 if splitk else 0), self.stream()))
        return out

    def gemm_nn(self, A, B):
        M, Kd = A.shape
        Nn = B.shape[1]
        out = self.empty(M, Nn)
        _lib.check(self.lib.romhc_gemm_nn(_ptr(A), A.stride(0), _ptr(B), B.stride(0), _ptr(out), Nn, M, Nn, Kd,
                                          self.stream()))
        return out

    def gemm_tn(self, A, B):
        Kd, M = A.shape
        Nn = B.shape[1]
        out = self.empty(M, Nn)
        _lib.check(self.lib.romhc_gemm_tn(_ptr(A), A.stride(0), _ptr(B), B.stride(0), _ptr(out), Nn, M, Nn, Kd,
                                          self.stream()))
        return out

    def tsqr_r(self, W):
        """R (b, b) upper triangular with W W^T = R^T R for a row block W (b <= 32, Dp): Householder TSQR on the device."""
        b, Dp = W.shape
        R = self.empty(b, b)
        _lib.check(self.lib.romhc_tsqr_r(_ptr(W), W.stride(0), b, Dp, _ptr(R), self.stream()))
        return R

    def row_dots(self, X, Y):
        """out[k] = X[k] . Y[k]; Y with a single row is shared by all rows of X (a GEMV that streams X once)."""
        out = self.empty(X.shape[0])
        ldy = 0 if Y.shape[0] == 1 else Y.stride(0)
        _lib.check(self.lib.romhc_row_dots(_ptr(X), X.stride(0), _ptr(Y), ldy, X.shape[0], X.shape[1], _ptr(out), self.stream()))
        return out

    def row_norms(self, X):
        out = self.empty(X.shape[0])
        _lib.check(self.lib.romhc_row_norms(_ptr(X), X.stride(0), X.shape[0], X.shape[1], _ptr(out), self.stream()))
        return out

    def column_mean(self, X):
        K, D = X.shape
        mean = self.empty(D)
        _lib.check(self.lib.romhc_column_mean(_ptr(X), X.stride(0), K, D, _ptr(mean), self.stream()))
        return mean

    def center_rows_(self, X, mean):
        K, D = X.shape
        _lib.check(self.lib.romhc_center_rows(_ptr(X), X.stride(0), K, D, _ptr(mean), self.stream()))
        return X

    # ---- K6 -----------------------------------------------------------------------------------------------------
    def evaluate(self, points, u_pad):
        pts = self.dev(points).reshape(-1, 2).contiguous()
        K, m = u_pad.shape[0], pts.shape[0]
        out = self.empty(K, m)
        _lib.check(self.lib.romhc_evaluate(self.handle, _ptr(pts), m, _ptr(u_pad), K, _ptr(out), self.stream()))
        return out

    def estimator(self, c, a_basis, invert):
        """c (n, K), a_basis (n, nb) -> (K, nb)."""
        n, K = c.shape
        out = self.empty(K, a_basis.shape[1])
        _lib.check(self.lib.romhc_estimator(_ptr(c), K, n, _ptr(a_basis), a_basis.shape[1], 1 if invert else 0,
                                            _ptr(out), self.stream()))
        return out

    def poly_features(self, basis_pad, terms):
        """terms (F, degree) int array (-1 = unused factor) -> (F, Dp) products of basis rows at every slot."""
        t = torch.as_tensor(np.ascontiguousarray(np.asarray(terms, dtype=np.int32)), device=self.device)
        F, deg = t.shape
        out = self.empty(F, basis_pad.shape[1])
        _lib.check(self.lib.romhc_poly_features(_ptr(basis_pad), basis_pad.stride(0), basis_pad.shape[0], basis_pad.shape[1],
                                                _ptr(t), F, deg, _ptr(out), out.stride(0), self.stream()))
        return out

    def argmax_dev(self, v):
        """np.argmax of a device vector without leaving the device: (index (1,) int64, value (1,)) tensors."""
        idx = self.empty(1, dtype=torch.int64)
        val = self.empty(1)
        _lib.check(self.lib.romhc_argmax(_ptr(v.contiguous()), v.shape[0], _ptr(idx), _ptr(val), self.stream()))
        return idx, val

    def argmax(self, v):
        idx, val = self.argmax_dev(v)
        return int(idx.item()), float(val.item())

    # ---- host-buffer entry points -----------------------------------------------------------------------------------
    def generate_solutions_host(self, y_host, out=None, return_stats=False):
        y = np.ascontiguousarray(np.asarray(y_host, dtype=np.float64).reshape(-1, self.nb))
        K = y.shape[0]
        U = self.host_result((K, self.D)) if out is None else out
        iters = np.empty(K, dtype=np.int32)
        relres = np.empty(K)
        _lib.check(self.lib.romhc_generate_solutions_host(
            self.handle, y.ctypes.data_as(C.c_void_p), K, C.c_void_p(U.ctypes.data) if isinstance(U, np.ndarray)
            else C.c_void_p(U.data_ptr()), iters.ctypes.data_as(C.c_void_p), relres.ctypes.data_as(C.c_void_p)))
        return (U, iters, relres) if return_stats else U

    def reduced_galerkin_host(self, y_host, Ahat_host, bhat_host, out=None):
        y = np.ascontiguousarray(np.asarray(y_host, dtype=np.float64).reshape(-1, self.nb))
        A = np.ascontiguousarray(np.asarray(Ahat_host, dtype=np.float64))
        b = np.ascontiguousarray(np.asarray(bhat_host, dtype=np.float64))
        K, n = y.shape[0], b.shape[0]
        Cc = np.empty((K, n)) if out is None else out          # pass pinned memory to let the copies overlap the solves
        info = np.empty(K, dtype=np.int32)
        _lib.check(self.lib.romhc_reduced_galerkin_host(
            self.handle, y.ctypes.data_as(C.c_void_p), A.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), n, K,
            Cc.ctypes.data_as(C.c_void_p), info.ctypes.data_as(C.c_void_p)))
        if info.any():
            raise np.linalg.LinAlgError("reduced Galerkin matrix is not positive definite")
        return Cc
