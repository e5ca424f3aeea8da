"""Snapshot-free weak greedy driven by the residual a-posteriori estimator (additive, not in the reference).

The reference's greedy (`ReducedBasis.py:112-139`, mirrored in `lib/ReducedBasis.py`) needs the K training snapshots up
front and sweeps their true H10 errors every round.  The classical reduced-basis alternative computes ONE snapshot per
round: with the affine operator A(y) = sum_q y_q A_q (q = subdomains) and the reduced solution u_n = c^T Phi,

    r(y) = b - sum_q y_q A_q Phi^T c,     ||r(y)||^2_{A_1^-1} = v^T G v,    v = [1, -(y_q c_j)_{j,q}],

where G = R^T A_1 R is the Gram matrix of the Riesz representers R = A_1^-1 [b, A_q phi_j] (A_1 = the H10 operator,
`SolutionsManagers.py:49`).  Since min(y) A_1 <= A(y) <= max(y) A_1,

    ||r||_{A_1^-1} / max(y)  <=  ||u - u_n||_{H10}  <=  ||r||_{A_1^-1} / min(y).

Device work per round, all through the C ABI: nb stencil applies A_q phi (`romhc_apply` with unit parameter vectors), nb
batched GMG-PCG solves with the H10 operator (`romhc_solve_rhs`, y = NULL), one tall-skinny Gram update
(`romhc_gemm_nt`), K reduced Galerkin solves (`romhc_reduced_solve`), the K quadratic forms (`romhc_gemm_nn`) and an
argmax; one snapshot solve (`romhc_solve`) for the selected parameter.  Offline cost: n snapshot solves instead of K.
"""
from __future__ import annotations

import numpy as np
import torch

from .lib.ReducedBasis import BaseReducedBasis

GREEDY_FOR_RESIDUAL = "residual"


class ReducedBasisGreedyResidual(BaseReducedBasis):
    def __init__(self, use_coercivity=True, relative=False, chunk=262144):
        self.greedy_for = GREEDY_FOR_RESIDUAL
        self.name = "Greedy residual"
        self.linestyle = "dotted"
        self.use_coercivity = use_coercivity      # bound = estimator / min(y) (upper bound of the H10 error)
        self.relative = relative                  # divide by ||u_n||_{H10} once the basis is not empty
        self.chunk = int(chunk)
        super().__init__()

    # ---- estimator pieces -------------------------------------------------------------------------------------------
    @staticmethod
    def _riesz(eng, rhs_pad):
        """A_1 g = rhs for every row (padded layout in, padded layout out)."""
        g, _, _ = eng.solve(None, rhs=rhs_pad.contiguous())
        return g

    def _extend(self, eng, phi_pad, state):
        """new basis vector phi (1, Dp): nb applies A_q phi, nb Riesz solves, Gram rows."""
        nb = eng.nb
        eye = torch.eye(nb, dtype=torch.float64, device=eng.device)
        rhs = eng.apply(eye, phi_pad.expand(nb, -1).contiguous())             # (nb, Dp): A_q phi
        g = self._riesz(eng, rhs)
        self._append(eng, rhs, g, state)

    @staticmethod
    def _append(eng, rhs, g, state):
        """append Riesz pairs (rhs_i, g_i = A_1^-1 rhs_i); G[i, j] = rhs_i . g_j (= g_i^T A_1 g_j)."""
        RHS = rhs if state["RHS"] is None else torch.cat([state["RHS"], rhs])
        R = g if state["R"] is None else torch.cat([state["R"], g])
        m_old, m = (0 if state["G"] is None else state["G"].shape[0]), RHS.shape[0]
        G = torch.zeros(m, m, dtype=torch.float64, device=eng.device)
        if m_old:
            G[:m_old, :m_old] = state["G"]
        new = eng.gemm_nt(rhs.contiguous(), R.contiguous())                     # (m - m_old, m)
        G[m_old:, :] = new
        G[:, m_old:] = new.T
        state.update(RHS=RHS, R=R, G=0.5 * (G + G.T))

    def estimate(self, eng, y, Phi_pad, state, return_coefs=False):
        """||r(y_k)||_{A_1^-1} for every row of y (K, nb) with the current basis; optionally the coefficients."""
        K, nb = y.shape
        n = 0 if Phi_pad is None else Phi_pad.shape[0]
        G = state["G"]
        if n == 0:
            est = torch.sqrt(G[0, 0]).expand(K).clone()
            return (est, None, None) if return_coefs else est
        Ahat, bhat = eng.project_operators(Phi_pad)
        est = torch.empty(K, dtype=torch.float64, device=eng.device)
        un = torch.empty(K, dtype=torch.float64, device=eng.device)
        A1hat = Ahat.sum(dim=0)
        Cs = []
        for k0 in range(0, K, self.chunk):
            yk = y[k0:k0 + self.chunk].contiguous()
            Cc = eng.reduced_solve(yk, Ahat, bhat)                             # (k, n)
            V = torch.empty(yk.shape[0], 1 + nb * n, dtype=torch.float64, device=eng.device)
            V[:, 0] = 1.0
            V[:, 1:] = -(Cc[:, :, None] * yk[:, None, :]).reshape(yk.shape[0], n * nb)   # index 1 + j * nb + q
            W = eng.gemm_nn(V, G)
            est[k0:k0 + self.chunk] = torch.sqrt(torch.clamp((W * V).sum(dim=1), min=0.0))
            un[k0:k0 + self.chunk] = torch.sqrt(torch.clamp((eng.gemm_nn(Cc.contiguous(), A1hat.contiguous()) * Cc).sum(dim=1), min=0.0))
            if return_coefs:
                Cs.append(Cc)
        return (est, un, torch.cat(Cs)) if return_coefs else (est, un)

    # ---- the builder ------------------------------------------------------------------------------------------------
    def build(self, n: int, sm, a2train, solutions2train=None, **kwargs):
        """n greedy rounds over the parameters `a2train` (K, nrb, ncb); `solutions2train` is accepted for signature
        compatibility with the reference's builders and ignored (no snapshot of the training set is needed)."""
        eng = sm._engine_()
        a2train = np.asarray(a2train, dtype=np.float64).reshape((-1,) + tuple(sm.blocks_geometry))
        y = eng.params(a2train)
        K = y.shape[0]
        ymin = y.min(dim=1).values
        state = {"RHS": None, "R": None, "G": None}
        b = eng.pad(np.asarray(sm.B_total, dtype=np.float64).reshape(1, -1))
        self._append(eng, b, self._riesz(eng, b), state)
        Q = None                                                                # (j, Dp) orthonormal rows (Euclidean)
        basis, a, self.selected_indices, self.max_estimates = [], [], [], []
        self.snapshot_solves = 0
        for _ in range(n):
            if Q is None:
                score = self.estimate(eng, y, None, state)
            else:
                est, un = self.estimate(eng, y, Q, state)
                score = est / un if self.relative else est
            if self.use_coercivity:
                score = score / ymin
            idx, val = eng.argmax(score.contiguous())
            u, _, _ = eng.solve(y[idx:idx + 1].contiguous())                    # the round's only snapshot
            self.snapshot_solves += 1
            phi = u.clone()
            if Q is not None:
                for _ in range(2):                                              # Gram-Schmidt, twice
                    c = eng.gemm_nt(Q.contiguous(), phi.contiguous(), splitk=True)          # (j, 1) = Q phi^T (split-K DMMA product)
                    phi = phi - eng.gemm_nn(c.T.contiguous(), Q.contiguous())                # (1, Dp)
            nrm = float(torch.linalg.vector_norm(phi))
            if not nrm > 1e-10 * float(torch.linalg.vector_norm(u)):
                break                                                           # already in the span: nothing to add
            phi = phi / nrm
            Q = phi if Q is None else torch.cat([Q, phi])
            self._extend(eng, phi, state)
            self.selected_indices.append(idx)
            self.max_estimates.append(val)
            basis.append(eng.unpad(u).cpu().numpy()[0])
            a.append(a2train[idx])
        self.basis_orth = eng.unpad(Q).cpu().numpy() if Q is not None else np.empty((0, 0))
        self._state, self._Q = state, Q
        super().set(basis=np.asarray(basis), a=a)
        return self

    def error_bounds(self, sm, a):
        """(lower, upper) bounds of ||u(a) - u_n(a)||_{H10} for every parameter in `a` with the built basis."""
        eng = sm._engine_()
        y = eng.params(np.asarray(a, dtype=np.float64).reshape((-1,) + tuple(sm.blocks_geometry)))
        est = self.estimate(eng, y, self._Q, self._state)
        est = est if self._Q is None else est[0]
        return (est / y.max(dim=1).values).cpu().numpy(), (est / y.min(dim=1).values).cpu().numpy()
