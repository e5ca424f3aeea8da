"""numpy twin of the CUDA GMG-PCG algorithm (romhighcontrast_b200/csrc/solver.cu) on the full vertex grid.

Test infrastructure: used to check single kernels (stencil apply, one V-cycle, iteration counts)
at sizes where running it takes milliseconds.  Same algorithm, different code: red/black GS V(nu,nu) with
nu = 2 / 3 / 4 sweeps on the finest / intermediate / small levels (the library defaults),
P1 anti-diagonal transfers, rediscretised coarse operators, dense coarsest solve, difference-form apply.
When halving the cells per subdomain gets stuck on an odd count with a coarsest grid too large for the dense solve
(N = 43: 129 x 129 cells), the deepest batched level hands over to a power-of-two hierarchy through a NON-NESTED
transfer: bilinear interpolation inside each subdomain (block interfaces are grid lines of both meshes), its
transpose as restriction (`coarsening_chain`, `bridge_matrix`).
"""
import numpy as np


def cell_coef(a, N):
    return np.kron(np.asarray(a, float), np.ones((N, N)))


class Level:
    def __init__(self, a, N):
        k = cell_coef(a, N)
        self.N = N
        self.R, self.C = k.shape
        R, C = self.R, self.C
        kp = np.zeros((R + 2, C + 2))
        kp[1:-1, 1:-1] = k
        ul, ur = kp[0:R + 1, 0:C + 1], kp[0:R + 1, 1:C + 2]
        dl, dr = kp[1:R + 2, 0:C + 1], kp[1:R + 2, 1:C + 2]
        self.diag = (ul + ur) + (dl + dr)
        self.wE = 0.5 * (ur + dr)
        self.wS = 0.5 * (dl + dr)
        self.mask = np.zeros((R + 1, C + 1), bool)
        self.mask[1:R, 1:C] = True
        rr, cc = np.meshgrid(np.arange(R + 1), np.arange(C + 1), indexing="ij")
        self.red = self.mask & ((rr + cc) % 2 == 0)
        self.black = self.mask & ((rr + cc) % 2 == 1)

    def offdiag(self, u):
        s = np.zeros_like(u)
        s[:, :-1] += self.wE[:, :-1] * u[:, 1:]
        s[:, 1:] += self.wE[:, :-1] * u[:, :-1]
        s[:-1, :] += self.wS[:-1, :] * u[1:, :]
        s[1:, :] += self.wS[:-1, :] * u[:-1, :]
        return s

    def apply(self, u):
        out = np.zeros_like(u)
        dE = u[:, :-1] - u[:, 1:]
        out[:, :-1] += self.wE[:, :-1] * dE
        out[:, 1:] -= self.wE[:, :-1] * dE
        dS = u[:-1, :] - u[1:, :]
        out[:-1, :] += self.wS[:-1, :] * dS
        out[1:, :] -= self.wS[:-1, :] * dS
        out[~self.mask] = 0
        return out

    def energy(self, u):
        dE = u[:, :-1] - u[:, 1:]
        dS = u[:-1, :] - u[1:, :]
        return float((self.wE[:, :-1] * dE * dE).sum() + (self.wS[:-1, :] * dS * dS).sum())

    def gs_half(self, z, r, color):
        s = self.offdiag(z)
        z[color] = (r[color] + s[color]) / self.diag[color]


def restrict(d):
    R, C = d.shape[0] - 1, d.shape[1] - 1
    dp = np.zeros((R + 3, C + 3))
    dp[1:-1, 1:-1] = d
    c = lambda dr, dc: dp[1 + dr:R + 2 + dr:2, 1 + dc:C + 2 + dc:2]
    out = c(0, 0) + 0.5 * (c(0, 1) + c(0, -1) + c(1, 0) + c(-1, 0) + c(-1, 1) + c(1, -1))
    out[0, :] = 0; out[-1, :] = 0; out[:, 0] = 0; out[:, -1] = 0
    return out


def prolong(e, shape):
    z = np.zeros(shape)
    z[::2, ::2] = e
    z[::2, 1::2] = 0.5 * (e[:, :-1] + e[:, 1:])
    z[1::2, ::2] = 0.5 * (e[:-1, :] + e[1:, :])
    z[1::2, 1::2] = 0.5 * (e[:-1, 1:] + e[1:, :-1])
    return z


DIRECT_MAX = 64
TAIL_MAX_DP = 4352
MAX_LEVELS = 12


def coarsening_chain(nrb, ncb, N, bridge=True):
    """Cells per subdomain on every level and the index of the level whose transfer to the next one is non-nested
    (None: all transfers are the nested factor-2 ones).  Mirrors Context::build_levels."""
    def halve(chain):
        while True:
            n = chain[-1]
            if n % 2 or nrb * (n // 2) < 2 or ncb * (n // 2) < 2 or len(chain) >= MAX_LEVELS:
                return chain
            chain.append(n // 2)
    dp = lambda n: (nrb * n + 1) * ((ncb * n + 7) // 8 * 8)
    chain = halve([N])
    n = chain[-1]
    if not bridge or n % 2 == 0 or n < 3 or (nrb * n - 1) * (ncb * n - 1) <= DIRECT_MAX:
        return chain, None
    batched = [j for j, m in enumerate(chain) if dp(m) > TAIL_MAX_DP]
    if not batched or len(chain) >= MAX_LEVELS:
        return chain, None
    j = batched[-1]
    nc = 1
    while 3 * (2 * nc) <= 2 * chain[j]:          # largest power of two with ratio >= 1.5
        nc *= 2
    return halve(chain[:j + 1] + [nc]), j


def bridge_matrix(nblocks, Nf, Nc):
    """1-D interpolation (nblocks*Nf + 1, nblocks*Nc + 1): coarse hat functions sampled at the fine vertices,
    weight (Nf - |i Nc - I Nf|) / Nf where positive."""
    i = np.arange(nblocks * Nf + 1)[:, None]
    I = np.arange(nblocks * Nc + 1)[None, :]
    return np.maximum(Nf - np.abs(i * Nc - I * Nf), 0) / Nf


class GMG:
    DIRECT_MAX = DIRECT_MAX
    TAIL_MAX_DP = TAIL_MAX_DP

    def __init__(self, a, N, coarse_sweeps=8, nu=2, nu_tail=4, nu_mid=3, bridge=True):
        a = np.asarray(a, float)
        self.nu, self.nu_tail, self.nu_mid = nu, nu_tail, (nu_mid or nu)
        nrb, ncb = a.shape
        chain, self.bridge_level = coarsening_chain(nrb, ncb, N, bridge)
        self.levels = [Level(a, n) for n in chain]
        if self.bridge_level is not None:
            nf, nc = chain[self.bridge_level], chain[self.bridge_level + 1]
            self.Py, self.Px = bridge_matrix(nrb, nf, nc), bridge_matrix(ncb, nf, nc)
        self.coarse_sweeps = coarse_sweeps
        L = self.levels[-1]
        D = (L.R - 1) * (L.C - 1)
        P = (L.C + 7) // 8 * 8
        in_tail = (L.R + 1) * P <= self.TAIL_MAX_DP
        self.direct = in_tail and D <= self.DIRECT_MAX
        self.smooth_only = not in_tail          # coarsest level handled by the strip kernels: one symmetric sweep
        if self.direct:
            M = np.zeros((D, D))
            for j in range(D):
                u = np.zeros((L.R + 1, L.C + 1)); u[1:-1, 1:-1].flat[j] = 1
                M[:, j] = L.apply(u)[1:-1, 1:-1].ravel()
            self.coarse_matrix = M

    def coarse_solve(self, r):
        L = self.levels[-1]
        z = np.zeros_like(r)
        if self.direct:
            z[1:-1, 1:-1] = np.linalg.solve(self.coarse_matrix, r[1:-1, 1:-1].ravel()).reshape(L.R - 1, L.C - 1)
            return z
        sweeps = (self.nu if len(self.levels) == 1 else self.nu_mid) if self.smooth_only else self.coarse_sweeps
        for _ in range(sweeps):
            L.gs_half(z, r, L.red); L.gs_half(z, r, L.black)
        for _ in range(sweeps):
            L.gs_half(z, r, L.black); L.gs_half(z, r, L.red)
        return z

    def vcycle(self, r, l=0):
        if l == len(self.levels) - 1:
            return self.coarse_solve(r)
        L = self.levels[l]
        in_tail = (L.R + 1) * ((L.C + 7) // 8 * 8) <= self.TAIL_MAX_DP
        nu = self.nu_tail if in_tail else (self.nu if l == 0 else self.nu_mid)
        z = np.zeros_like(r)
        for _ in range(nu):
            L.gs_half(z, r, L.red); L.gs_half(z, r, L.black)
        d = r - L.apply(z); d[~L.mask] = 0
        if l == self.bridge_level:
            d[L.black] = 0                      # just relaxed: the kernels take the residual there as exactly zero
            rc = self.Py.T @ d @ self.Px
            rc[~self.levels[l + 1].mask] = 0
            z = z + self.Py @ self.vcycle(rc, l + 1) @ self.Px.T
        else:
            z = z + prolong(self.vcycle(restrict(d), l + 1), r.shape)
        z[~L.mask] = 0
        for _ in range(nu):
            L.gs_half(z, r, L.black); L.gs_half(z, r, L.red)
        return z


def pcg(a, N, tol=1e-12, maxit=1000, coarse_sweeps=8, nu=2, nu_tail=4, nu_mid=3, bridge=True):
    g = GMG(a, N, coarse_sweeps, nu, nu_tail, nu_mid, bridge)
    L = g.levels[0]
    b = np.zeros((L.R + 1, L.C + 1)); b[1:-1, 1:-1] = 1.0 / N ** 2
    x = np.zeros_like(b); r = b.copy()
    z = g.vcycle(r); p = z.copy(); rz = (r * z).sum(); rz0 = rz
    it = 0
    for it in range(1, maxit + 1):
        Ap = L.apply(p); al = rz / (p * Ap).sum()
        x += al * p; r -= al * Ap
        z = g.vcycle(r); rzn = (r * z).sum()
        if not rzn > tol ** 2 * rz0:
            break
        p = z + (rzn / rz) * p; rz = rzn
    return x[1:-1, 1:-1].ravel(), it
