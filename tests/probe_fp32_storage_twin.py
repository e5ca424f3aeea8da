"""Probe (CPU only): would fp32 STORAGE of the vectors internal to the multigrid preconditioner (z_A, z_B, coarse
residuals and corrections; arithmetic stays fp64) change the PCG iteration count or the attainable accuracy?
numpy twin of the solver, (4,4) subdomains, N = 16, contrast 1e6 and 1e10.  Result (round 1): iteration counts equal
in 7 of 8 cases (+1 in one), relative error against the sparse direct oracle unchanged (1e-13 .. 1e-14).
Usage: python tests/probe_fp32_storage_twin.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, gmg_twin as T
f32=lambda a: a.astype(np.float32).astype(np.float64)
class G32(T.GMG):
    def vcycle(self, r, l=0):
        if l == len(self.levels) - 1:
            return f32(self.coarse_solve(r))
        L = self.levels[l]
        in_tail = (L.R + 1) * ((L.C + 7) // 8 * 8) <= self.TAIL_MAX_DP
        nu = self.nu_tail if in_tail else (self.nu if l == 0 else self.nu_mid)
        z = np.zeros_like(r)
        for _ in range(nu):
            L.gs_half(z, r, L.red); L.gs_half(z, r, L.black)
        d = r - L.apply(z); d[~L.mask] = 0
        z = f32(z)                                   # z_A stored as fp32
        rc = f32(T.restrict(d))                      # coarse residual stored as fp32
        z = z + T.prolong(self.vcycle(rc, l + 1), r.shape)
        z[~L.mask] = 0
        for _ in range(nu):
            L.gs_half(z, r, L.black); L.gs_half(z, r, L.red)
        return f32(z)                                # z_B stored as fp32
def pcg(g, N, tol=1e-12, maxit=100):
    L = g.levels[0]
    b = np.zeros((L.R + 1, L.C + 1)); b[1:-1, 1:-1] = 1.0 / N ** 2
    x = np.zeros_like(b); r = b.copy()
    z = g.vcycle(r); p = z.copy(); rz = (r * z).sum(); rz0 = rz
    for it in range(1, maxit + 1):
        Ap = L.apply(p); al = rz / (p * Ap).sum()
        x += al * p; r -= al * Ap
        z = g.vcycle(r); rzn = (r * z).sum()
        if not rzn > tol ** 2 * rz0: break
        p = z + (rzn / rz) * p; rz = rzn
    return x, it
from oracle import FEMOracle
geo,N=(4,4),16
rng=np.random.default_rng(1)
for cmax in (1e6,1e10):
    for s in range(4):
        a=10**rng.uniform(0,np.log10(cmax),geo)
        x64,it64=pcg(T.GMG(a,N),N); x32,it32=pcg(G32(a,N),N)
        uo=FEMOracle(geo,N).generate_solutions(a[None])[0]
        e=lambda x: np.linalg.norm(x[1:-1,1:-1].ravel()-uo)/np.linalg.norm(uo)
        print(f"cmax {cmax:g}: iterations fp64 {it64} / fp32-storage {it32}; rel err vs oracle {e(x64):.1e} / {e(x32):.1e}")
