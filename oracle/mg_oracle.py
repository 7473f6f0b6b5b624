"""CPU restatement of the geometric multigrid V-cycle of csrc/fsb_mg.cu.  TEST INFRASTRUCTURE ONLY (see fem_oracle.py).

The reference reaches multigrid through PETSc GAMG (SolverBase.py:643-672, solve_amg); an algebraic hierarchy cannot be
reproduced without PETSc, so what is pinned here is the library's own geometric scheme: nestedness of dolfin's box meshes
and Galerkin = re-discretisation (tests/test_oracle_forms.py), and this V-cycle, which the CUDA path must match to
rounding when given the same level matrices and dampings."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def prolongation(ncells_fine, ncomp=1):
    """P1 interpolation from the box with ncells/2 cells per axis: fine vertex 2C + d is the coarse vertex C (d = 0) or the
    midpoint of the coarse edge (C, C + d)."""
    nf = list(ncells_fine) + [0] * (3 - len(ncells_fine))
    df = [k + 1 for k in nf]
    dc = [k // 2 + 1 for k in nf]
    f = np.arange(df[0] * df[1] * df[2])
    i, j, k = f % df[0], (f // df[0]) % df[1], f // (df[0] * df[1])
    di, dj, dk = i & 1, j & 1, k & 1
    c0 = (i >> 1) + dc[0] * ((j >> 1) + dc[1] * (k >> 1))
    c1 = ((i >> 1) + di) + dc[0] * (((j >> 1) + dj) + dc[1] * ((k >> 1) + dk))
    P = sp.coo_matrix((np.full(2 * f.size, 0.5), (np.concatenate([f, f]), np.concatenate([c0, c1]))),
                      shape=(f.size, dc[0] * dc[1] * dc[2])).tocsr()          # d = 0: the two halves add up to 1
    return sp.kron(P, sp.identity(ncomp)).tocsr() if ncomp > 1 else P


CHEB_RATIO = 10.0


def chebyshev(L, x, rhs, deg, zero):
    """Chebyshev iteration of degree `deg` for D^-1 A on [lmax / CHEB_RATIO, lmax], lmax = 4 / (3 omega)."""
    lmax = 4.0 / (3.0 * L['omega'])
    lmin = lmax / CHEB_RATIO
    theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
    sigma = theta / delta
    rho0 = 1.0 / sigma
    d = None
    for k in range(deg):
        r = rhs if (k == 0 and zero) else rhs - L['A'] @ x
        if k == 0:
            d = L['dinv'] * r / theta
        else:
            rho1 = 1.0 / (2.0 * sigma - rho0)
            d = rho1 * rho0 * d + (2.0 * rho1 / delta) * L['dinv'] * r
            rho0 = rho1
        x = d.copy() if (k == 0 and zero) else x + d
    return x


def vcycle(levels, transfers, b, nu=2, coarse_sweeps=24, l=0):
    """levels[l] = dict(A, dinv, omega, bc mask); transfers[l] = P from level l+1 to l.  Zero initial guess; Chebyshev
    smoothing of degree nu before and after the coarse correction, damped Jacobi on the coarsest level."""
    L = levels[l]
    if l == len(levels) - 1:
        x = None
        for s in range(coarse_sweeps):
            x = L['omega'] * L['dinv'] * b if s == 0 else x + L['omega'] * L['dinv'] * (b - L['A'] @ x)
        return x
    x = chebyshev(L, None, b, nu, True)
    bc = transfers[l].T @ (b - L['A'] @ x)
    bc[levels[l + 1]['bc']] = 0.0
    corr = transfers[l] @ vcycle(levels, transfers, bc, nu, coarse_sweeps, l + 1)
    corr[L['bc']] = 0.0
    return chebyshev(L, x + corr, b, nu, False)
