"""Regression pins for the oracle's restatements of the round-1 terms: small fixed inputs -> arrays stored in forms_expected.npz.
The oracle is the checker of the CUDA path, so an accidental edit of the checker must not go unnoticed; every array here was produced
by the oracle at the commit that also passed its closed-form KATs (tests/test_oracle_forms.py).  Run from the repo root:
    python tests/golden/make_forms_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import fem_oracle as fo  # noqa: E402
from oracle import fem_oracle_p2 as fp  # noqa: E402
from oracle import mg_oracle as mo  # noqa: E402


def inputs():
    c, t = fo.unit_cube_mesh(2, 3, 2)
    rng = np.random.default_rng(2026)
    c = c + 0.05 * (rng.random(c.shape) * 2 - 1)
    T = 300 + 40 * c[:, 0] + 25 * c[:, 2] ** 2
    u = 1e-3 * rng.standard_normal((c.shape[0], 3))
    vel = np.array([0.7, -0.3, 0.2])
    return c, t, T, u, vel


def compute():
    c, t, T, u, vel = inputs()
    nv = c.shape[0]
    fv, opp, _ = fo.exterior_facets(t)
    out = {"coords": c, "cells": t, "T": T, "u": u}
    out["thermal_load"] = fo.thermal_load(c, t, 7.5e5, T, 293.0)
    cn, xn, _ = fp.p2_dofmap(c, t)
    out["thermal_load_p2"] = fp.thermal_load(c, t, cn, xn.shape[0], 7.5e5, 293.0 + 40 * xn[:, 0] + 25 * xn[:, 2] ** 2, 293.0)
    mu, lam = fo.lame(2e11, 0.27)
    out["von_mises_cells"] = fo.von_mises_cells(c, t, u, mu, lam)
    J, r = fo.radiation_terms(c, fv, T, 0.9 * 5.670367e-8, 280.0)
    out["radiation_J"], out["radiation_r"] = J, r
    kf, dkf = (lambda x: 0.6 * (1 + 0.02 * (x - 300.0))), (lambda x: 0.012 + 0 * x)
    Jk, Rk = fo.nonlinear_k_terms(c, t, T, kf, dkf)
    out["nonlinear_k_J"], out["nonlinear_k_R"] = fo.conform(Jk, *fo.csr_pattern(t, nv)).data, Rk
    out["supg_local"] = fo.local_supg(c, t, vel, 5.0, mass=2.5, adv=1.5)
    out["supg_source"] = fo.supg_source(c, t, 4.0, vel, 5.0)
    Af, bf = fo.supg_facet_terms(c, fv, opp, vel, 5.0, g=2.0, h=3.0)
    out["supg_facet_A"], out["supg_facet_b"] = fo.conform(Af, *fo.csr_pattern(t, nv)).data, bf
    out["advection_nodal"] = fo.local_advection_nodal(c, t, np.stack([-c[:, 1], c[:, 0], 0.3 * c[:, 2]], axis=1), 2.0)
    # one multigrid V-cycle on the 4^3 / 2^3 heat hierarchy with fixed dampings
    levels, transfers = [], []
    for n in (4, 2):
        cc, tt = fo.unit_cube_mesh(n, n, n)
        z0 = np.nonzero(cc[:, 2] == 0)[0]
        A, b = fo.heat_system(cc, tt, 20.0, [(z0, 350.0)], source=1000.0)
        bc = np.zeros(cc.shape[0], dtype=bool)
        bc[z0] = True
        levels.append({"A": A.tocsr(), "dinv": 1.0 / A.diagonal(), "omega": 0.62, "bc": bc})
    transfers.append(mo.prolongation((4, 4, 4)))
    res = np.sin(np.arange(levels[0]["A"].shape[0]) * 0.37)
    res[levels[0]["bc"]] = 0.0
    out["mg_residual"], out["mg_vcycle"] = res, mo.vcycle(levels, transfers, res)
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "forms_expected.npz"), **compute())
    print("wrote forms_expected.npz")
