"""GPU parity tests for the terms the reference's own example scripts exercise beyond plain diffusion/elasticity:
thermal stress and the von Mises projection (examples/test_linear_elasticity.py:123-141), the quasi-static transient
loop with a time-dependent stress boundary (:117-121), radiation solved by Newton's method
(examples/test_heat_transfer.py:195-222), point sources (ScalarTransportSolver.py:150-158) and `.pvd` output
(SolverBase.py:570-577).  Kernel-level checks against the oracle on the same inputs, then the API-level cases.
"""
import copy
import math
import os
import xml.etree.ElementTree as ET

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu

from oracle import fem_oracle as fo  # noqa: E402  (checker only)
from oracle import fem_oracle_p2 as fp  # noqa: E402
from fenicssolver_b200 import LinearElasticitySolver, ScalarTransportSolver, SolverBase, _lib  # noqa: E402
from fenicssolver_b200.dolfin_compat import (AutoSubDomain, BoxMesh, Constant, Expression, FunctionSpace, Point,  # noqa: E402
                                             PointSource, SubDomain, UnitSquareMesh, VectorFunctionSpace, near)

TOL = 1e-10
VAL_TOL = 1e-13
QUIET = {'logging_level': 40, 'logging_file': None, 'plotting_freq': 0, 'saving_freq': 0, 'plotting_interactive': False}


def jitter(coords, n, seed=0, amp=0.2):
    rng = np.random.default_rng(seed)
    return coords + amp / n * (rng.random(coords.shape) * 2 - 1)


def small_mesh(dim, n=3, seed=1):
    c, t = (fo.unit_square_mesh(n, n + 1) if dim == 2 else fo.unit_cube_mesh(n, n + 1, n))
    return jitter(c, n + 1, seed), t


def close(dev, ref, tol=VAL_TOL):
    scale = np.abs(ref).max()
    assert np.abs(dev - ref).max() <= tol * scale, "max abs err %.3e vs scale %.3e" % (np.abs(dev - ref).max(), scale)


# ----------------------------------------------------------------------------------- kernels
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("nodal", [False, True])
def test_thermal_load_p1(ctx, dim, nodal):
    c, t = small_mesh(dim)
    nv = c.shape[0]
    T = 293.0 + 40 * c[:, 0] + 25 * c[:, -1] ** 2 if nodal else 343.0
    m = _lib.DeviceMesh.upload(ctx, c, t)
    b = _lib.DeviceVector(ctx, nv * dim)
    if nodal:
        _lib.assemble_thermal_load(m, b, 7.5e5, T=_lib.DeviceVector.from_numpy(ctx, T), T_ref=293.0)
    else:
        _lib.assemble_thermal_load(m, b, 7.5e5, T_const=T, T_ref=293.0)
    close(b.numpy(), fo.thermal_load(c, t, 7.5e5, T, 293.0))
    _lib.assemble_thermal_load(m, b, 7.5e5, T_const=343.0, T_ref=293.0, scale=-1.0)     # accumulates, scale applies
    if not nodal:
        assert np.abs(b.numpy()).max() <= 1e-12 * 7.5e5 * 50


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("nodal", [False, True])
def test_thermal_load_p2(ctx, dim, nodal):
    c, t = small_mesh(dim)
    cn, xn, _ = fp.p2_dofmap(c, t)
    nn = xn.shape[0]
    T = 293.0 + 40 * xn[:, 0] + 25 * xn[:, -1] ** 2 if nodal else 343.0
    m = _lib.DeviceMesh.upload_p2(ctx, c, cn, nn)
    b = _lib.DeviceVector(ctx, nn * dim)
    if nodal:
        _lib.assemble_thermal_load(m, b, 7.5e5, T=_lib.DeviceVector.from_numpy(ctx, T), T_ref=293.0)
    else:
        _lib.assemble_thermal_load(m, b, 7.5e5, T_const=T, T_ref=293.0)
    ref = fp.thermal_load(c, t, cn, nn, 7.5e5, T, 293.0)
    close(b.numpy(), ref)
    # divergence theorem: sum_a x_a[i] b[(a,i)] = beta int (T - T_ref) dx for every axis i (x is in the P2 space)
    if nodal:
        vol, _ = fo.p1_geometry(c, t)
        _, M, _, _ = fp.reference_tensors(dim)
        integral = 7.5e5 * float(np.sum(vol[:, None] * np.einsum("ij,cj->ci", M, (T - 293.0)[cn])))
        got = (xn * b.numpy().reshape(-1, dim)).sum(axis=0)
        assert np.allclose(got, integral, rtol=1e-11)


@pytest.mark.parametrize("dim", [2, 3])
def test_von_mises_load_p1_and_p2(ctx, dim):
    c, t = small_mesh(dim)
    nv = c.shape[0]
    mu, lam = fo.lame(2e11, 0.27)
    rng = np.random.default_rng(5)
    u = 1e-3 * rng.standard_normal((nv, dim))
    m = _lib.DeviceMesh.upload(ctx, c, t)
    b = _lib.DeviceVector(ctx, nv)
    _lib.assemble_von_mises_load(m, _lib.DeviceVector.from_numpy(ctx, u.ravel()), mu, lam, b)
    vol, _ = fo.p1_geometry(c, t)
    vm = fo.von_mises_cells(c, t, u, mu, lam)
    ref = np.zeros(nv)
    np.add.at(ref, t.ravel(), np.repeat(vm * vol / (dim + 1), dim + 1))
    close(b.numpy(), ref)
    # degree 2: same quadrature rule restated in numpy; a quadratic field with a non-polynomial von Mises stress
    cn, xn, _ = fp.p2_dofmap(c, t)
    nn = xn.shape[0]
    u2 = 1e-3 * np.stack([xn[:, 0] ** 2 + 0.3 * xn[:, 1], xn[:, 0] * xn[:, -1]] + ([0.5 * xn[:, 1] ** 2] if dim == 3 else []), axis=1)
    m2 = _lib.DeviceMesh.upload_p2(ctx, c, cn, nn)
    b2 = _lib.DeviceVector(ctx, nv)
    _lib.assemble_von_mises_load(m2, _lib.DeviceVector.from_numpy(ctx, u2.ravel()), mu, lam, b2)
    close(b2.numpy(), fp.von_mises_load(c, t, cn, u2, mu, lam), 1e-12)
    # a linear field in the P2 space gives the P1 answer exactly (constant stress)
    ul = 1e-3 * (xn @ rng.standard_normal((dim, dim)))
    b3 = _lib.DeviceVector(ctx, nv)
    _lib.assemble_von_mises_load(m2, _lib.DeviceVector.from_numpy(ctx, ul.ravel()), mu, lam, b3)
    vml = fo.von_mises_cells(c, t, ul[:nv], mu, lam)
    ref3 = np.zeros(nv)
    np.add.at(ref3, t.ravel(), np.repeat(vml * vol / (dim + 1), dim + 1))
    close(b3.numpy(), ref3, 1e-12)


@pytest.mark.parametrize("dim", [2, 3])
def test_facet_radiation_terms(ctx, dim):
    c, t = small_mesh(dim)
    nv = c.shape[0]
    fv, _, _ = fo.exterior_facets(t)
    T = 300.0 + 60 * c[:, 0] + 20 * c[:, -1] ** 2
    mco, Ta = 0.9 * 5.670367e-8, 280.0
    m = _lib.DeviceMesh.upload(ctx, c, t)
    A = _lib.DeviceMatrix.create(m, 1)
    r = _lib.DeviceVector(ctx, nv)
    Td = _lib.DeviceVector.from_numpy(ctx, T)
    _lib.assemble_facet_radiation(m, A, r, Td, fv, mco, Ta, rscale=-1.0)
    Jf, rf = fo.radiation_terms(c, fv, T, mco, Ta)
    rref = np.zeros(nv)
    np.add.at(rref, fv.ravel(), rf.ravel())
    close(r.numpy(), -rref, 1e-12)
    rp, ci, va = A.download_csr()
    Jref = fo.conform(fo._scatter(fv, Jf, nv), *fo.csr_pattern(t, nv))
    close(va, Jref.data, 1e-12)
    # derivative check: J is d r / d T (finite differences along a random direction)
    d = np.random.default_rng(2).standard_normal(nv)
    eps = 1e-3
    _, rp_ = fo.radiation_terms(c, fv, T + eps * d, mco, Ta)
    _, rm_ = fo.radiation_terms(c, fv, T - eps * d, mco, Ta)
    fd = np.zeros(nv)
    np.add.at(fd, fv.ravel(), ((rp_ - rm_) / (2 * eps)).ravel())
    Jd = sp.csr_matrix((va, ci.astype(np.int64), rp), shape=(nv, nv)) @ d
    assert np.abs(Jd - fd).max() <= 1e-6 * np.abs(fd).max()


# ----------------------------------------------------------------------------------- test_heat_transfer.py:195-222
cx_min, cy_min, cx_max, cy_max = 0, 0, 1, 1
top = AutoSubDomain(lambda x: near(x[1], cy_max))
bottom = AutoSubDomain(lambda x: near(x[1], cy_min))
left = AutoSubDomain(lambda x: near(x[0], cx_min))
right = AutoSubDomain(lambda x: near(x[0], cx_max))


def radiation_settings(n, transient=None):
    T_hot, T_cold, T_ambient = 360, 300, 300
    mesh = UnitSquareMesh(n, n)
    Q = FunctionSpace(mesh, "CG", 1)
    bcs = {"hot": {'boundary': top, 'boundary_id': 1, 'values': {'temperature': {'variable': 'temperature', 'type': 'Dirichlet', 'value': Constant(T_hot)}}},
           "left": {'boundary': left, 'boundary_id': 3, 'values': {'temperature': {'variable': 'temperature', 'type': 'heatFlux', 'value': Constant(0)}}},
           "right": {'boundary': right, 'boundary_id': 4, 'values': {'temperature': {'variable': 'temperature', 'type': 'symmetry', 'value': None}}},
           "cold": {'boundary': bottom, 'boundary_id': 2, 'values': {'temperature': {'variable': 'temperature', 'type': 'Dirichlet', 'value': Constant(T_cold)}}}}
    return {'solver_name': 'ScalarEquationSolver', 'mesh': None, 'function_space': Q, 'periodic_boundary': None, 'fe_degree': 1,
            'boundary_conditions': bcs, 'body_source': None, 'initial_values': {'temperature': T_ambient},
            'material': {'density': 1000, 'specific_heat_capacity': 4200, 'thermal_conductivity': 0.1},
            'solver_settings': {'transient_settings': transient or {'transient': False, 'starting_time': 0, 'time_step': 0.1, 'ending_time': 1},
                                'reference_values': {'temperature': T_ambient},
                                'solver_parameters': {"relative_tolerance": 1e-9, "maximum_iterations": 500, "monitor_convergence": True}},
            'scalar_name': 'temperature', 'report_settings': QUIET,
            'radiation_settings': {'ambient_temperature': T_ambient - 20, 'emissivity': 0.9}}, mesh


def test_radiation_example_newton_matches_oracle():
    """test_radiation(): Dirichlet 360/300 on top/bottom, grey-body radiation to 280 K over the whole exterior surface,
    emissivity 0.9, conductivity 0.6 set on the solver after construction (:209-216)."""
    n = 20
    settings, mesh = radiation_settings(n)
    solver = ScalarTransportSolver.ScalarTransportSolver(settings)
    solver.material['conductivity'] = 0.6
    solver.material['emissivity'] = 0.9
    T = solver.solve()
    info = solver.solve_info
    assert info['converged'] == 1 and 2 <= info['newton_iterations'] <= 8
    res = info['newton_residuals']
    assert res[-1] <= 1e-11 * res[0]
    c, t = fo.unit_square_mesh(n, n)
    fv, _, _ = fo.exterior_facets(t)
    tp, bt = np.nonzero(c[:, 1] == 1)[0], np.nonzero(c[:, 1] == 0)[0]
    To, its = fo.solve_radiation_newton(c, t, 0.6, [(tp, 360.0), (bt, 300.0)], fv, 0.9 * 5.670367e-8, 280.0, 300.0)
    assert fo.relative_l2(T.vector().get_local(), To) < TOL
    # the side walls lose heat: colder than the linear conduction profile there
    lin = 300 + 60 * c[:, 1]
    side = (c[:, 0] == 0) & (c[:, 1] > 0.2) & (c[:, 1] < 0.8)
    assert np.all(T.values[side] < lin[side])
    assert np.allclose(solver.radiation_flux(300.0), 0.9 * 5.670367e-8 * (280.0 ** 4 - 300.0 ** 4))


def test_radiation_transient_steps_match_oracle():
    """Radiation inside the Crank-Nicolson loop: every step is a Newton solve (the nonlinear term is fully implicit,
    ScalarTransportSolver.py:359)."""
    n, nsteps = 8, 3
    k, c_ = 0.6, 1000 * 4200.0
    dt = c_ / (n * n) / k
    settings, mesh = radiation_settings(n, {'transient': True, 'starting_time': 0.0, 'time_step': dt, 'ending_time': dt * (nsteps - 0.5)})
    solver = ScalarTransportSolver.ScalarTransportSolver(settings)
    solver.material['conductivity'] = k
    T = solver.solve()
    assert solver.current_step == nsteps
    c, t = fo.unit_square_mesh(n, n)
    nv = c.shape[0]
    fv, _, _ = fo.exterior_facets(t)
    tp, bt = np.nonzero(c[:, 1] == 1)[0], np.nonzero(c[:, 1] == 0)[0]
    dofs = np.concatenate([tp, bt]); vals = np.concatenate([np.full(tp.size, 360.0), np.full(bt.size, 300.0)])
    K = fo.assemble_matrix(t, fo.local_laplace(c, t, k), nv)
    M = fo.assemble_matrix(t, fo.local_mass(c, t, c_), nv)
    m_ = 0.9 * 5.670367e-8
    Tn = np.full(nv, 300.0)
    for _ in range(nsteps):
        Alin = (M / dt + 0.5 * K).tocsr()
        blin = (M / dt) @ Tn - 0.5 * (K @ Tn)
        X = Tn.copy()
        X[dofs] = vals
        for it in range(30):
            Jf, rf = fo.radiation_terms(c, fv, X, m_, 280.0)
            res = Alin @ X - blin
            np.add.at(res, fv.ravel(), rf.ravel())
            A, rhs = fo.apply_dirichlet((Alin + fo._scatter(fv, Jf, nv)).tocsr(), -res, dofs, np.zeros(dofs.size), symmetric=True)
            if it and np.linalg.norm(rhs) < 1e-13 * r0:
                break
            r0 = np.linalg.norm(rhs) if it == 0 else r0
            X = X + fo.solve_direct(A, rhs)
        Tn = X
    assert fo.relative_l2(T.values, Tn) < TOL


def test_point_source_matches_oracle():
    """settings['point_source'] as a list of (point, magnitude) (ScalarTransportSolver.py:150-158): delta loads through
    the basis functions of the containing cell, Dirichlet rows imposed afterwards."""
    n = 12
    settings, mesh = radiation_settings(n)
    del settings['radiation_settings']
    pts = [((0.31, 0.42), 50.0), (Point(0.75, 0.5), -20.0)]
    settings['point_source'] = pts
    solver = ScalarTransportSolver.ScalarTransportSolver(settings)
    solver.material['conductivity'] = 0.6
    T = solver.solve()
    c, t = fo.unit_square_mesh(n, n)
    nv = c.shape[0]
    b = np.zeros(nv)
    vol, G = fo.p1_geometry(c, t)
    for p, mag in pts:
        x = np.asarray(p.x if isinstance(p, Point) else p)
        lam0 = 1.0 + np.einsum("cai,ci->ca", G, x[None, :] - c[t[:, 0]])       # l_a(x) = l_a(x_0) + G_a.(x - x_0)
        lam0[:, 1:] -= 1.0
        cell = int(np.nonzero(lam0.min(axis=1) >= -1e-12)[0][0])
        np.add.at(b, t[cell], mag * lam0[cell])
    tp, bt = np.nonzero(c[:, 1] == 1)[0], np.nonzero(c[:, 1] == 0)[0]
    A = fo.conform(fo.assemble_matrix(t, fo.local_laplace(c, t, 0.6), nv), *fo.csr_pattern(t, nv))
    Ab, bb = fo.apply_dirichlet(A, b, np.concatenate([tp, bt]), np.concatenate([np.full(tp.size, 360.0), np.full(bt.size, 300.0)]), symmetric=True)
    assert fo.relative_l2(T.values, fo.solve_direct(Ab, bb)) < TOL
    assert abs(b.sum() - 30.0) < 1e-12
    # a PointSource object is accepted as well
    settings2, _ = radiation_settings(n)
    del settings2['radiation_settings']
    settings2['point_source'] = PointSource(settings2['function_space'], Point(0.31, 0.42), 50.0)
    s2 = ScalarTransportSolver.ScalarTransportSolver(settings2)
    s2.material['conductivity'] = 0.6
    assert np.isfinite(s2.solve().values).all()


# ----------------------------------------------------------------------------------- test_linear_elasticity.py:35-141
xmax = 10.0


class Left(SubDomain):
    def inside(self, x, on_boundary):
        return near(x[0], 0.0)


class Right(SubDomain):
    def inside(self, x, on_boundary):
        return near(x[0], xmax)


def elasticity_settings(fe_degree, n, has_thermal_stress, has_body_source, transient):
    mesh = BoxMesh(Point(0, 0, 0), Point(xmax, 1, 1), *n)
    rho = 7800
    bf = Expression(("10*rho", "0", "0.0"), omega=100, rho=rho, degree=2)
    from collections import OrderedDict
    bcs = OrderedDict()
    bcs["fixed"] = {'boundary': Left(), 'boundary_id': 1, 'type': 'Dirichlet', 'value': (Constant(0), None, None)}
    bcs["displ"] = {'boundary': Right(), 'boundary_id': 2, 'type': 'Dirichlet', 'value': Constant((0, 0, 1e-3))}
    V = VectorFunctionSpace(mesh, "Lagrange", fe_degree)
    s = copy.deepcopy(SolverBase.default_case_settings)
    s['material'] = {'name': 'steel', 'elastic_modulus': 2e11, 'poisson_ratio': 0.27, 'density': 7800, 'thermal_expansion_coefficient': 2e-6}
    s['function_space'] = V
    s['boundary_conditions'] = bcs
    s['temperature_distribution'] = None
    s['solver_settings']['reference_values'] = {'temperature': 293}
    s['report_settings'] = QUIET
    if transient:
        dt, f = 0.001, 100
        s['solver_settings']['transient_settings'] = {'transient': True, 'starting_time': 0.0, 'time_step': dt, 'ending_time': 0.005}
        dynamic_stress = lambda t: Constant((1e8 * math.sin(f * math.pi * 2 * t), 0, 0))       # noqa: E731
        bcs["tensile"] = {'boundary': Right(), 'boundary_id': 2, 'type': 'stress', 'value': dynamic_stress}
    if has_thermal_stress:
        s['temperature_distribution'] = Expression("343", degree=fe_degree)
    if has_body_source:
        s['body_source'] = bf
    return s, mesh


def oracle_elasticity(fe_degree, n, has_thermal_stress, has_body_source, T_nodal=None, stress=None):
    c, t = fo.box_mesh((0, 0, 0), (xmax, 1, 1), *n)
    mu, lam = fo.lame(2e11, 0.27)
    beta = 2e11 / (1 - 2 * 0.27) * 2e-6
    fv, opp, _ = fo.exterior_facets(t)
    rsel = c[fv].mean(axis=1)[:, 0] == xmax
    if fe_degree == 1:
        nn, xn = c.shape[0], c
        A = fo.assemble_matrix(t, fo.local_elasticity(c, t, mu, lam), nn, 3)
        b = np.zeros(3 * nn)
        if has_thermal_stress:
            b += fo.thermal_load(c, t, beta, 343.0 if T_nodal is None else T_nodal(xn), 293.0)
        if has_body_source:
            b -= fo.assemble_source(c, t, np.array([10 * 7800.0, 0, 0]), ncomp=3)
        if stress is not None:
            b -= fo.assemble_facet_load(c, fv[rsel], np.asarray(stress), nn, 3)
    else:
        cn, xn, edges = fp.p2_dofmap(c, t)
        nn = xn.shape[0]
        A = fp.assemble_matrix(cn, fp.local_elasticity(c, t, mu, lam), nn, 3)
        b = np.zeros(3 * nn)
        if has_thermal_stress:
            b += fp.thermal_load(c, t, cn, nn, beta, 343.0 if T_nodal is None else T_nodal(xn), 293.0)
        if has_body_source:
            b -= fp.assemble_source(c, t, cn, nn, np.array([10 * 7800.0, 0, 0]), ncomp=3)
        if stress is not None:
            b -= fp.assemble_facet_load(c, fv[rsel], fp.facet_nodes(fv[rsel], edges, c.shape[0]), np.asarray(stress), nn, 3)
    lv, rv = np.nonzero(xn[:, 0] == 0)[0], np.nonzero(xn[:, 0] == xmax)[0]
    dofs = np.concatenate([lv * 3, (rv[:, None] * 3 + np.arange(3)).ravel()])
    vals = np.concatenate([np.zeros(lv.size), np.tile([0, 0, 1e-3], rv.size)])
    Ab, bb = fo.apply_dirichlet(A, b, dofs, vals, symmetric=True)
    return fo.solve_direct(Ab, bb), c, t, xn


@pytest.mark.parametrize("fe_degree,n", [(1, (12, 3, 3)), (2, (6, 2, 2))])
@pytest.mark.parametrize("has_thermal_stress,has_body_source", [(True, True), (True, False)])
def test_linear_elasticity_thermal_stress(fe_degree, n, has_thermal_stress, has_body_source):
    """test(has_thermal_stress=True, ...) of the reference example (:163-167): Expression("343") against the reference
    temperature 293, body force as an Expression, on P1 and on the example's own degree-2 space."""
    s, mesh = elasticity_settings(fe_degree, n, has_thermal_stress, has_body_source, False)
    solver = LinearElasticitySolver.LinearElasticitySolver(s)
    u = solver.solve()
    assert solver.solve_info["converged"] == 1
    uo, c, t, xn = oracle_elasticity(fe_degree, n, has_thermal_stress, has_body_source)
    assert fo.relative_l2(u.vector().get_local(), uo) < 1e-9
    assert abs(solver.thermal_stress(343.0) - 2e11 / (1 - 0.54) * 2e-6 * 50) < 1e-3
    # a free bar heated uniformly expands stress-free: thermal load alone, clamp only what removes rigid modes
    if fe_degree == 1 and not has_body_source:
        vm = solver.von_Mises(u)
        vo = fo.von_mises_projection(c, t, uo, *fo.lame(2e11, 0.27))
        assert fo.relative_l2(vm.values, vo) < 1e-8
        assert vm.values.shape == (c.shape[0],)


def test_von_mises_of_a_degree_2_displacement():
    s, mesh = elasticity_settings(2, (4, 2, 2), True, True, False)
    solver = LinearElasticitySolver.LinearElasticitySolver(s)
    u = solver.solve()
    vm = solver.von_Mises(u)
    c, t = fo.box_mesh((0, 0, 0), (xmax, 1, 1), 4, 2, 2)
    cn, xn, _ = fp.p2_dofmap(c, t)
    nv = c.shape[0]
    b = fp.von_mises_load(c, t, cn, u.vector().get_local(), *fo.lame(2e11, 0.27))
    M = fo.assemble_matrix(t, fo.local_mass(c, t, 1.0), nv)
    assert fo.relative_l2(vm.values, fo.solve_direct(M, b)) < 1e-8
    assert vm.values.min() > 0


def test_nodal_temperature_distribution_thermal_stress():
    """A temperature field (the commented-out variant at :126, "dT * x[1]/ymax") instead of the constant."""
    n = (8, 3, 3)
    s, mesh = elasticity_settings(1, n, False, False, False)
    s['temperature_distribution'] = Expression("293 + dT * x[1]/ymax", dT=100, ymax=1.0, degree=1)
    solver = LinearElasticitySolver.LinearElasticitySolver(s)
    u = solver.solve()
    uo, *_ = oracle_elasticity(1, n, True, False, T_nodal=lambda x: 293 + 100 * x[:, 1])
    assert fo.relative_l2(u.vector().get_local(), uo) < 1e-9


def test_linear_elasticity_transient_dynamic_stress(tmp_path):
    """test(..., transient=True) (:117-121, :164): five quasi-static steps; the stress on the right face follows
    sin(2 pi f t) evaluated at the reference's one-step-behind time (SolverBase.py:453-465); the Dirichlet condition
    with the same boundary id keeps its rows.  The last step is checked against the oracle, every step is saved."""
    n = (8, 2, 2)
    s, mesh = elasticity_settings(1, n, True, True, True)
    pvd = os.path.join(str(tmp_path), "elastic_displacement.pvd")
    s['report_settings'] = dict(QUIET, saving_freq=1, result_filename=pvd)
    solver = LinearElasticitySolver.LinearElasticitySolver(s)
    u = solver.solve()
    nsteps = solver.current_step
    assert nsteps >= 5
    t_last = solver.get_current_time(nsteps - 1) if nsteps - 1 else solver.get_current_time(0)
    stress = (1e8 * math.sin(100 * math.pi * 2 * t_last), 0, 0)
    uo, c, t, xn = oracle_elasticity(1, n, True, True, stress=stress)
    assert fo.relative_l2(u.vector().get_local(), uo) < 1e-9
    root = ET.parse(pvd).getroot()
    pieces = root.findall("./Collection/DataSet")
    assert len(pieces) == nsteps - 1                     # step 0 is not saved (current_step > 0, SolverBase.py:531)
    vtu = ET.parse(os.path.join(str(tmp_path), pieces[-1].attrib["file"])).getroot()
    piece = vtu.find("./UnstructuredGrid/Piece")
    assert int(piece.attrib["NumberOfPoints"]) == c.shape[0] and int(piece.attrib["NumberOfCells"]) == t.shape[0]
    arr = np.array(piece.find("./PointData/DataArray").text.split(), dtype=np.float64).reshape(-1, 3)
    assert np.abs(arr - u.values).max() <= 1e-14 * np.abs(u.values).max()


@pytest.mark.parametrize("dim", [2, 3])
def test_nonlinear_conductivity_terms(ctx, dim):
    c, t = small_mesh(dim)
    nv = c.shape[0]
    rng = np.random.default_rng(3)
    T = 300 + 60 * rng.random(nv)
    kf, dkf = (lambda x: 0.6 * (1 + 0.02 * (x - 300.0)) + 1e-5 * x ** 2), (lambda x: 0.6 * 0.02 + 2e-5 * x)
    m = _lib.DeviceMesh.upload(ctx, c, t)
    A = _lib.DeviceMatrix.create(m, 1)
    r = _lib.DeviceVector(ctx, nv)
    V = lambda a: _lib.DeviceVector.from_numpy(ctx, a)       # noqa: E731
    _lib.assemble_scalar_nonlinear_k(m, A, r, V(T), V(kf(T)), V(dkf(T)), rscale=-1.0)
    J, R = fo.nonlinear_k_terms(c, t, T, kf, dkf)
    close(r.numpy(), -R, 1e-12)
    _, _, va = A.download_csr()
    close(va, fo.conform(J, *fo.csr_pattern(t, nv)).data, 1e-12)


def test_nonlinear_conductivity_example_matches_oracle_newton():
    """examples/test_heat_transfer.py:53-56 with `nonlinear = True`: conductivity = lambda T: (T-T_ambient)/T_ambient * 0.6,
    shifted so that k stays positive on [300, 360]."""
    n = 16
    settings, mesh = radiation_settings(n)
    del settings['radiation_settings']
    kfun = lambda T: (0.2 + (T - 300) / 300) * 0.6       # noqa: E731
    solver = ScalarTransportSolver.ScalarTransportSolver(settings)
    solver.material['conductivity'] = kfun
    T = solver.solve()
    info = solver.solve_info
    assert info['converged'] == 1 and info['newton_iterations'] <= 10
    c, t = fo.unit_square_mesh(n, n)
    tp, bt = np.nonzero(c[:, 1] == 1)[0], np.nonzero(c[:, 1] == 0)[0]
    To, _ = fo.solve_nonlinear_k_newton(c, t, kfun, lambda T: 0.6 / 300 + 0 * T, [(tp, 360.0), (bt, 300.0)], 300.0)
    assert fo.relative_l2(T.values, To) < TOL
    # Kirchhoff profile (second-order accurate)
    q = c[:, 1] * (0.12 * 60 + 0.001 * 3600)
    s = (-0.12 + np.sqrt(0.12 ** 2 + 2 * 0.002 * q)) / 0.002
    assert np.abs(T.values - 300 - s).max() < 0.2


# ----------------------------------------------------------------------------------- SUPG (ScalarTransportSolver.py:252-274)
@pytest.mark.parametrize("dim", [2, 3])
def test_supg_extra_terms(ctx, dim):
    c, t = small_mesh(dim)
    nv = c.shape[0]
    vel = np.array([0.7, -0.3, 0.2][:dim])
    Pe = 5.0
    m = _lib.DeviceMesh.upload(ctx, c, t)
    A = _lib.DeviceMatrix.create(m, 1)
    _lib.assemble_scalar_supg(m, A, vel, Pe, mass=2.5, adv=1.5)
    _, _, va = A.download_csr()
    ref = fo.conform(fo.assemble_matrix(t, fo.local_supg(c, t, vel, Pe, mass=2.5, adv=1.5), nv), *fo.csr_pattern(t, nv))
    close(va, ref.data)
    x = np.random.default_rng(1).standard_normal(nv)
    y = _lib.DeviceVector(ctx, nv)
    _lib.assemble_scalar_supg(m, None, vel, Pe, mass=2.5, adv=1.5, x=_lib.DeviceVector.from_numpy(ctx, x), y=y)
    close(y.numpy(), ref @ x, 1e-12)
    b = _lib.DeviceVector(ctx, nv)
    _lib.assemble_source_supg(m, b, 4.0, vel, Pe)
    close(b.numpy(), fo.supg_source(c, t, 4.0, vel, Pe))
    fv, opp, _ = fo.exterior_facets(t)
    A.zero()
    b2 = _lib.DeviceVector(ctx, nv)
    _lib.assemble_facet_supg(m, A, b2, fv, opp, vel, Pe, g=2.0, h=3.0)
    Af, bf = fo.supg_facet_terms(c, fv, opp, vel, Pe, g=2.0, h=3.0)
    close(b2.numpy(), bf, 1e-12)
    _, _, va = A.download_csr()
    close(va, fo.conform(Af, *fo.csr_pattern(t, nv)).data, 1e-12)


def supg_case(n, transient=None):
    mesh = UnitSquareMesh(n, n)
    Q = FunctionSpace(mesh, "CG", 1)
    bcs = {"hot": {'boundary': top, 'boundary_id': 1, 'type': 'Dirichlet', 'value': Constant(360)},
           "cold": {'boundary': bottom, 'boundary_id': 2, 'type': 'HTC', 'value': Constant(100), 'ambient': Constant(300)},
           "left": {'boundary': left, 'boundary_id': 3, 'type': 'heatFlux', 'value': Constant(40.0)},
           "right": {'boundary': right, 'boundary_id': 4, 'type': 'symmetry', 'value': None}}
    return {'solver_name': 'ScalarTransportSolver', 'mesh': None, 'function_space': Q, 'periodic_boundary': None, 'fe_degree': 1,
            'boundary_conditions': bcs, 'body_source': 2000.0, 'initial_values': {'temperature': 300},
            'material': {'density': 1000, 'specific_heat_capacity': 4200, 'thermal_conductivity': 0.6},
            'convective_velocity': Constant((0.5e-5, -2e-5)),
            'advection_settings': {'stabilization_method': 'SPUG', 'Pe': 50.0},
            'solver_settings': {'transient_settings': transient or {'transient': False, 'starting_time': 0, 'time_step': 0.1, 'ending_time': 1},
                                'reference_values': {'temperature': 300}, 'solver_parameters': {}},
            'scalar_name': 'temperature', 'report_settings': QUIET}


def supg_oracle_system(n, dt=None, Tn=None):
    c, t = fo.unit_square_mesh(n, n)
    nv = c.shape[0]
    k, cap, vel, Pe = 0.6, 1000 * 4200.0, np.array([0.5e-5, -2e-5]), 50.0
    fv, opp, _ = fo.exterior_facets(t)
    mid = c[fv].mean(axis=1)
    fb, fl = mid[:, 1] == 0, mid[:, 0] == 0
    K = fo.assemble_matrix(t, fo.local_laplace(c, t, k), nv)
    C = fo.assemble_matrix(t, fo.local_advection(c, t, vel, cap) + fo.local_supg(c, t, vel, Pe, adv=cap), nv)
    Ah, bh = fo.supg_facet_terms(c, fv[fb], opp[fb], vel, Pe, g=100.0 * 300.0, h=100.0)
    _, bl = fo.supg_facet_terms(c, fv[fl], opp[fl], vel, Pe, g=40.0)
    R = fo._scatter(fv[fb], fo.local_facet_mass(c, fv[fb], 100.0), nv) + Ah
    loads = (fo.assemble_source(c, t, 2000.0) + fo.supg_source(c, t, 2000.0, vel, Pe) + fo.assemble_facet_load(c, fv[fb], 100.0 * 300.0, nv)
             + bh + fo.assemble_facet_load(c, fv[fl], 40.0, nv) + bl)
    tp = np.nonzero(c[:, 1] == 1)[0]
    if dt is None:
        A, b = K + C + R, loads
    else:
        M = fo.assemble_matrix(t, fo.local_mass(c, t, cap / dt) + fo.local_supg(c, t, vel, Pe, mass=cap / dt), nv)
        A, b = M + 0.5 * K + C + R, M @ Tn - 0.5 * (K @ Tn) + loads
    Ab, bb = fo.apply_dirichlet(A.tocsr(), b, tp, np.full(tp.size, 360.0), symmetric=False)
    return fo.solve_direct(Ab, bb)


def test_supg_steady_with_htc_flux_and_source_matches_oracle():
    """using_convective_velocity + HTC of examples/test_heat_transfer.py:139-160 with advection_settings
    {'stabilization_method': 'SPUG', 'Pe': ...}: every integral carries the stabilised test function."""
    n = 16
    solver = ScalarTransportSolver.ScalarTransportSolver(supg_case(n))
    T = solver.solve()
    assert solver.solve_info['converged'] == 1
    To = supg_oracle_system(n)
    assert fo.relative_l2(T.values, To) < TOL
    plain = supg_case(n)
    plain['advection_settings'] = {'stabilization_method': None}
    Tg = ScalarTransportSolver.ScalarTransportSolver(plain).solve()
    assert fo.relative_l2(Tg.values, To) > 1e-6          # the stabilisation does change the discrete answer
    bad = supg_case(n)
    bad['advection_settings'] = {'stabilization_method': 'IP', 'alpha': 0.1}
    with pytest.raises(SolverBase.SolverError):
        ScalarTransportSolver.ScalarTransportSolver(bad).solve()


def test_supg_transient_matches_oracle():
    n, nsteps = 8, 3
    dt = 1000 * 4200.0 / (n * n) / 0.6
    s = supg_case(n, {'transient': True, 'starting_time': 0.0, 'time_step': dt, 'ending_time': dt * (nsteps - 0.5)})
    solver = ScalarTransportSolver.ScalarTransportSolver(s)
    T = solver.solve()
    assert solver.current_step == nsteps
    Tn = np.full((n + 1) ** 2, 300.0)
    for _ in range(nsteps):
        Tn = supg_oracle_system(n, dt, Tn)
    assert fo.relative_l2(T.values, Tn) < TOL


def test_new_entry_points_empty_inputs_and_errors(ctx):
    """Empty lists are no-ops, wrong sizes / degrees / parameters come back as SolverError with a message."""
    c, t = small_mesh(3)
    nv = c.shape[0]
    m = _lib.DeviceMesh.upload(ctx, c, t)
    A = _lib.DeviceMatrix.create(m, 1)
    b, T = _lib.DeviceVector(ctx, nv), _lib.DeviceVector(ctx, nv)
    b3 = _lib.DeviceVector(ctx, 3 * nv)
    empty_f = np.zeros((0, 3), dtype=np.int32)
    _lib.assemble_facet_radiation(m, A, b, T, empty_f, 5e-8, 300.0)                       # nf = 0
    _lib.assemble_facet_supg(m, A, b, empty_f, np.zeros(0, dtype=np.int32), [1.0, 0, 0], 10.0, g=1.0, h=1.0)
    b.add_entries(np.zeros(0, dtype=np.int64), np.zeros(0))
    assert np.all(b.numpy() == 0.0) and np.all(A.download_csr()[2] == 0.0)
    with pytest.raises(_lib.SolverError):
        b.add_entries([nv], [1.0])                                                       # index out of range
    with pytest.raises(_lib.SolverError):
        _lib.assemble_thermal_load(m, b, 1.0, T_const=1.0)                               # rhs must have ncomp == dim
    with pytest.raises(_lib.SolverError):
        _lib.assemble_thermal_load(m, b3, 1.0, T=b3)                                     # temperature must be nodal scalar
    with pytest.raises(_lib.SolverError):
        _lib.assemble_von_mises_load(m, b, 1.0, 1.0, b)                                  # displacement size
    with pytest.raises(_lib.SolverError):
        _lib.assemble_scalar_supg(m, A, [1.0, 0, 0], 0.0, adv=1.0)                       # Peclet must be positive
    with pytest.raises(_lib.SolverError):
        _lib.assemble_scalar_nonlinear_k(m, A, b, T, b3, T)                              # nodal vector sizes
    cn, xn, _ = fp.p2_dofmap(c, t)
    m2 = _lib.DeviceMesh.upload_p2(ctx, c, cn, xn.shape[0])
    A2 = _lib.DeviceMatrix.create(m2, 1)
    v2 = _lib.DeviceVector(ctx, xn.shape[0])
    fv, _, _ = fo.exterior_facets(t)
    with pytest.raises(_lib.SolverError):
        _lib.assemble_facet_radiation(m2, A2, v2, v2, fv, 5e-8, 300.0)                   # degree-1 only
    with pytest.raises(_lib.SolverError):
        _lib.assemble_scalar_supg(m2, A2, [1.0, 0, 0], 5.0, adv=1.0)
    with pytest.raises(_lib.SolverError):
        ctx.dist_set_halo(1, 2, [1], [0, 1], [0], [1], [1])                              # no fsb_dist_init
    with pytest.raises(_lib.SolverError):
        _lib.Multigrid(ctx, [A], [(3, 3, 3)], 3)                                         # matrix does not match the box
    # a single-level hierarchy is legal: the "V-cycle" is the coarse smoother
    mb = _lib.DeviceMesh.box(ctx, (2, 2, 2), (0, 0, 0), (1, 1, 1))
    Ab = _lib.DeviceMatrix.create(mb, 1)
    Ab.assemble_scalar(kscale=1.0, mass=1.0)
    mg = _lib.Multigrid(ctx, [Ab], [(2, 2, 2)], 3)
    rhs, x = _lib.DeviceVector(ctx, 27), _lib.DeviceVector(ctx, 27)
    rhs.fill(1.0)
    info = mg.solve(rhs, x, rtol=1e-12, maxit=200)
    assert info["converged"] == 1
    rp, ci, va = Ab.download_csr()
    M = sp.csr_matrix((va, ci.astype(np.int64), rp), shape=(27, 27))
    assert np.abs(M @ x.numpy() - 1.0).max() < 1e-9


def test_elasticity_surface_source_is_a_normal_load_on_the_whole_surface():
    """settings['surface_source'] = {'value': p} (LinearElasticitySolver.py:110-115): dot(mesh_normal*p, v)*ds over every
    exterior facet, with the reference's load sign."""
    n = (6, 3, 3)
    s, mesh = elasticity_settings(1, n, False, False, False)
    s['surface_source'] = {'value': 2.5e6, 'direction': None}
    solver = LinearElasticitySolver.LinearElasticitySolver(s)
    u = solver.solve()
    c, t = fo.box_mesh((0, 0, 0), (xmax, 1, 1), *n)
    nv = c.shape[0]
    mu, lam = fo.lame(2e11, 0.27)
    A = fo.assemble_matrix(t, fo.local_elasticity(c, t, mu, lam), nv, 3)
    fv, opp, _ = fo.exterior_facets(t)
    meas, nrm = fo.facet_measure(c, fv, opp)
    b = -fo.assemble_facet_load(c, fv, 2.5e6 * nrm, nv, 3)
    lv, rv = np.nonzero(c[:, 0] == 0)[0], np.nonzero(c[:, 0] == xmax)[0]
    dofs = np.concatenate([lv * 3, (rv[:, None] * 3 + np.arange(3)).ravel()])
    vals = np.concatenate([np.zeros(lv.size), np.tile([0, 0, 1e-3], rv.size)])
    Ab, bb = fo.apply_dirichlet(A, b, dofs, vals, symmetric=True)
    assert fo.relative_l2(u.vector().get_local(), fo.solve_direct(Ab, bb)) < 1e-9


@pytest.mark.parametrize("dim", [2, 3])
def test_advection_by_a_velocity_field(ctx, dim):
    """fsb_assemble_advection_nodal against the oracle (matrix and matrix-free action), then through the solver with an
    Expression velocity (a rotating flow), ScalarTransportSolver.py:130-139."""
    c, t = small_mesh(dim)
    nv = c.shape[0]
    vel = np.stack([-c[:, 1], c[:, 0]] + ([0.3 * c[:, 2] + c[:, 0]] if dim == 3 else []), axis=1)
    m = _lib.DeviceMesh.upload(ctx, c, t)
    A = _lib.DeviceMatrix.create(m, 1)
    vd = _lib.DeviceVector.from_numpy(ctx, vel.ravel())
    _lib.assemble_advection_nodal(m, A, vd, scale=2.5)
    ref = fo.conform(fo.assemble_matrix(t, fo.local_advection_nodal(c, t, vel, 2.5), nv), *fo.csr_pattern(t, nv))
    _, _, va = A.download_csr()
    close(va, ref.data)
    x = np.random.default_rng(4).standard_normal(nv)
    y = _lib.DeviceVector(ctx, nv)
    _lib.assemble_advection_nodal(m, None, vd, scale=2.5, x=_lib.DeviceVector.from_numpy(ctx, x), y=y)
    close(y.numpy(), ref @ x, 1e-12)
    # a constant field reproduces the constant-velocity kernel
    A.zero()
    _lib.assemble_advection_nodal(m, A, _lib.DeviceVector.from_numpy(ctx, np.tile([0.4, -0.7, 0.2][:dim], nv)), scale=1.0)
    B = _lib.DeviceMatrix.create(m, 1)
    B.assemble_scalar(kscale=0.0, adv=1.0, vel=np.array([0.4, -0.7, 0.2][:dim]))
    close(A.download_csr()[2], B.download_csr()[2], 1e-12)


def test_rotating_flow_expression_velocity_matches_oracle():
    n = 16
    settings, mesh = radiation_settings(n)
    del settings['radiation_settings']
    settings['convective_velocity'] = Expression(("-w*(x[1]-0.5)", "w*(x[0]-0.5)"), w=2e-6, degree=1)
    solver = ScalarTransportSolver.ScalarTransportSolver(settings)
    solver.material['conductivity'] = 0.6
    T = solver.solve()
    assert solver.solve_info['converged'] == 1
    c, t = fo.unit_square_mesh(n, n)
    nv = c.shape[0]
    vel = np.stack([-2e-6 * (c[:, 1] - 0.5), 2e-6 * (c[:, 0] - 0.5)], axis=1)
    cap = 1000 * 4200.0
    A = fo.assemble_matrix(t, fo.local_laplace(c, t, 0.6) + fo.local_advection_nodal(c, t, vel, cap), nv)
    tp, bt = np.nonzero(c[:, 1] == 1)[0], np.nonzero(c[:, 1] == 0)[0]
    Ab, bb = fo.apply_dirichlet(fo.conform(A, *fo.csr_pattern(t, nv)), np.zeros(nv), np.concatenate([tp, bt]),
                                np.concatenate([np.full(tp.size, 360.0), np.full(bt.size, 300.0)]), symmetric=False)
    To = fo.solve_direct(Ab, bb)
    assert fo.relative_l2(T.values, To) < TOL
    assert np.abs(To - (300 + 60 * c[:, 1])).max() > 1.0         # the swirl does bend the isotherms


def test_elasticity_stress_tensor_boundary():
    """bcs['tensile'] = {'type': 'stress', 'value': Constant(((..),(..),(..)))}: the traction is S.n per facet
    (LinearElasticitySolver.py:190-196), with the reference's load sign; degree 1 and the example's degree 2."""
    S = np.array([[1e8, 2e7, 0.0], [2e7, 0.0, 0.0], [0.0, 0.0, 5e6]])
    for fe_degree, n in ((1, (8, 3, 3)), (2, (4, 2, 2))):
        s, mesh = elasticity_settings(fe_degree, n, False, False, False)
        s['boundary_conditions']["fixed"]['value'] = Constant((0, 0, 0))
        s['boundary_conditions']["displ"] = {'boundary': Right(), 'boundary_id': 2, 'type': 'stress', 'value': Constant(S)}
        solver = LinearElasticitySolver.LinearElasticitySolver(s)
        u = solver.solve()
        assert solver.solve_info['converged'] == 1
        c, t = fo.box_mesh((0, 0, 0), (xmax, 1, 1), *n)
        mu, lam = fo.lame(2e11, 0.27)
        fv, opp, _ = fo.exterior_facets(t)
        rsel = c[fv].mean(axis=1)[:, 0] == xmax
        meas, nrm = fo.facet_measure(c, fv[rsel], opp[rsel])
        if fe_degree == 1:
            nn, xn = c.shape[0], c
            A = fo.assemble_matrix(t, fo.local_elasticity(c, t, mu, lam), nn, 3)
            b = -fo.assemble_facet_load(c, fv[rsel], nrm @ S.T, nn, 3)
        else:
            cn, xn, edges = fp.p2_dofmap(c, t)
            nn = xn.shape[0]
            A = fp.assemble_matrix(cn, fp.local_elasticity(c, t, mu, lam), nn, 3)
            b = -fp.assemble_facet_load(c, fv[rsel], fp.facet_nodes(fv[rsel], edges, c.shape[0]), nrm @ S.T, nn, 3)
        lv = np.nonzero(xn[:, 0] == 0)[0]
        dofs = (lv[:, None] * 3 + np.arange(3)).ravel()
        Ab, bb = fo.apply_dirichlet(A, b, dofs, np.zeros(dofs.size), symmetric=True)
        assert fo.relative_l2(u.vector().get_local(), fo.solve_direct(Ab, bb)) < 1e-9


def test_conductivity_field_expression():
    """material['conductivity'] = Expression(...) : a scalar field k(x), P1-interpolated (the reference's tensor-weighted example
    passes an Expression too, examples/test_heat_transfer.py:91); steady and Crank-Nicolson."""
    n = 12
    kexpr = "0.6 * (1 + x[0] + 2*x[1]*x[1])"
    c, t = fo.unit_square_mesh(n, n)
    nv = c.shape[0]
    kn = 0.6 * (1 + c[:, 0] + 2 * c[:, 1] ** 2)
    kcell = kn[t].mean(axis=1)
    tp, bt = np.nonzero(c[:, 1] == 1)[0], np.nonzero(c[:, 1] == 0)[0]
    dofs = np.concatenate([tp, bt]); vals = np.concatenate([np.full(tp.size, 360.0), np.full(bt.size, 300.0)])
    K = fo.conform(fo.assemble_matrix(t, fo.local_laplace(c, t, kcell), nv), *fo.csr_pattern(t, nv))
    settings, mesh = radiation_settings(n)
    del settings['radiation_settings']
    settings['body_source'] = 250.0
    solver = ScalarTransportSolver.ScalarTransportSolver(settings)
    solver.material['conductivity'] = Expression(kexpr, degree=1)
    T = solver.solve()
    assert not solver.nonlinear and solver.solve_info['converged'] == 1
    Ab, bb = fo.apply_dirichlet(K, fo.assemble_source(c, t, 250.0), dofs, vals, symmetric=True)
    assert fo.relative_l2(T.values, fo.solve_direct(Ab, bb)) < TOL
    # transient: theta K(k) on the left, -(1 - theta) K(k) T_prev on the right
    nsteps, cap = 3, 1000 * 4200.0
    dt = cap / (n * n) / 0.6
    settings, mesh = radiation_settings(n, {'transient': True, 'starting_time': 0.0, 'time_step': dt, 'ending_time': dt * (nsteps - 0.5)})
    del settings['radiation_settings']
    solver = ScalarTransportSolver.ScalarTransportSolver(settings)
    solver.material['conductivity'] = Expression(kexpr, degree=1)
    T = solver.solve()
    M = fo.assemble_matrix(t, fo.local_mass(c, t, cap), nv)
    Tn = np.full(nv, 300.0)
    for _ in range(nsteps):
        Ab, bb = fo.apply_dirichlet((M / dt + 0.5 * K).tocsr(), (M / dt) @ Tn - 0.5 * (K @ Tn), dofs, vals, symmetric=True)
        Tn = fo.solve_direct(Ab, bb)
    assert fo.relative_l2(T.values, Tn) < TOL
