"""GPU parity tests for the degree-2 (P2) path, kernel level through the C-ABI, against oracle/fem_oracle_p2.py
(reference tensors by exact barycentric integration, an independent route from the CUDA tables).
Bars: pattern bit-exact; values 1e-13 of the matrix scale; solutions 1e-10 relative L2."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu

from oracle import fem_oracle as fo  # noqa: E402
from oracle import fem_oracle_p2 as p2  # noqa: E402
from fenicssolver_b200 import _lib  # noqa: E402


def mesh_case(dim, seed=0):
    n = (5, 4) if dim == 2 else (3, 3, 2)
    c0, t = (fo.rectangle_mesh(0, 0, 2, 1, *n) if dim == 2 else fo.box_mesh((0, 0, 0), (2, 1, 1), *n))
    rng = np.random.default_rng(seed)
    c = c0 + 0.12 / max(n) * (rng.random(c0.shape) * 2 - 1)
    hi = [2, 1, 1][:dim]
    bnd = np.any((c0 == 0) | (c0 == np.array(hi)), axis=1)
    c[bnd] = c0[bnd]
    cn, xc, edges = p2.p2_dofmap(c, t)
    return c, t, cn, xc, edges


def device_csr(A):
    rp, ci, va = A.download_csr()
    n = rp.size - 1
    return sp.csr_matrix((va, ci.astype(np.int64), rp), shape=(n, n)), rp, ci


def close(dev, ref, tol=1e-13):
    assert dev.shape == ref.shape
    assert np.abs(dev - ref).max() <= tol * np.abs(ref).max()


@pytest.mark.parametrize("asm_mode", [0, 1, 2])
@pytest.mark.parametrize("dim", [2, 3])
def test_p2_pattern_and_scalar_terms(ctx, dim, asm_mode):
    c, t, cn, xc, edges = mesh_case(dim)
    nn = xc.shape[0]
    m = _lib.DeviceMesh.upload_p2(ctx, c, cn, nn)
    rp0, ci0 = p2.csr_pattern(cn, nn)
    rng = np.random.default_rng(1)
    kt = rng.random((dim, dim)) + dim * np.eye(dim)
    vel = rng.random(dim) - 0.5
    ctx.set_option("asm_mode", asm_mode)
    try:
        for kw in (dict(kscale=3.0), dict(kscale=0.0, mass=2.5), dict(kscale=0.0, adv=1.5, vel=vel),
                   dict(kscale=0.7, ktensor=kt, mass=1.25, adv=4.0, vel=vel)):
            A = _lib.DeviceMatrix.create(m, 1)
            A.assemble_scalar(**kw)
            dev, rp, ci = device_csr(A)
            assert np.array_equal(rp, rp0) and np.array_equal(ci, ci0)
            Ke = kw.get("kscale", 1.0) * p2.local_laplace(c, t, kw.get("ktensor", 1.0))
            if kw.get("mass"):
                Ke = Ke + p2.local_mass(c, t, kw["mass"])
            if kw.get("adv"):
                Ke = Ke + p2.local_advection(c, t, kw["vel"], kw["adv"])
            ref = fo.conform(p2.assemble_matrix(cn, Ke, nn), rp0, ci0)
            close(dev.data, ref.data)
    finally:
        ctx.set_option("asm_mode", 1)
    # matrix-free action (Crank-Nicolson right-hand side)
    xh = rng.random(nn)
    x, y = _lib.DeviceVector.from_numpy(ctx, xh), _lib.DeviceVector(ctx, nn)
    _lib.apply_scalar(m, x, y, kscale=-0.3, mass=7.0)
    close(y.numpy(), p2.assemble_matrix(cn, -0.3 * p2.local_laplace(c, t) + p2.local_mass(c, t, 7.0), nn) @ xh)


@pytest.mark.parametrize("dim", [2, 3])
def test_p2_elasticity_matrix(ctx, dim):
    c, t, cn, xc, edges = mesh_case(dim, seed=2)
    nn = xc.shape[0]
    mu, lam = fo.lame(2e11, 0.27)
    m = _lib.DeviceMesh.upload_p2(ctx, c, cn, nn)
    A = _lib.DeviceMatrix.create(m, dim)
    A.assemble_elasticity(mu, lam)
    dev, rp, ci = device_csr(A)
    rp0, ci0 = p2.csr_pattern(cn, nn, dim)
    assert np.array_equal(rp, rp0) and np.array_equal(ci, ci0)
    ref = fo.conform(p2.assemble_matrix(cn, p2.local_elasticity(c, t, mu, lam), nn, dim), rp0, ci0)
    close(dev.data, ref.data)


@pytest.mark.parametrize("dim", [2, 3])
def test_p2_rhs_and_facet_terms(ctx, dim):
    c, t, cn, xc, edges = mesh_case(dim, seed=3)
    nn, nv = xc.shape[0], c.shape[0]
    m = _lib.DeviceMesh.upload_p2(ctx, c, cn, nn)
    fv, opp, _ = fo.exterior_facets(t)
    fn = p2.facet_nodes(fv, edges, nv).astype(np.int32)
    rng = np.random.default_rng(4)
    sel = rng.random(fv.shape[0]) < 0.6
    b = _lib.DeviceVector(ctx, nn)
    _lib.assemble_source(m, b, 1000.0)
    _lib.assemble_facet_load(m, b, fn[sel], 36.0)
    ref = p2.assemble_source(c, t, cn, nn, 1000.0) + p2.assemble_facet_load(c, fv[sel], fn[sel], 36.0, nn)
    close(b.numpy(), ref)
    gvec = rng.random(dim)
    bv = _lib.DeviceVector(ctx, nn * dim)
    _lib.assemble_source(m, bv, gvec, ncomp=dim, scale=-1.0)
    _lib.assemble_facet_load(m, bv, fn[sel], gvec, ncomp=dim)
    _lib.assemble_facet_load(m, bv, fn[sel], 1e6, ncomp=dim, opp=opp[sel], normal=True)
    meas, nrm = fo.facet_measure(c, fv[sel], opp[sel])
    ref = (-p2.assemble_source(c, t, cn, nn, gvec, ncomp=dim) + p2.assemble_facet_load(c, fv[sel], fn[sel], gvec, nn, dim)
           + p2.assemble_facet_load(c, fv[sel], fn[sel], 1e6 * nrm, nn, dim))
    close(bv.numpy(), ref)
    Sn = rng.random(nn)
    bn = _lib.DeviceVector(ctx, nn)
    _lib.assemble_source_nodal(m, bn, _lib.DeviceVector.from_numpy(ctx, Sn), scale=2.0)
    close(bn.numpy(), 2.0 * (p2.assemble_matrix(cn, p2.local_mass(c, t), nn) @ Sn))
    assert abs(_lib.facet_area(m, fn[sel]) - fo.boundary_area(c, fv[sel])) < 1e-12 * fo.boundary_area(c, fv[sel])
    A = _lib.DeviceMatrix.create(m, 1)
    A.assemble_facet_mass(fn[sel], 100.0)
    dev, rp, ci = device_csr(A)
    close(dev.data, fo.conform(fo._scatter(fn[sel], p2.local_facet_mass(c, fv[sel], 100.0), nn), rp, ci).data)


@pytest.mark.parametrize("dim", [2, 3])
def test_p2_poisson_quadratic_solution_is_exact(ctx, dim):
    """-div(k grad u) = f with a quadratic u: P2 reproduces it at every node (vertices and edge midpoints)."""
    c, t, cn, xc, edges = mesh_case(dim, seed=5)
    nn, nv = xc.shape[0], c.shape[0]
    rng = np.random.default_rng(6)
    Q = rng.random((dim, dim)); Q = Q + Q.T
    lin = rng.random(dim)
    u = np.einsum("ni,ij,nj->n", xc, Q, xc) + xc @ lin + 1.0
    k = 3.0
    m = _lib.DeviceMesh.upload_p2(ctx, c, cn, nn)
    A = _lib.DeviceMatrix.create(m, 1)
    A.assemble_scalar(kscale=k)
    b = _lib.DeviceVector(ctx, nn)
    _lib.assemble_source(m, b, -k * 2 * np.trace(Q))
    fv, _, _ = fo.exterior_facets(t)
    bd = np.unique(p2.facet_nodes(fv, edges, nv))
    x = _lib.DeviceVector(ctx, nn)
    A.apply_dirichlet(b, bd, u[bd], symmetric=True, x=x)
    info = A.solve(b, x, "cg", rtol=1e-13)
    assert info["converged"] == 1
    assert fo.relative_l2(x.numpy(), u) < 1e-10
    # and it agrees with the oracle's direct solve of the same system
    Ao, bo = fo.apply_dirichlet(p2.assemble_matrix(cn, p2.local_laplace(c, t, k), nn), p2.assemble_source(c, t, cn, nn, -k * 2 * np.trace(Q)), bd, u[bd], True)
    assert fo.relative_l2(x.numpy(), fo.solve_direct(Ao, bo)) < 1e-10


def test_p2_elasticity_solve_matches_oracle(ctx):
    """Cantilever-like box, P2 vector space (the reference example's element), body force + traction,
    reference load sign: against the oracle's direct solve."""
    c, t = fo.box_mesh((0, 0, 0), (4, 1, 1), 6, 2, 2)
    cn, xc, edges = p2.p2_dofmap(c, t)
    nn, nv = xc.shape[0], c.shape[0]
    mu, lam = fo.lame(2e11, 0.27)
    fv, opp, _ = fo.exterior_facets(t)
    fn = p2.facet_nodes(fv, edges, nv).astype(np.int32)
    mid = c[fv].mean(axis=1)
    left, right = mid[:, 0] == 0, mid[:, 0] == 4
    m = _lib.DeviceMesh.upload_p2(ctx, c, cn, nn)
    A = _lib.DeviceMatrix.create(m, 3)
    A.assemble_elasticity(mu, lam)
    b = _lib.DeviceVector(ctx, 3 * nn)
    _lib.assemble_source(m, b, (0.0, 0.0, -7800 * 9.81), ncomp=3, scale=-1.0)
    _lib.assemble_facet_load(m, b, fn[right], (1e6, 0.0, 0.0), ncomp=3, scale=-1.0)
    nodes = np.unique(fn[left])
    dofs = (nodes[:, None] * 3 + np.arange(3)).ravel()
    x = _lib.DeviceVector(ctx, 3 * nn)
    A.apply_dirichlet(b, dofs, 0.0, symmetric=True, x=x)
    info = A.solve(b, x, "cg", rtol=1e-13, maxit=200000)
    assert info["converged"] == 1
    Ko = p2.assemble_matrix(cn, p2.local_elasticity(c, t, mu, lam), nn, 3)
    bo = -p2.assemble_source(c, t, cn, nn, np.array([0.0, 0.0, -7800 * 9.81]), ncomp=3) - p2.assemble_facet_load(c, fv[right], fn[right], np.array([1e6, 0.0, 0.0]), nn, 3)
    Ao, bo = fo.apply_dirichlet(Ko, bo, dofs, np.zeros(dofs.size), True)
    assert fo.relative_l2(x.numpy(), fo.solve_direct(Ao, bo)) < 1e-9     # kappa ~ 1e7 with Jacobi: 1e-13 residual -> ~1e-10 error


# ------------------------------------------------------------------------------------ through the solver API
def test_reference_elasticity_example_with_its_own_degree_2_space():
    """examples/test_linear_elasticity.py:105-106 builds VectorFunctionSpace(mesh, "Lagrange", 2): the same case
    through LinearElasticitySolver on the P2 path (fixed end, bending force vector, body force, reference sign)."""
    import copy
    from collections import OrderedDict
    from fenicssolver_b200 import LinearElasticitySolver, SolverBase
    from fenicssolver_b200.dolfin_compat import AutoSubDomain, BoxMesh, Constant, Point, VectorFunctionSpace, near
    n = (10, 2, 2)
    mesh = BoxMesh(Point(0, 0, 0), Point(10, 1, 1), *n)
    V = VectorFunctionSpace(mesh, "Lagrange", 2)
    bcs = OrderedDict()
    bcs["fixed"] = {'boundary': AutoSubDomain(lambda x: near(x[0], 0.0)), 'boundary_id': 1, 'type': 'Dirichlet', 'value': (0, 0, 0)}
    bcs["bending"] = {'boundary': AutoSubDomain(lambda x: near(x[0], 10.0)), 'boundary_id': 2, 'type': 'force', 'value': Constant((0, 1e6, 0))}
    s = copy.deepcopy(SolverBase.default_case_settings)
    s['material'] = {'name': 'steel', 'elastic_modulus': 2e11, 'poisson_ratio': 0.27, 'density': 7800, 'thermal_expansion_coefficient': 2e-6}
    s['function_space'] = V
    s['boundary_conditions'] = bcs
    s['temperature_distribution'] = None
    s['solver_settings']['reference_values'] = {'temperature': 293}
    s['report_settings'] = {'logging_level': 30, 'logging_file': None, 'plotting_freq': 0, 'saving_freq': 0}
    s['body_source'] = (10 * 7800.0, 0.0, 0.0)
    solver = LinearElasticitySolver.LinearElasticitySolver(s)
    u = solver.solve()
    assert solver.solve_info["converged"] == 1
    c, t = fo.box_mesh((0, 0, 0), (10, 1, 1), *n)
    cn, xc, edges = p2.p2_dofmap(c, t)
    nn, nv = xc.shape[0], c.shape[0]
    mu, lam = fo.lame(2e11, 0.27)
    A = p2.assemble_matrix(cn, p2.local_elasticity(c, t, mu, lam), nn, 3)
    fv, opp, _ = fo.exterior_facets(t)
    rsel = c[fv].mean(axis=1)[:, 0] == 10
    fn = p2.facet_nodes(fv, edges, nv)
    b = -p2.assemble_facet_load(c, fv[rsel], fn[rsel], np.array([0, 1e6, 0]), nn, 3)
    b -= p2.assemble_source(c, t, cn, nn, np.array([10 * 7800.0, 0, 0]), ncomp=3)
    ln = np.flatnonzero(xc[:, 0] == 0)
    dofs = (ln[:, None] * 3 + np.arange(3)).ravel()
    Ab, bb = fo.apply_dirichlet(A, b, dofs, np.zeros(dofs.size), symmetric=True)
    uo = fo.solve_direct(Ab, bb)
    assert fo.relative_l2(u.vector().get_local(), uo) < 1e-9
    assert u.values.shape == (nn, 3)
    assert u.compute_vertex_values().shape == (3 * nv,)
    # P2 bends further than P1 on the same mesh (P1 locks): tip deflection grows by a clear margin
    s1 = copy.deepcopy({k: v for k, v in s.items() if k != 'function_space'})
    s1['function_space'] = VectorFunctionSpace(mesh, "Lagrange", 1)
    u1 = LinearElasticitySolver.LinearElasticitySolver(s1).solve()
    tip = np.flatnonzero(c[:, 0] == 10)
    assert abs(u.values[tip, 1].mean()) > 1.5 * abs(u1.values[tip, 1].mean())


def test_heat_settings_dict_with_fe_degree_2():
    """examples/test_heat_transfer.py-style settings with a degree-2 space (Dirichlet expression + heatFlux + HTC +
    body source) through ScalarTransportSolver, against the degree-2 oracle."""
    from fenicssolver_b200 import ScalarTransportSolver
    from fenicssolver_b200.dolfin_compat import AutoSubDomain, Constant, Expression, FunctionSpace, UnitSquareMesh, near
    nx, ny = 7, 5
    k, htc, Ta = 0.2, 12.0, 280.0
    mesh = UnitSquareMesh(nx, ny)

    def bc(kind, value, **kw):
        return {'temperature': dict(variable='temperature', type=kind, value=value, **kw)}
    bcs = {
        "hot": {'boundary': AutoSubDomain(lambda x: near(x[1], 1)), 'boundary_id': 1,
                'values': bc('Dirichlet', Expression("300 + 40*x[0]*x[0]", degree=2))},
        "flux": {'boundary': AutoSubDomain(lambda x: near(x[0], 1)), 'boundary_id': 2, 'values': bc('heatFlux', Constant(25.0))},
        "htc": {'boundary': AutoSubDomain(lambda x: near(x[1], 0)), 'boundary_id': 3, 'values': bc('HTC', Constant(htc), ambient=Constant(Ta))},
    }
    s = {'solver_name': 'ScalarEquationSolver', 'mesh': None, 'function_space': FunctionSpace(mesh, "CG", 2),
         'periodic_boundary': None, 'fe_degree': 2, 'boundary_conditions': bcs, 'body_source': None,
         'initial_values': {'temperature': 300},
         'material': {'density': 1000, 'specific_heat_capacity': 4200, 'thermal_conductivity': k},
         'solver_settings': {'transient_settings': {'transient': False, 'starting_time': 0, 'time_step': 0.1, 'ending_time': 1},
                             'reference_values': {'temperature': 300},
                             'solver_parameters': {"relative_tolerance": 1e-9, "maximum_iterations": 500}},
         'scalar_name': 'temperature'}
    solver = ScalarTransportSolver.ScalarTransportSolver(s)
    solver.material['conductivity'] = k
    T = solver.solve()
    assert solver.solve_info["converged"] == 1
    c, t = fo.unit_square_mesh(nx, ny)
    cn, xc, edges = p2.p2_dofmap(c, t)
    nn, nv = xc.shape[0], c.shape[0]
    fv, opp, _ = fo.exterior_facets(t)
    fn = p2.facet_nodes(fv, edges, nv)
    mid = c[fv].mean(axis=1)
    right, bottom = mid[:, 0] == 1, mid[:, 1] == 0
    A = p2.assemble_matrix(cn, p2.local_laplace(c, t, k), nn) + fo._scatter(fn[bottom], p2.local_facet_mass(c, fv[bottom], htc), nn)
    b = p2.assemble_facet_load(c, fv[right], fn[right], 25.0, nn) + p2.assemble_facet_load(c, fv[bottom], fn[bottom], htc * Ta, nn)
    d = np.flatnonzero(xc[:, 1] == 1)
    Ab, bb = fo.apply_dirichlet(A, b, d, 300 + 40 * xc[d, 0] ** 2, symmetric=True)
    assert fo.relative_l2(T.vector().get_local(), fo.solve_direct(Ab, bb)) < 1e-10
    assert T.values.shape == (nn,) and T.compute_vertex_values().shape == (nv,)
