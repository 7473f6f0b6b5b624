"""Multigrid-preconditioned CG (`solver_parameters['preconditioner'] = 'gmg'`): the counterpart of the reference's
CG + GAMG elasticity path (SolverBase.py:643-672) on generated box meshes.  Same answers as the oracle's direct solve,
iteration counts independent of the mesh size."""
import copy

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import fem_oracle as fo  # noqa: E402  (checker only)
from fenicssolver_b200 import LinearElasticitySolver, ScalarTransportSolver, SolverBase  # noqa: E402
from fenicssolver_b200.dolfin_compat import AutoSubDomain, BoxMesh, Constant, Mesh, Point, UnitCubeMesh, UnitSquareMesh, near  # noqa: E402

QUIET = {'logging_level': 40, 'logging_file': None, 'plotting_freq': 0, 'saving_freq': 0, 'plotting_interactive': False}


def heat_settings(mesh, precond, transient=None, dim=3):
    ax = dim - 1
    return {'solver_name': 'ScalarTransportSolver', 'scalar_name': 'temperature', 'mesh': mesh, 'fe_degree': 1, 'fe_family': 'CG',
            'material': {'density': 1000, 'specific_heat_capacity': 500, 'thermal_conductivity': 20},
            'boundary_conditions': {
                'inlet': {'boundary': lambda x: near(x[ax], 0.0), 'boundary_id': 1, 'type': 'Dirichlet', 'value': 350},
                'outlet': {'boundary': lambda x: near(x[ax], 1.0), 'boundary_id': 2, 'type': 'HTC', 'value': 400.0, 'ambient': 300.0},
                'side': {'boundary': lambda x: near(x[0], 0.0), 'boundary_id': 3, 'type': 'heatFlux', 'value': 2000.0}},
            'body_source': 1000, 'initial_values': {'temperature': 293},
            'solver_settings': {'transient_settings': transient or {'transient': False, 'starting_time': 0, 'time_step': 0.01, 'ending_time': 0.03},
                                'reference_values': {'temperature': 293}, 'solver_parameters': {'preconditioner': precond}},
            'report_settings': QUIET}


def heat_oracle(c, t, dim):
    nv = c.shape[0]
    ax = dim - 1
    fv, opp, _ = fo.exterior_facets(t)
    mid = c[fv].mean(axis=1)
    inlet = np.nonzero(c[:, ax] == 0)[0]
    A, b = fo.heat_system(c, t, 20.0, [(inlet, 350.0)], source=1000.0, neumann=[(fv[mid[:, 0] == 0], 2000.0)],
                          robin=[(fv[mid[:, ax] == 1], 400.0, 300.0)])
    return fo.solve_direct(A, b)


@pytest.mark.parametrize("N", [8, 16, 32])
def test_gmg_heat_matches_direct_solve_with_mesh_independent_iterations(N):
    solver = ScalarTransportSolver.ScalarTransportSolver(heat_settings(UnitCubeMesh(N, N, N), 'gmg'))
    T = solver.solve()
    info = solver.solve_info
    assert info['converged'] == 1 and info['iterations'] <= 30, info
    assert info['mg_levels'] == {8: 3, 16: 4, 32: 5}[N]
    c, t = fo.unit_cube_mesh(N, N, N)
    assert fo.relative_l2(T.values, heat_oracle(c, t, 3)) < 1e-10
    if N == 32:
        jac = ScalarTransportSolver.ScalarTransportSolver(heat_settings(UnitCubeMesh(N, N, N), 'jacobi'))
        Tj = jac.solve()
        assert jac.solve_info['iterations'] > 5 * info['iterations']
        assert fo.relative_l2(T.values, Tj.values) < 1e-10
        w = [solver._mg.omega(l) for l in range(info['mg_levels'])]
        assert all(0.3 < x < 0.8 for x in w), w


@pytest.mark.parametrize("case", ["heat3d", "heat2d", "elasticity"])
def test_vcycle_matches_the_numpy_restatement(case):
    """One V(2,2) cycle on a random residual: the CUDA transfer/smoothing kernels against oracle/mg_oracle.py on the
    downloaded level matrices with the library's own dampings."""
    import scipy.sparse as sp
    from oracle import mg_oracle as mo
    from fenicssolver_b200 import _lib
    if case == "heat3d":
        solver, n, ncomp = ScalarTransportSolver.ScalarTransportSolver(heat_settings(UnitCubeMesh(8, 8, 8), 'gmg')), (8, 8, 8), 1
    elif case == "heat2d":
        solver, n, ncomp = ScalarTransportSolver.ScalarTransportSolver(heat_settings(UnitSquareMesh(16, 8), 'gmg', dim=2)), (16, 8), 1
    else:
        s = copy.deepcopy(SolverBase.default_case_settings)
        s.update({'mesh': BoxMesh(Point(0, 0, 0), Point(2, 1, 1), 8, 4, 4), 'material': {'elastic_modulus': 2e11, 'poisson_ratio': 0.27, 'density': 7800},
                  'boundary_conditions': {'clamp': {'boundary': AutoSubDomain(lambda x: near(x[0], 0.0)), 'boundary_id': 1, 'type': 'Dirichlet',
                                                    'value': Constant((0, 0, 0))}},
                  'body_source': (0.0, 0.0, -7800 * 9.81), 'report_settings': QUIET})
        s['solver_settings'] = dict(s['solver_settings'], solver_parameters={'preconditioner': 'gmg'})
        solver, n, ncomp = LinearElasticitySolver.LinearElasticitySolver(s), (8, 4, 4), 3
    solver.solve()
    mg = solver._mg
    levels, transfers, nl = [], [], list(n)
    for l, A in enumerate(mg.matrices):
        rp, ci, va = A.download_csr()
        M = sp.csr_matrix((va, ci.astype(np.int64), rp), shape=(rp.size - 1, rp.size - 1))
        d = M.diagonal()
        rows = np.repeat(np.arange(M.shape[0]), np.diff(M.indptr))
        offdiag = np.bincount(rows, weights=np.abs(M.data) * (rows != M.indices), minlength=M.shape[0])
        levels.append({'A': M, 'dinv': 1.0 / d, 'omega': mg.omega(l), 'bc': (d == 1.0) & (offdiag == 0.0)})
        if l + 1 < len(mg.matrices):
            transfers.append(mo.prolongation(nl, ncomp))
            nl = [k // 2 for k in nl]
    assert all(0.3 < L['omega'] < 0.8 for L in levels), [L['omega'] for L in levels]
    rng = np.random.default_rng(0)
    r = rng.standard_normal(levels[0]['A'].shape[0])
    r[levels[0]['bc']] = 0.0
    ctx = solver.device_space().ctx
    z = _lib.DeviceVector(ctx, r.size)
    mg.apply(_lib.DeviceVector.from_numpy(ctx, r), z, 2)
    zo = mo.vcycle(levels, transfers, r, nu=2)
    assert np.abs(z.numpy() - zo).max() <= 1e-11 * np.abs(zo).max()


def test_gmg_two_dimensional_heat():
    N = 32
    solver = ScalarTransportSolver.ScalarTransportSolver(heat_settings(UnitSquareMesh(N, N), 'gmg', dim=2))
    T = solver.solve()
    assert solver.solve_info['converged'] == 1 and solver.solve_info['iterations'] <= 30
    c, t = fo.unit_square_mesh(N, N)
    assert fo.relative_l2(T.values, heat_oracle(c, t, 2)) < 1e-10


def test_gmg_elasticity_cantilever():
    """The case solve_amg exists for: 3-D elasticity, clamped at x = 0, gravity and an end pressure."""
    n = (32, 8, 8)

    def settings(precond):
        mesh = BoxMesh(Point(0, 0, 0), Point(4, 1, 1), *n)
        s = copy.deepcopy(SolverBase.default_case_settings)
        s.update({'mesh': mesh, 'material': {'elastic_modulus': 2e11, 'poisson_ratio': 0.27, 'density': 7800},
                  'boundary_conditions': {'clamp': {'boundary': AutoSubDomain(lambda x: near(x[0], 0.0)), 'boundary_id': 1, 'type': 'Dirichlet',
                                                    'value': Constant((0, 0, 0))},
                                          'tip': {'boundary': AutoSubDomain(lambda x: near(x[0], 4.0)), 'boundary_id': 2, 'type': 'pressure', 'value': 1e5}},
                  'body_source': (0.0, 0.0, -7800 * 9.81), 'report_settings': QUIET})
        s['solver_settings'] = dict(s['solver_settings'], solver_parameters={'preconditioner': precond})
        return s
    mg = LinearElasticitySolver.LinearElasticitySolver(settings('gmg'))
    u = mg.solve()
    assert mg.solve_info['converged'] == 1 and mg.solve_info['iterations'] <= 60, mg.solve_info
    jac = LinearElasticitySolver.LinearElasticitySolver(settings('jacobi'))
    uj = jac.solve()
    assert jac.solve_info['iterations'] > 8 * mg.solve_info['iterations']
    assert fo.relative_l2(u.vector().get_local(), uj.vector().get_local()) < 1e-8
    c, t = fo.box_mesh((0, 0, 0), (4, 1, 1), *n)
    nv = c.shape[0]
    mu, lam = fo.lame(2e11, 0.27)
    A = fo.assemble_matrix(t, fo.local_elasticity(c, t, mu, lam), nv, 3)
    fv, opp, _ = fo.exterior_facets(t)
    sel = c[fv].mean(axis=1)[:, 0] == 4.0
    meas, nrm = fo.facet_measure(c, fv[sel], opp[sel])
    b = -fo.assemble_source(c, t, np.array([0, 0, -7800 * 9.81]), ncomp=3) - fo.assemble_facet_load(c, fv[sel], 1e5 * nrm, nv, 3)
    lv = np.nonzero(c[:, 0] == 0)[0]
    dofs = (lv[:, None] * 3 + np.arange(3)).ravel()
    Ab, bb = fo.apply_dirichlet(A, b, dofs, np.zeros(dofs.size), symmetric=True)
    assert fo.relative_l2(u.vector().get_local(), fo.solve_direct(Ab, bb)) < 1e-9


def test_gmg_transient_steps_match_jacobi_path():
    N, nsteps = 16, 3
    dt = 1000 * 500.0 / (N * N) / 20.0
    tr = {'transient': True, 'starting_time': 0.0, 'time_step': dt, 'ending_time': dt * (nsteps - 0.5)}
    a = ScalarTransportSolver.ScalarTransportSolver(heat_settings(UnitCubeMesh(N, N, N), 'gmg', dict(tr)))
    b = ScalarTransportSolver.ScalarTransportSolver(heat_settings(UnitCubeMesh(N, N, N), 'jacobi', dict(tr)))
    Ta, Tb = a.solve(), b.solve()
    assert a.current_step == nsteps and a.solve_info['iterations'] < b.solve_info['iterations']
    assert fo.relative_l2(Ta.values, Tb.values) < 1e-10


def test_gmg_needs_a_box_mesh_and_a_symmetric_problem():
    base = UnitCubeMesh(4, 4, 4)
    s = heat_settings(Mesh(base.coordinates().copy(), base.cells().copy()), 'gmg')
    with pytest.raises(SolverBase.SolverError):
        ScalarTransportSolver.ScalarTransportSolver(s).solve()
    s = heat_settings(UnitCubeMesh(4, 4, 4), 'gmg')
    s['convective_velocity'] = (0.0, 0.0, 1e-3)
    with pytest.raises(SolverBase.SolverError):
        ScalarTransportSolver.ScalarTransportSolver(s).solve()


def test_solve_amg_defaults_to_multigrid_on_even_boxes():
    """LinearElasticitySolver.solve_form -> solve_amg (3-D): multigrid-preconditioned CG when the box can be coarsened and
    no preconditioner was named, Jacobi-CG otherwise; the settings dict is left as the user wrote it."""
    def settings(n):
        s = copy.deepcopy(SolverBase.default_case_settings)
        s.update({'mesh': BoxMesh(Point(0, 0, 0), Point(4, 1, 1), *n), 'material': {'elastic_modulus': 2e11, 'poisson_ratio': 0.27, 'density': 7800},
                  'boundary_conditions': {'clamp': {'boundary': AutoSubDomain(lambda x: near(x[0], 0.0)), 'boundary_id': 1, 'type': 'Dirichlet',
                                                    'value': Constant((0, 0, 0))}},
                  'body_source': (0.0, 0.0, -7800 * 9.81), 'report_settings': QUIET})
        s['solver_settings'] = dict(s['solver_settings'], solver_parameters={})
        return s
    even = LinearElasticitySolver.LinearElasticitySolver(settings((16, 4, 4)))
    even.solve()
    assert even.solve_info.get('mg_levels') == 2 and even.solve_info['converged'] == 1
    assert 'preconditioner' not in even.solver_settings['solver_parameters']
    odd = LinearElasticitySolver.LinearElasticitySolver(settings((15, 3, 3)))
    odd.solve()
    assert 'mg_levels' not in odd.solve_info and odd.solve_info['converged'] == 1
    named = settings((16, 4, 4))
    named['solver_settings']['solver_parameters']['preconditioner'] = 'jacobi'
    jac = LinearElasticitySolver.LinearElasticitySolver(named)
    uj = jac.solve()
    assert 'mg_levels' not in jac.solve_info
    assert fo.relative_l2(even.result.vector().get_local(), uj.vector().get_local()) < 1e-8
