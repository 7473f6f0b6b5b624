"""GPU parity tests, API level: the reference's example cases through ScalarTransportSolver /
LinearElasticitySolver / main.load_settings, checked against the CPU oracle's direct (LU) solve — the
solver the reference's scalar path really uses — and against analytic answers.

Each test mirrors a script under /root/reference/examples (cited); the reference's own scripts assert
nothing, so the numbers come from the oracle and the known-answer tests of SURVEY 8c.
Tolerance: 1e-10 relative L2 (BASELINE.json north_star).
"""
import copy
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import fem_oracle as fo  # noqa: E402  (checker only)
from fenicssolver_b200 import LinearElasticitySolver, ScalarTransportSolver, SolverBase  # noqa: E402
from fenicssolver_b200.dolfin_compat import (AutoSubDomain, BoxMesh, Constant, FunctionSpace, Point, SubDomain,  # noqa: E402
                                             UnitCubeMesh, UnitSquareMesh, VectorFunctionSpace, near)
from fenicssolver_b200.main import load_settings, main  # noqa: E402

TOL = 1e-10


def write_dolfin_xml(tmp_path, g):
    """Re-create data/mesh.xml + marker files in dolfin-XML from the golden arrays (the GPU box has no
    /root/reference)."""
    c, t = g["coords"], g["cells"]
    p = os.path.join(tmp_path, "mesh.xml")
    with open(p, "w") as f:
        f.write('<?xml version="1.0" encoding="UTF-8"?>\n\n<dolfin xmlns:dolfin="http://www.fenicsproject.org">\n')
        f.write('  <mesh celltype="tetrahedron" dim="3">\n    <vertices size="%d">\n' % c.shape[0])
        for i, x in enumerate(c):
            f.write('      <vertex index="%d" x="%.16e" y="%.16e" z="%.16e"/>\n' % (i, x[0], x[1], x[2]))
        f.write('    </vertices>\n    <cells size="%d">\n' % t.shape[0])
        for i, v in enumerate(t):
            f.write('      <tetrahedron index="%d" v0="%d" v1="%d" v2="%d" v3="%d"/>\n' % (i, v[0], v[1], v[2], v[3]))
        f.write('    </cells>\n  </mesh>\n</dolfin>\n')
    for name, dim, vals in (("mesh_facet_region.xml", 2, g["facet_tags"]), ("mesh_physical_region.xml", 3, g["cell_tags"])):
        with open(os.path.join(tmp_path, name), "w") as f:
            f.write('<?xml version="1.0" encoding="UTF-8"?>\n<dolfin xmlns:dolfin="http://fenicsproject.org">\n')
            f.write('  <mesh_function type="uint" dim="%d" size="%d">\n' % (dim, vals.size))
            for i, v in enumerate(vals):
                f.write('    <entity index="%d" value="%d"/>\n' % (i, v))
            f.write('  </mesh_function>\n</dolfin>\n')
    return p


def test_customized_case_settings_json_fixture(tmp_path, golden_dir):
    """examples/test_customized_case_settings.py:52-66 with data/TestHeatTransfer.json: exact answer 350-2.5z."""
    g = np.load(os.path.join(golden_dir, "fixture_mesh.npz"))
    e = np.load(os.path.join(golden_dir, "fixture_expected.npz"))
    mesh_path = write_dolfin_xml(str(tmp_path), g)
    settings = json.load(open(os.path.join(golden_dir, "TestHeatTransfer.json")))
    settings["mesh"] = mesh_path
    case = os.path.join(str(tmp_path), "case.json")
    json.dump(settings, open(case, "w"))
    settings = load_settings(case)
    solver = ScalarTransportSolver.ScalarTransportSolver(settings)
    T = solver.solve()
    assert (solver.boundary_facets.values == 1).sum() == 100 and (solver.boundary_facets.values == 2).sum() == 100
    assert fo.relative_l2(T.vector().get_local(), e["analytic"]) < TOL
    assert fo.relative_l2(T.vector().get_local(), e["solution"]) < TOL
    assert solver.solve_info["converged"] == 1
    # main() dispatches on solver_name and returns after plot() (batch: no-op)
    s2 = main(case)
    assert fo.relative_l2(s2.result.values, e["analytic"]) < TOL


# ---------------------------------------------------------------------------------- test_heat_transfer.py
def heat_settings(bcs, nx=40, ny=40, material=None, transient=None):
    mesh = UnitSquareMesh(nx, ny)
    Q = FunctionSpace(mesh, "CG", 1)
    return {'solver_name': 'ScalarEquationSolver',
            'mesh': None, 'function_space': Q, 'periodic_boundary': None, 'fe_degree': 1,
            'boundary_conditions': bcs, 'body_source': None,
            'initial_values': {'temperature': 300},
            'material': material or {'density': 1000, 'specific_heat_capacity': 4200, 'thermal_conductivity': 0.1},
            'solver_settings': {
                'transient_settings': transient or {'transient': False, 'starting_time': 0, 'time_step': 0.1, 'ending_time': 1},
                'reference_values': {'temperature': 300},
                'solver_parameters': {"relative_tolerance": 1e-9, "maximum_iterations": 500, "monitor_convergence": True},
            },
            'scalar_name': 'temperature',
            }, mesh


top = AutoSubDomain(lambda x: near(x[1], 1))
bottom = AutoSubDomain(lambda x: near(x[1], 0))
left = AutoSubDomain(lambda x: near(x[0], 0))
right = AutoSubDomain(lambda x: near(x[0], 1))


def square_sets(nx, ny):
    c, t = fo.unit_square_mesh(nx, ny)
    fv, opp, _ = fo.exterior_facets(t)
    mid = c[fv].mean(axis=1)
    sel = {"top": mid[:, 1] == 1, "bottom": mid[:, 1] == 0, "left": mid[:, 0] == 0, "right": mid[:, 0] == 1}
    return c, t, fv, sel


def test_heat_transfer_dirichlet_plus_flux():
    """examples/test_heat_transfer.py:47-57,151-153,170: top Dirichlet 360, bottom heatFlux, k=0.6 ->
    T = 360 + 60(1-y) exactly (SURVEY 8c KAT 5), new-style 'values' boundary dicts."""
    k, T_hot, T_cold = 0.6, 360, 300
    heat_flux = (T_hot - T_cold) / 1.0 * k
    bcs = {
        "hot": {'boundary': top, 'boundary_id': 1, 'values': {'temperature': {'variable': 'temperature', 'type': 'Dirichlet', 'value': Constant(T_hot)}}},
        "left": {'boundary': left, 'boundary_id': 3, 'values': {'temperature': {'variable': 'temperature', 'type': 'heatFlux', 'value': Constant(0)}}},
        "right": {'boundary': right, 'boundary_id': 4, 'values': {'temperature': {'variable': 'temperature', 'type': 'symmetry', 'value': None}}},
        "cold": {'boundary': bottom, 'boundary_id': 2, 'values': {'temperature': {'variable': 'temperature', 'type': 'heatFlux', 'value': Constant(heat_flux)}}},
    }
    settings, mesh = heat_settings(bcs)
    solver = ScalarTransportSolver.ScalarTransportSolver(settings)
    solver.material['conductivity'] = k
    T = solver.solve()
    y = mesh.coordinates()[:, 1]
    assert fo.relative_l2(T.vector().get_local(), 360 + 60 * (1 - y)) < TOL
    c, t, fv, sel = square_sets(40, 40)
    A, b = fo.heat_system(c, t, k, [(np.unique(fv[sel["top"]]), 360.0)], neumann=[(fv[sel["bottom"]], heat_flux)])
    assert fo.relative_l2(T.vector().get_local(), fo.solve_direct(A, b)) < TOL


def test_heat_transfer_htc_and_convection():
    """examples/test_heat_transfer.py:154-166,224 (the default test()): heatFlux on top, HTC on the bottom,
    convective velocity (0.005,-0.005) -> nonsymmetric system, BiCGStab; against the oracle's LU."""
    k, htc, Ta = 0.6, 100.0, 300.0
    heat_flux = 36.0
    bcs = {
        "hot": {'boundary': top, 'boundary_id': 1, 'values': {'temperature': {'variable': 'temperature', 'type': 'heatFlux', 'value': Constant(heat_flux)}}},
        "left": {'boundary': left, 'boundary_id': 3, 'values': {'temperature': {'variable': 'temperature', 'type': 'heatFlux', 'value': Constant(0)}}},
        "right": {'boundary': right, 'boundary_id': 4, 'values': {'temperature': {'variable': 'temperature', 'type': 'symmetry', 'value': None}}},
        "cold": {'boundary': bottom, 'boundary_id': 2, 'values': {'temperature': {'variable': 'temperature', 'type': 'HTC', 'value': Constant(htc), 'ambient': Constant(Ta)}}},
    }
    settings, mesh = heat_settings(bcs, 24, 24)
    settings['convective_velocity'] = Constant((0.4e-6, -0.4e-6))      # global Peclet c|v|L/k ~ 4: well-posed with a flux inflow
    solver = ScalarTransportSolver.ScalarTransportSolver(settings)
    solver.material['conductivity'] = k
    T = solver.solve()
    c, t, fv, sel = square_sets(24, 24)
    A, b = fo.heat_system(c, t, k, [], neumann=[(fv[sel["top"]], heat_flux)], robin=[(fv[sel["bottom"]], htc, Ta)],
                          velocity=np.array([0.4e-6, -0.4e-6]), capacity=1000 * 4200.0, symmetric=False)
    assert fo.relative_l2(T.vector().get_local(), fo.solve_direct(A, b)) < TOL
    assert solver.solve_info["converged"] == 1


def test_heat_transfer_anisotropic_tensor():
    """examples/test_heat_transfer.py:91,139-140 with a constant anisotropic tensor (get_material_value :326-330)."""
    K = [[2.0, 0.5], [0.5, 1.0]]
    bcs = {"hot": {'boundary': top, 'boundary_id': 1, 'type': 'Dirichlet', 'value': 360},
           "cold": {'boundary': bottom, 'boundary_id': 2, 'type': 'Dirichlet', 'value': 300},
           "left": {'boundary': left, 'boundary_id': 3, 'type': 'heatFlux', 'value': 50.0}}
    settings, mesh = heat_settings(bcs, 16, 12)
    settings['body_source'] = 1000.0
    solver = ScalarTransportSolver.ScalarTransportSolver(settings)
    solver.material['conductivity'] = K
    T = solver.solve()
    c, t, fv, sel = square_sets(16, 12)
    A, b = fo.heat_system(c, t, np.array(K), [(np.unique(fv[sel["top"]]), 360.0), (np.unique(fv[sel["bottom"]]), 300.0)],
                          source=1000.0, neumann=[(fv[sel["left"]], 50.0)])
    assert fo.relative_l2(T.vector().get_local(), fo.solve_direct(A, b)) < TOL


def test_electrostatics_old_style_bcs():
    """examples/test_electrostatics.py:73-78,108: flat BC dicts, electric_potential, Dirichlet top/bottom."""
    bcs = {"hot": {'boundary': top, 'boundary_id': 1, 'type': 'Dirichlet', 'value': Constant(360)},
           "left": {'boundary': left, 'boundary_id': 3, 'type': 'flux', 'value': Constant(0)},
           "right": {'boundary': right, 'boundary_id': 4, 'type': 'flux', 'value': Constant(0)},
           "cold": {'boundary': bottom, 'boundary_id': 2, 'type': 'Dirichlet', 'value': Constant(300)}}
    settings, mesh = heat_settings(bcs, 40, 40, material={'relative_electric_permittivity': 11.7})
    settings['scalar_name'] = 'electric_potential'
    settings['initial_values'] = {'electric_potential': 300}
    solver = ScalarTransportSolver.ScalarTransportSolver(settings)
    V = solver.solve()
    y = mesh.coordinates()[:, 1]
    assert fo.relative_l2(V.vector().get_local(), 300 + 60 * y) < TOL
    assert abs(solver.conductivity() - 11.7 * 8.854187817e-12) < 1e-25


def test_unit_cube_heat_kat4_plain_data_mesh():
    """SURVEY 8c KAT 4 / config C2 at small N through the JSON-able plain-data mesh description."""
    N = 12
    settings = {'solver_name': 'ScalarTransportSolver', 'scalar_name': 'temperature',
                'mesh': {'type': 'UnitCubeMesh', 'n': [N, N, N]}, 'fe_degree': 1, 'fe_family': 'CG',
                'material': {'density': 1000, 'specific_heat_capacity': 500, 'thermal_conductivity': 20},
                'boundary_conditions': {
                    'bottom': {'boundary': lambda x: near(x[2], 0.0), 'boundary_id': 1, 'type': 'Dirichlet', 'value': 350},
                    'top': {'boundary': lambda x: near(x[2], 1.0), 'boundary_id': 2, 'type': 'Dirichlet', 'value': 300}},
                'body_source': 1000, 'initial_values': {'temperature': 293},
                'solver_settings': {'transient_settings': {'transient': False, 'starting_time': 0, 'time_step': 0.01, 'ending_time': 0.03},
                                    'reference_values': {'temperature': 293},
                                    'solver_parameters': {'relative_tolerance': 1e-7, 'maximum_iterations': 500}}}
    solver = ScalarTransportSolver.ScalarTransportSolver(settings)
    T = solver.solve()
    z = solver.mesh.coordinates()[:, 2]
    assert fo.relative_l2(T.values, 350 - 50 * z + 1000 * z * (1 - z) / 40) < TOL
    assert solver.krylov_parameters()["rtol"] == 1e-12      # parity mode: the loose JSON tolerance does not apply
    assert solver.solve_info["rnorm"] <= 1e-12 * solver.solve_info["bnorm"]


def test_transient_crank_nicolson_matches_oracle_stepping():
    """SolverBase.solve_transient (:492-542) + the CN form (ScalarTransportSolver.py:287-293): five steps of
    advection-diffusion with per-step re-assembly, against the oracle stepping the same recurrences with LU."""
    N = 8
    k, rho, cp = 0.6, 1000.0, 4200.0
    c_ = rho * cp
    h = 1.0 / N
    dt = c_ * h * h / k
    vel = (k / (c_ * h), 0.0, -0.5 * k / (c_ * h))          # cell Peclet ~ 1
    bcs = {'bottom': {'boundary': lambda x: near(x[2], 0.0), 'boundary_id': 1, 'type': 'Dirichlet', 'value': 360},
           'top': {'boundary': lambda x: near(x[2], 1.0), 'boundary_id': 2, 'type': 'Dirichlet', 'value': 300},
           'side': {'boundary': lambda x: near(x[0], 0.0), 'boundary_id': 3, 'type': 'heatFlux', 'value': 5.0}}
    nsteps = 5
    settings = {'solver_name': 'ScalarTransportSolver', 'scalar_name': 'temperature', 'mesh': UnitCubeMesh(N, N, N),
                'material': {'density': rho, 'specific_heat_capacity': cp, 'thermal_conductivity': k},
                'boundary_conditions': bcs, 'body_source': 2.0, 'initial_values': {'temperature': 300},
                'convective_velocity': vel,
                'solver_settings': {'transient_settings': {'transient': True, 'starting_time': 0.0, 'time_step': dt, 'ending_time': dt * (nsteps - 0.5)},
                                    'reference_values': {'temperature': 300}, 'solver_parameters': {}},
                'report_settings': {'logging_level': 30, 'logging_file': None, 'plotting_freq': 0, 'saving_freq': 0}}
    solver = ScalarTransportSolver.ScalarTransportSolver(settings)
    T = solver.solve()
    assert solver.current_step == nsteps
    # oracle
    c, t = fo.unit_cube_mesh(N, N, N)
    nv = c.shape[0]
    fv, opp, _ = fo.exterior_facets(t)
    mid = c[fv].mean(axis=1)
    z0, z1 = np.nonzero(c[:, 2] == 0)[0], np.nonzero(c[:, 2] == 1)[0]
    K = fo.assemble_matrix(t, fo.local_laplace(c, t, k), nv)
    M = fo.assemble_matrix(t, fo.local_mass(c, t, c_), nv)
    C = fo.assemble_matrix(t, fo.local_advection(c, t, np.array(vel), c_), nv)
    loads = fo.assemble_source(c, t, 2.0) + fo.assemble_facet_load(c, fv[mid[:, 0] == 0], 5.0, nv)
    Tn = np.full(nv, 300.0)
    dofs = np.concatenate([z0, z1]); vals = np.concatenate([np.full(z0.size, 360.0), np.full(z1.size, 300.0)])
    for _ in range(nsteps):
        A = M / dt + 0.5 * K + C
        b = (M / dt) @ Tn - 0.5 * (K @ Tn) + loads
        Ab, bb = fo.apply_dirichlet(A, b, dofs, vals, symmetric=False)
        Tn = fo.solve_direct(Ab, bb)
    assert fo.relative_l2(T.values, Tn) < TOL


# ---------------------------------------------------------------------------------- test_linear_elasticity.py
class Left(SubDomain):
    def inside(self, x, on_boundary):
        return near(x[0], 0.0)


class Right(SubDomain):
    def inside(self, x, on_boundary):
        return near(x[0], 10.0)


def elasticity_case(boundary_type, has_body_source):
    nx, ny, nz = 12, 3, 3
    mesh = BoxMesh(Point(0, 0, 0), Point(10, 1, 1), nx, ny, nz)
    V = VectorFunctionSpace(mesh, "Lagrange", 1)       # the reference uses degree 2 (:105-106); P2 is a next-tier item
    from collections import OrderedDict
    bcs = OrderedDict()
    bcs["fixed"] = {'boundary': Left(), 'boundary_id': 1, 'type': 'Dirichlet', 'value': (Constant(0), None, None)}
    if boundary_type == 1:
        bcs["displ"] = {'boundary': Right(), 'boundary_id': 2, 'type': 'Dirichlet', 'value': Constant((0, 0, 1e-3))}
    elif boundary_type == 2:
        bcs["fixed"]['value'] = Constant((0, 0, 0))
        bcs["tensile"] = {'boundary': Right(), 'boundary_id': 2, 'type': 'stress', 'value': Constant((1e8, 0, 0))}
    elif boundary_type == 3:
        bcs["fixed"]['value'] = (0, 0, 0)
        bcs["bending"] = {'boundary': Right(), 'boundary_id': 2, 'type': 'force', 'value': Constant((0, 1e6, 0))}
    elif boundary_type == 4:
        bcs["fixed"]['value'] = (0, 0, 0)
        bcs["push"] = {'boundary': Right(), 'boundary_id': 2, 'type': 'force', 'value': 2.5e5}
    s = copy.deepcopy(SolverBase.default_case_settings)
    s['material'] = {'name': 'steel', 'elastic_modulus': 2e11, 'poisson_ratio': 0.27, 'density': 7800, 'thermal_expansion_coefficient': 2e-6}
    s['function_space'] = V
    s['boundary_conditions'] = bcs
    s['temperature_distribution'] = None
    s['solver_settings']['reference_values'] = {'temperature': 293}
    s['report_settings'] = {'logging_level': 30, 'logging_file': None, 'plotting_freq': 0, 'saving_freq': 0}
    if has_body_source:
        s['body_source'] = (10 * 7800.0, 0.0, 0.0)
    return s, mesh, (nx, ny, nz)


@pytest.mark.parametrize("boundary_type,has_body_source", [(1, False), (2, False), (3, True), (4, False)])
def test_linear_elasticity_cases(boundary_type, has_body_source):
    """examples/test_linear_elasticity.py:38-131 (per-component Dirichlet, prescribed displacement, stress, force
    vector, scalar force over the boundary area, body force), including the reference's flipped load sign."""
    s, mesh, n = elasticity_case(boundary_type, has_body_source)
    solver = LinearElasticitySolver.LinearElasticitySolver(s)
    u = solver.solve()
    assert solver.solve_info["converged"] == 1
    c, t = fo.box_mesh((0, 0, 0), (10, 1, 1), *n)
    nv = c.shape[0]
    mu, lam = fo.lame(2e11, 0.27)
    A = fo.assemble_matrix(t, fo.local_elasticity(c, t, mu, lam), nv, 3)
    fv, opp, _ = fo.exterior_facets(t)
    mid = c[fv].mean(axis=1)
    rsel = mid[:, 0] == 10
    lv, rv = np.nonzero(c[:, 0] == 0)[0], np.nonzero(c[:, 0] == 10)[0]
    b = np.zeros(3 * nv)
    if boundary_type == 1:
        dofs = np.concatenate([lv * 3, (rv[:, None] * 3 + np.arange(3)).ravel()])
        vals = np.concatenate([np.zeros(lv.size), np.tile([0, 0, 1e-3], rv.size)])
    else:
        dofs = (lv[:, None] * 3 + np.arange(3)).ravel()
        vals = np.zeros(dofs.size)
        if boundary_type == 2:
            b -= fo.assemble_facet_load(c, fv[rsel], np.array([1e8, 0, 0]), nv, 3)
        elif boundary_type == 3:
            b -= fo.assemble_facet_load(c, fv[rsel], np.array([0, 1e6, 0]), nv, 3)
        else:
            meas, nrm = fo.facet_measure(c, fv[rsel], opp[rsel])
            b -= fo.assemble_facet_load(c, fv[rsel], (2.5e5 / meas.sum()) * nrm, nv, 3)
    if has_body_source:
        b -= fo.assemble_source(c, t, np.array([10 * 7800.0, 0, 0]), ncomp=3)
    Ab, bb = fo.apply_dirichlet(A, b, dofs, vals, symmetric=True)
    uo = fo.solve_direct(Ab, bb)
    assert fo.relative_l2(u.vector().get_local(), uo) < TOL
    assert u.values.shape == (nv, 3)
    if boundary_type == 2:
        assert u.values[rv, 0].mean() < 0       # reference sign flip: a "tensile" stress compresses the bar


def test_conventional_load_sign_switch():
    s, mesh, n = elasticity_case(2, False)
    s['reference_load_sign'] = False
    u = LinearElasticitySolver.LinearElasticitySolver(s).solve()
    rv = np.nonzero(mesh.coordinates()[:, 0] == 10)[0]
    assert u.values[rv, 0].mean() > 0


def test_unsupported_features_raise_solver_error():
    bcs = {"hot": {'boundary': top, 'boundary_id': 1, 'type': 'Dirichlet', 'value': 1.0},
           "odd": {'boundary': bottom, 'boundary_id': 2, 'type': 'no_such_type', 'value': 1.0}}
    settings, _ = heat_settings(bcs, 4, 4)
    with pytest.raises(SolverBase.SolverError):
        ScalarTransportSolver.ScalarTransportSolver(settings).solve()
    with pytest.raises(SolverBase.SolverError):
        ScalarTransportSolver.ScalarTransportSolver("not a dict")
    bcs2 = {"hot": {'boundary': top, 'boundary_id': 1, 'type': 'Dirichlet', 'value': 1.0}}
    settings, _ = heat_settings(bcs2, 4, 4)
    settings['material']['capacity'] = lambda T: 4.2e6 * (1 + 1e-3 * T)            # nonlinear capacity: unsupported in the reference too
    with pytest.raises(SolverBase.SolverError):
        ScalarTransportSolver.ScalarTransportSolver(settings).solve()


def test_fenics_tutorial_poisson_published_output():
    """ft01_poisson.py of the FEniCS tutorial through the solver API on the GPU: -Laplace u = -6 on UnitSquareMesh(8, 8), u_D = 1 + x^2 +
    2 y^2 on the whole boundary.  dolfin's published output: error_L2 = 0.00823509807335, error_max = 1.33226762955e-15."""
    from oracle import fem_oracle_p2 as fp
    mesh = UnitSquareMesh(8, 8)
    Q = FunctionSpace(mesh, "P", 1)
    from fenicssolver_b200.dolfin_compat import Expression
    u_D = Expression('1 + x[0]*x[0] + 2*x[1]*x[1]', degree=2)
    boundary = AutoSubDomain(lambda x, on_boundary: on_boundary)
    settings = {'solver_name': 'ScalarTransportSolver', 'scalar_name': 'temperature', 'mesh': None, 'function_space': Q,
                'boundary_conditions': {'all': {'boundary': boundary, 'boundary_id': 1, 'type': 'Dirichlet', 'value': u_D}},
                'body_source': Constant(-6.0), 'initial_values': {'temperature': 0.0},
                'material': {'conductivity': 1.0, 'capacity': 1.0},
                'solver_settings': {'transient_settings': {'transient': False, 'starting_time': 0, 'time_step': 0.1, 'ending_time': 1},
                                    'reference_values': {'temperature': 0.0}, 'solver_parameters': {}},
                'report_settings': {'logging_level': 40, 'logging_file': None, 'plotting_freq': 0, 'saving_freq': 0}}
    u = ScalarTransportSolver.ScalarTransportSolver(settings).solve()
    c, t = fo.unit_square_mesh(8, 8)
    exact = lambda x: 1 + x[..., 0] ** 2 + 2 * x[..., 1] ** 2       # noqa: E731
    assert np.abs(u.compute_vertex_values(mesh) - exact(c)).max() < 1e-11          # error_max (Krylov tolerance instead of LU)
    pts, w = fp._collapsed_rule(2, 6)
    vol, _ = fo.p1_geometry(c, t)
    e = exact(np.einsum('pa,cai->cpi', pts, c[t])) - np.einsum('pa,ca->cp', pts, u.values[t])
    error_L2 = float(np.sqrt(np.sum(vol[:, None] * w[None, :] * e ** 2)))
    assert abs(error_L2 - 0.00823509807335) < 1e-13
