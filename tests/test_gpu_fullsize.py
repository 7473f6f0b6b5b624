"""GPU tests at BASELINE.json's full single-GPU sizes (configs[1..3]), through the public API, checked with
size-independent properties because the oracle cannot factorise these systems in seconds:

* C2  3D heat 256^3: nodally exact 1-D profile (SURVEY 8c KAT 4), closed-form nnz, integer-exact pattern rows;
* C3  elasticity 128^3, 3 dofs per node: constant-strain patch test (a linear displacement field is reproduced);
* C4  advection-diffusion 128^3: the steady solution is a fixed point of the Crank-Nicolson step
      (A T* = b(T*) when (K + cC) T* = loads), which exercises mass, stiffness, advection, the matrix-free
      right-hand side, both Dirichlet variants, CG and BiCGStab with per-step re-assembly.
Tolerance 1e-10 relative L2 (north_star)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from fenicssolver_b200 import LinearElasticitySolver, ScalarTransportSolver  # noqa: E402
from fenicssolver_b200.dolfin_compat import Function, near  # noqa: E402

QUIET = {'logging_level': 40, 'logging_file': None, 'plotting_freq': 0, 'saving_freq': 0, 'plotting_interactive': False}


def rel_l2(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(b))


def test_c2_heat_256_exact_profile_and_pattern():
    N = 256
    s = {'solver_name': 'ScalarTransportSolver', 'scalar_name': 'temperature', 'mesh': {'type': 'UnitCubeMesh', 'n': [N, N, N]},
         'material': {'density': 1000, 'specific_heat_capacity': 500, 'thermal_conductivity': 20},
         'boundary_conditions': {'inlet': {'boundary': lambda x: near(x[2], 0.0), 'boundary_id': 1, 'type': 'Dirichlet', 'value': 350},
                                 'outlet': {'boundary': lambda x: near(x[2], 1.0), 'boundary_id': 2, 'type': 'Dirichlet', 'value': 300}},
         'body_source': 1000, 'initial_values': {'temperature': 293},
         'solver_settings': {'transient_settings': {'transient': False, 'starting_time': 0, 'time_step': 0.01, 'ending_time': 0.03},
                             'reference_values': {'temperature': 293}, 'solver_parameters': {'relative_tolerance': 1e-7, 'maximum_iterations': 500}},
         'report_settings': QUIET}
    solver = ScalarTransportSolver.ScalarTransportSolver(s)
    T = solver.solve().vector().get_local()
    nv = (N + 1) ** 3
    z = (np.arange(nv) // ((N + 1) ** 2)) / N
    assert rel_l2(T, 350 - 50 * z + 1000 * z * (1 - z) / 40) < 1e-10
    assert solver.solve_info["converged"] == 1
    sizes = solver.device_space().A.sizes()
    assert sizes["nrows"] == nv == 16974593
    assert sizes["nnz"] == nv + 2 * (3 * N * (N + 1) ** 2 + 3 * N * N * (N + 1) + N ** 3) == 253036801
    rp, ci, _ = solver.device_space().A.download_csr(values=False)
    p = N + 1
    r = 100 + 100 * p + 100 * p * p                       # an interior row: the 15 structural neighbours, sorted
    # the six tets per hex share the v0-v7 diagonal: edges run along 1, p, p^2, 1+p, p+p^2, 1+p^2 and 1+p+p^2
    dirs = [1, p, p * p, 1 + p, p + p * p, 1 + p * p, 1 + p + p * p]
    got = (ci[rp[r]:rp[r + 1]] - r).tolist()
    assert got == sorted([0] + dirs + [-d for d in dirs])
    assert np.all(np.diff(rp) >= 4) and np.diff(rp).max() == 15


def test_c3_elasticity_128_patch_test():
    N = 128
    G = np.array([[0.010, 0.020, -0.010], [0.000, -0.020, 0.030], [0.015, 0.000, 0.010]]) * 1e-3
    t0 = np.array([1e-4, -2e-4, 3e-4])
    exprs = tuple("%r*x[0] + %r*x[1] + %r*x[2] + %r" % (float(G[i, 0]), float(G[i, 1]), float(G[i, 2]), float(t0[i])) for i in range(3))
    s = {'solver_name': 'LinearElasticitySolver', 'mesh': {'type': 'UnitCubeMesh', 'n': [N, N, N]},
         'material': {'name': 'steel', 'elastic_modulus': 2e11, 'poisson_ratio': 0.27, 'density': 7800},
         'boundary_conditions': {'all': {'boundary': lambda x, on_boundary: on_boundary, 'boundary_id': 1, 'type': 'Dirichlet', 'value': exprs}},
         'body_source': None, 'initial_values': {},
         'solver_settings': {'transient_settings': {'transient': False, 'starting_time': 0, 'time_step': 0.01, 'ending_time': 0.03},
                             'reference_values': {}, 'solver_parameters': {}},
         'report_settings': QUIET}
    solver = LinearElasticitySolver.LinearElasticitySolver(s)
    u = solver.solve()
    X = solver.mesh.coordinates()
    assert u.values.shape == (2146689, 3)
    assert rel_l2(u.values, X @ G.T + t0) < 1e-10
    sizes = solver.device_space().A.sizes()
    assert sizes["bs"] == 3 and sizes["nnzb"] == 31802497 and sizes["nnz"] == 286222473 and sizes["nrows"] == 6440067
    assert solver.solve_info["converged"] == 1


def test_c4_advection_diffusion_128_steady_state_is_a_crank_nicolson_fixed_point():
    N = 128
    k, rho, cp = 0.6, 1000.0, 4200.0
    c_ = rho * cp
    vel = (2.0 * k / c_, -1.0 * k / c_, 3.0 * k / c_)            # global Peclet ~ 4: the steady problem is well posed
    dt = c_ / (N * N) / k                                        # c h^2 / k
    def settings(transient, initial):
        return {'solver_name': 'ScalarTransportSolver', 'scalar_name': 'temperature', 'mesh': {'type': 'UnitCubeMesh', 'n': [N, N, N]},
                'material': {'density': rho, 'specific_heat_capacity': cp, 'thermal_conductivity': k},
                'boundary_conditions': {'hot': {'boundary': lambda x: near(x[2], 0.0), 'boundary_id': 1, 'type': 'Dirichlet', 'value': 360},
                                        'cold': {'boundary': lambda x: near(x[2], 1.0), 'boundary_id': 2, 'type': 'Dirichlet', 'value': 300},
                                        'side': {'boundary': lambda x: near(x[0], 0.0), 'boundary_id': 3, 'type': 'heatFlux', 'value': 5.0}},
                'body_source': 20.0, 'initial_values': {'temperature': initial}, 'convective_velocity': vel,
                'solver_settings': {'transient_settings': {'transient': transient, 'starting_time': 0.0, 'time_step': dt, 'ending_time': 1.5 * dt},
                                    'reference_values': {'temperature': 300}, 'solver_parameters': {}},
                'report_settings': QUIET}
    steady = ScalarTransportSolver.ScalarTransportSolver(settings(False, 300))
    Ts = steady.solve().vector().get_local()
    assert steady.solve_info["converged"] == 1 and Ts.min() > 299 and Ts.max() < 400
    tr = ScalarTransportSolver.ScalarTransportSolver(settings(True, Ts))
    T2 = tr.solve().vector().get_local()
    assert tr.current_step == 2                                   # two Crank-Nicolson steps, matrix re-assembled in each
    assert rel_l2(T2, Ts) < 1e-10
