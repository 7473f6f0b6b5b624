"""Degree-2 at scale: elasticity on a BoxMesh N^3 with VectorFunctionSpace(mesh, "Lagrange", 2) (the reference example's own
space, examples/test_linear_elasticity.py:105-106) and the scalar heat problem on FunctionSpace(mesh, "CG", 2): set-up, assembly and
Krylov timings, SpMV bandwidth.  python tools/p2_run.py [N]"""
import copy
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fenicssolver_b200 import LinearElasticitySolver, ScalarTransportSolver, SolverBase  # noqa: E402
from fenicssolver_b200.dolfin_compat import AutoSubDomain, BoxMesh, Constant, FunctionSpace, Point, UnitCubeMesh, VectorFunctionSpace, near  # noqa: E402

QUIET = {'logging_level': 40, 'logging_file': None, 'plotting_freq': 0, 'saving_freq': 0, 'plotting_interactive': False}
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32


def report(name, solver, wall):
    sp_ = solver.device_space()
    sizes = sp_.A.sizes()
    info = solver.solve_info
    it = max(info["iterations"], 1)
    bs = sizes["bs"]
    bytes_spmv = sizes["nnzb"] * (8 * bs * bs + 4) + (sizes["nrows"] // bs) * (8 + 16 * bs)
    print(json.dumps({"case": name, "dofs": sizes["nrows"], "nnz": sizes["nnz"], "blocks_per_row": sizes["nnzb"] / (sizes["nrows"] / bs),
                      "iterations": info["iterations"], "converged": info["converged"], "wall_s": wall,
                      "timings_s": {k: round(v, 4) for k, v in solver.timings.items()},
                      "ms_per_iteration": info["solve_ms"] / it,
                      "spmv_ms": info["spmv_ms"] / it if info["spmv_ms"] else None,
                      "spmv_GBps": bytes_spmv / (info["spmv_ms"] / it * 1e-3) / 1e9 if info["spmv_ms"] else None}), flush=True)


def elasticity():
    mesh = BoxMesh(Point(0, 0, 0), Point(4, 1, 1), N, N, N)
    s = copy.deepcopy(SolverBase.default_case_settings)
    s.update({'material': {'elastic_modulus': 2e11, 'poisson_ratio': 0.27, 'density': 7800},
              'function_space': VectorFunctionSpace(mesh, "Lagrange", 2), 'report_settings': QUIET,
              'boundary_conditions': {'clamp': {'boundary': AutoSubDomain(lambda x: near(x[0], 0.0)), 'boundary_id': 1, 'type': 'Dirichlet',
                                                'value': Constant((0, 0, 0))}},
              'body_source': (0.0, 0.0, -7800 * 9.81)})
    s['solver_settings'] = dict(s['solver_settings'], solver_parameters={'preconditioner': 'jacobi'})
    solver = LinearElasticitySolver.LinearElasticitySolver(s)
    solver.device_space().ctx.set_option("profile", 1)
    t0 = time.perf_counter()
    solver.solve()
    report("P2 elasticity %d^3" % N, solver, time.perf_counter() - t0)


def heat():
    mesh = UnitCubeMesh(N, N, N)
    s = {'solver_name': 'ScalarTransportSolver', 'scalar_name': 'temperature', 'mesh': None, 'function_space': FunctionSpace(mesh, "CG", 2),
         'material': {'density': 1000, 'specific_heat_capacity': 500, 'thermal_conductivity': 20},
         'boundary_conditions': {'inlet': {'boundary': lambda x: near(x[2], 0.0), 'boundary_id': 1, 'type': 'Dirichlet', 'value': 350},
                                 'outlet': {'boundary': lambda x: near(x[2], 1.0), 'boundary_id': 2, 'type': 'Dirichlet', 'value': 300}},
         'body_source': 1000, 'initial_values': {'temperature': 293},
         'solver_settings': {'transient_settings': {'transient': False, 'starting_time': 0, 'time_step': 0.01, 'ending_time': 0.03},
                             'reference_values': {'temperature': 293}, 'solver_parameters': {}},
         'report_settings': QUIET}
    solver = ScalarTransportSolver.ScalarTransportSolver(s)
    solver.device_space().ctx.set_option("profile", 1)
    t0 = time.perf_counter()
    T = solver.solve()
    z = solver.function_space.node_coordinates()[:, 2]
    err = np.linalg.norm(T.values - (350 - 50 * z + 1000 * z * (1 - z) / 40)) / np.linalg.norm(T.values)
    report("P2 heat %d^3 (quadratic exact profile reproduced to %.1e)" % (N, err), solver, time.perf_counter() - t0)


if __name__ == "__main__":
    heat()
    elasticity()
