"""Run BASELINE.json configs C3 (elasticity cantilever 128^3) and C4 (transient advection-diffusion 128^3,
200 Crank-Nicolson steps with per-step re-assembly) through the public API and print one JSON line each.
Not the bench (bench.py measures C2); these are the figures quoted in DESIGN.md."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fenicssolver_b200 import LinearElasticitySolver, ScalarTransportSolver  # noqa: E402
from fenicssolver_b200.dolfin_compat import near  # noqa: E402

QUIET = {'logging_level': 40, 'logging_file': None, 'plotting_freq': 0, 'saving_freq': 0, 'plotting_interactive': False}
N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
NSTEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 200
PRECOND = os.environ.get("FSB_PRECOND", "jacobi")        # 'gmg': CG preconditioned by geometric multigrid (C3)


def c3():
    s = {'solver_name': 'LinearElasticitySolver', 'mesh': {'type': 'UnitCubeMesh', 'n': [N, N, N]},
         'material': {'name': 'steel', 'elastic_modulus': 2e11, 'poisson_ratio': 0.27, 'density': 7800},
         'boundary_conditions': {'clamp': {'boundary': lambda x: near(x[0], 0.0), 'boundary_id': 1, 'type': 'Dirichlet', 'value': (0, 0, 0)}},
         'body_source': (0.0, 0.0, -7800 * 9.81), 'initial_values': {},
         'solver_settings': {'transient_settings': {'transient': False, 'starting_time': 0, 'time_step': 0.01, 'ending_time': 0.03},
                             'reference_values': {}, 'solver_parameters': {'preconditioner': PRECOND}},
         'report_settings': QUIET}
    solver = LinearElasticitySolver.LinearElasticitySolver(s)
    solver.device_space().ctx.set_option("profile", 1)
    t0 = time.perf_counter()
    u = solver.solve()
    tip = u.values[np.argmax(solver.mesh.coordinates()[:, 0] + solver.mesh.coordinates()[:, 2])]
    dt = time.perf_counter() - t0
    info = solver.solve_info
    sizes = solver.device_space().A.sizes()
    bytes_spmv = sizes["nnzb"] * 76 + sizes["nrows"] // 3 * 56
    it = max(info["iterations"], 1)
    print(json.dumps({"config": "C3 elasticity cantilever %d^3 P1, %d DoF, %s-CG rtol 1e-12 (reference load sign)" % (N, sizes["nrows"], PRECOND),
                      "iterations": info["iterations"], "converged": info["converged"], "wall_s": dt,
                      "timings": solver.timings, "solve_ms": info["solve_ms"], "ms_per_iteration": info["solve_ms"] / it,
                      "spmv_ms": info["spmv_ms"] / it, "spmv_GBps": (bytes_spmv / (info["spmv_ms"] / it * 1e-3) / 1e9) if info["spmv_ms"] else None,
                      "Mdof_per_s": sizes["nrows"] / (solver.timings["assemble"] + solver.timings["solve"]) / 1e6,
                      "tip_displacement": tip.tolist()}), flush=True)


def c4():
    k, rho, cp = 0.6, 1000.0, 4200.0
    c_ = rho * cp
    h = 1.0 / N
    dt = c_ * h * h / k
    vel = (0.0, 0.0, 2 * k / (c_ * h) * 0.5)                 # cell Peclet 0.5
    s = {'solver_name': 'ScalarTransportSolver', 'scalar_name': 'temperature', 'mesh': {'type': 'UnitCubeMesh', 'n': [N, N, N]},
         'material': {'density': rho, 'specific_heat_capacity': cp, 'thermal_conductivity': k},
         'boundary_conditions': {'hot': {'boundary': lambda x: near(x[2], 0.0), 'boundary_id': 1, 'type': 'Dirichlet', 'value': 360},
                                 'cold': {'boundary': lambda x: near(x[2], 1.0), 'boundary_id': 2, 'type': 'Dirichlet', 'value': 300}},
         'body_source': None, 'initial_values': {'temperature': 300}, 'convective_velocity': vel,
         'solver_settings': {'transient_settings': {'transient': True, 'starting_time': 0.0, 'time_step': dt, 'ending_time': dt * (NSTEPS - 0.5)},
                             'reference_values': {'temperature': 300}, 'solver_parameters': {}},
         'report_settings': QUIET}
    solver = ScalarTransportSolver.ScalarTransportSolver(s)
    iters = []
    orig = solver.solve_current_step

    def counted():
        orig()
        iters.append(solver.solve_info["iterations"])
    solver.solve_current_step = counted
    t0 = time.perf_counter()
    T = solver.solve()
    Th = T.vector().get_local()
    wall = time.perf_counter() - t0
    ndof = Th.size
    print(json.dumps({"config": "C4 transient advection-diffusion %d^3 P1, %d DoF, %d Crank-Nicolson steps, re-assembly every step, BiCGStab rtol 1e-12" % (N, ndof, NSTEPS),
                      "steps": solver.current_step, "wall_s": wall, "ms_per_step": wall / max(solver.current_step, 1) * 1e3,
                      "iterations_per_step": {"first": iters[0], "mean": float(np.mean(iters)), "last": iters[-1]},
                      "Mdof_steps_per_s": ndof * solver.current_step / wall / 1e6, "T_min": float(Th.min()), "T_max": float(Th.max()),
                      "symbolic_s": solver.timings.get("symbolic")}), flush=True)


if __name__ == "__main__":
    which = sys.argv[3] if len(sys.argv) > 3 else "both"
    if which in ("c3", "both"):
        c3()
    if which in ("c4", "both"):
        c4()
