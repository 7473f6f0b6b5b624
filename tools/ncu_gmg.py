"""One multigrid-preconditioned solve of config C2 at size N through the solver API, for an ncu launch list
(ncu --metrics gpu__time_duration.sum -c 3000 python tools/ncu_gmg.py 256)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from fenicssolver_b200 import ScalarTransportSolver  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
s = bench.case_settings(N)
s['solver_settings']['solver_parameters']['preconditioner'] = 'gmg'
solver = ScalarTransportSolver.ScalarTransportSolver(s)
solver.solve()
print(solver.solve_info, solver.timings)
