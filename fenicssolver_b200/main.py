"""JSON / dict case loader and solver dispatcher (the reference's main.py:65-95 API).

    from fenicssolver_b200.main import load_settings, main
    main('case.json')            # or main(settings_dict)
"""
from __future__ import annotations

import json
import os.path
import sys


def load_settings(case_input):
    if isinstance(case_input, (dict)):
        settings = case_input
    elif isinstance(case_input, (str, bytes, os.PathLike)) and os.path.exists(case_input):
        with open(case_input, 'r') as f:
            settings = json.loads(f.read())
    else:
        raise TypeError('{} is not supported by Fenics as case input, only path string or dict'.format(type(case_input)))
    return settings


def main(case_input):
    settings = load_settings(case_input)
    solver_name = settings['solver_name']
    if solver_name == "ScalarTransportSolver":
        from . import ScalarTransportSolver
        solver = ScalarTransportSolver.ScalarTransportSolver(settings)
        solver.solve()
    elif solver_name == "LinearElasticitySolver":
        from . import LinearElasticitySolver
        solver = LinearElasticitySolver.LinearElasticitySolver(settings)
        solver.solve()
    elif solver_name == "CoupledNavierStokesSolver":
        # dispatched by the reference (main.py:80-83); a mixed saddle-point system, outside this hot path
        from ._lib import SolverError
        raise SolverError('CoupledNavierStokesSolver is outside the B200 hot path (see DESIGN.md)')
    else:
        raise NameError('Solver name : {} is not supported by Fenics'.format(solver_name))
    solver.plot()
    return solver


if __name__ == "__main__":
    if len(sys.argv) < 2:
        print("Not enough input argument, Usage: `python -m fenicssolver_b200 case_input`")
    else:
        main(sys.argv[1])
