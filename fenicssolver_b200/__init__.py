"""B200-native FEM assemble-and-solve behind the FenicsSolver settings API.

    from fenicssolver_b200 import ScalarTransportSolver, LinearElasticitySolver, SolverBase
    from fenicssolver_b200.main import load_settings, main

Unlike the reference's __init__ (which calls main(sys.argv) on import when argv has two entries,
FenicsSolver/__init__.py:12-13, breaking `import` under pytest), the command-line hook lives in
__main__.py: `python -m fenicssolver_b200 case.json`.
"""
__version__ = "0.1"

from ._lib import SolverError  # noqa: F401
from . import SolverBase, ScalarTransportSolver, LinearElasticitySolver  # noqa: F401
from .main import load_settings, main  # noqa: F401
