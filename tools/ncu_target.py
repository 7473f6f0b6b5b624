"""Minimal production-config run for ncu: 3D heat at size N, assemble once, CG for a few iterations."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fenicssolver_b200 import _lib

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
ctx = _lib.Context(0)
m = _lib.DeviceMesh.box(ctx, (N, N, N), (0, 0, 0), (1, 1, 1))
if os.environ.get("FSB_NCU_K1"):
    fv, opp, cell, fid = m.exterior_facets()          # K1: boundary search + facet ids (set-up launch list)
    m.boundary_geometry()
A = _lib.DeviceMatrix.create(m, 1)
nv = m.sizes()[2]
A.assemble_scalar(kscale=20.0)
b = _lib.DeviceVector(ctx, nv)
x = _lib.DeviceVector(ctx, nv)
x.fill(293.0)
_lib.assemble_source(m, b, 1000.0)
p = N + 1
z0 = np.arange(p * p, dtype=np.int64)
dofs = np.concatenate([z0, z0 + p * p * N])
vals = np.concatenate([np.full(z0.size, 350.0), np.full(z0.size, 300.0)])
A.apply_dirichlet(b, dofs, vals, True, x)
print(A.solve(b, x, "cg", rtol=1e-12, maxit=iters))
