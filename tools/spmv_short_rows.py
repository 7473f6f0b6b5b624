"""SpMV configuration sweep on the squeezed (drop_zeros) operand of config C2: rows per tile x lanes per row x
pipeline stages.  Prints achieved GB/s per configuration (algorithmic bytes 12 B/nnz + 24 B/row of the operand)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fenicssolver_b200 import _lib, backend  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ITERS = 48
ctx = backend.get_context()
ctx.set_option("profile", 1)
m = _lib.DeviceMesh.box(ctx, (N, N, N), (0, 0, 0), (1, 1, 1))
nv = (N + 1) ** 3
A = _lib.DeviceMatrix.create(m, 1)
A.assemble_scalar(kscale=20.0)
b = _lib.DeviceVector(ctx, nv)
_lib.assemble_source(m, b, 1000.0)
plane = (N + 1) ** 2
dofs = np.concatenate([np.arange(plane), np.arange(nv - plane, nv)])
vals = np.concatenate([np.full(plane, 350.0), np.full(plane, 300.0)])
x = _lib.DeviceVector(ctx, nv)
A.apply_dirichlet(b, dofs, vals, symmetric=True, x=x)


def run(dz, rows, lpr, nst):
    ctx.set_option("drop_zeros", dz)
    ctx.set_option("spmv_rows", rows)
    ctx.set_option("spmv_lpr", lpr)
    ctx.set_option("spmv_stages", nst)
    x.fill(293.0)
    try:
        best = None
        for _ in range(2):
            info = A.solve(b, x, "cg", rtol=1e-30, maxit=ITERS)
            ms = info["spmv_ms"] / max(info["iterations"], 1)
            best = ms if best is None else min(best, ms)
        nnz = info["operand_nnzb"]
        gb = (12 * nnz + 24 * nv) / 1e9
        print("drop_zeros=%d rows=%3d lpr=%d stages=%d : %.4f ms  %7.1f GB/s  (nnz %d, it %.4f ms)"
              % (dz, rows, lpr, nst, best, gb / best * 1e3, nnz, info["solve_ms"] / max(info["iterations"], 1)), flush=True)
    except _lib.SolverError as e:
        print("drop_zeros=%d rows=%3d lpr=%d stages=%d : %s" % (dz, rows, lpr, nst, e), flush=True)


run(0, 0, 0, 0)
run(1, 0, 0, 0)
for rows in (512, 256, 128):
    for lpr in (1, 2):
        if rows == 512 and lpr == 2:
            continue
        for nst in (2, 3, 4):
            run(1, rows, lpr, nst)
run(0, 0, 0, 0)
