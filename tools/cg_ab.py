"""A/B of the CG variants and the SpMV L2 hints on config C2's system (single GPU): time per iteration and per SpMV
over a fixed number of iterations."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fenicssolver_b200 import _lib, backend  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ITERS = int(sys.argv[2]) if len(sys.argv) > 2 else 200
NZ = int(sys.argv[3]) if len(sys.argv) > 3 else N          # N x N x NZ cells: NZ = N/8 is one rank's slab of an 8-GPU run
VARIANTS = [int(v) for v in sys.argv[4].split(",")] if len(sys.argv) > 4 else [1, 2, 3]
ctx = backend.get_context()
ctx.set_option("profile", 1)
if os.environ.get("FSB_CG_UMODE"):
    ctx.set_option("cg_umode", int(os.environ["FSB_CG_UMODE"]))
m = _lib.DeviceMesh.box(ctx, (N, N, NZ), (0, 0, 0), (1, 1, NZ / N))
nv = (N + 1) ** 2 * (NZ + 1)
A = _lib.DeviceMatrix.create(m, 1)
A.assemble_scalar(kscale=20.0)
b = _lib.DeviceVector(ctx, nv)
_lib.assemble_source(m, b, 1000.0)
plane = (N + 1) ** 2
dofs = np.concatenate([np.arange(plane), np.arange(nv - plane, nv)])
vals = np.concatenate([np.full(plane, 350.0), np.full(plane, 300.0)])
x = _lib.DeviceVector(ctx, nv)
A.apply_dirichlet(b, dofs, vals, symmetric=True, x=x)
nnz = A.sizes()["nnz"]
for variant in VARIANTS:
    for hint in (1,):
        for dz in (0, 1):
            ctx.set_option("cg_variant", variant)
            ctx.set_option("spmv_hint", hint)
            ctx.set_option("drop_zeros", dz)
            best = None
            for _ in range(2):
                x.fill(293.0)
                A.apply_dirichlet(b, dofs, vals, symmetric=True, x=x) if False else None
                info = A.solve(b, x, "cg", rtol=1e-30, maxit=ITERS)
                it = max(info["iterations"], 1)
                cur = (info["solve_ms"] / it, info["spmv_ms"] / it)
                best = cur if best is None or cur[0] < best[0] else best
            gb = (12 * info["operand_nnzb"] + 24 * nv) / 1e9
            print("N=%d NZ=%d variant=%d hint=%d drop_zeros=%d : iteration %.4f ms  spmv %.4f ms (%.0f GB/s)  vector part %.4f ms"
                  % (N, NZ, variant, hint, dz, best[0], best[1], gb / best[1] * 1e3, best[0] - best[1]), flush=True)
