"""Single GPU: does the update phase of the persistent CG kernel depend on where the owned range starts (a rank with a lower
ghost plane starts one plane = an odd number of doubles into its vectors) and on the relative placement of the work vectors?
Emulates rank 1 of a 2-GPU 256^3 run on one GPU by restricting the owned rows of a 256x256x128 problem."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fenicssolver_b200 import _lib, backend  # noqa: E402

N, NZ, ITERS = 256, int(sys.argv[1]) if len(sys.argv) > 1 else 128, 200
ctx = backend.get_context()
ctx.set_option("profile", 1)
plane = (N + 1) ** 2
for ghost_lo, ghost_hi in ((0, 1), (1, 0)):
    for skew in (0, 4, 8, 16):
        ctx.set_option("cg_debug", skew)
        m = _lib.DeviceMesh.box(ctx, (N, N, NZ), (0, 0, 0), (1, 1, NZ / N))
        nv = plane * (NZ + 1)
        A = _lib.DeviceMatrix.create(m, 1)
        A.set_owned_rows(ghost_lo * plane, nv - ghost_hi * plane)
        A.assemble_scalar(kscale=20.0)
        b = _lib.DeviceVector(ctx, nv)
        _lib.assemble_source(m, b, 1000.0)
        x = _lib.DeviceVector(ctx, nv)
        for variant in (3,):
            ctx.set_option("cg_variant", variant)
            best = None
            for _ in range(2):
                x.fill(293.0)
                info = A.solve(b, x, "cg", rtol=1e-30, maxit=ITERS)
                it = max(info["iterations"], 1)
                cur = (info["solve_ms"] / it, info["spmv_ms"] / it)
                best = cur if best is None or cur[0] < best[0] else best
            print("ghost_lo=%d ghost_hi=%d debug=%4d variant=%d: iteration %.1f us, SpMV %.1f us, rest %.1f us"
                  % (ghost_lo, ghost_hi, skew, variant, best[0] * 1e3, best[1] * 1e3, (best[0] - best[1]) * 1e3), flush=True)
        del A, b, x, m
