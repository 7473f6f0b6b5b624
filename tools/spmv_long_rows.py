"""SpMV configuration sweep on degree-2 operands (long rows: ~28 entries / blocks per row on average, up to ~90): rows per tile x
lanes per row x stages x tile form (flat = two-phase: products over the tile's non-zeros, then row sums; rows = LPR lanes per
row), scalar CSR (P2 heat) and 3x3 BSR (P2 elasticity); every configuration's y is compared with the plain kernel's.
    python tools/spmv_long_rows.py [N]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fenicssolver_b200 import _lib, backend  # noqa: E402
from fenicssolver_b200.dolfin_compat import FunctionSpace, UnitCubeMesh  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 48
ctx = backend.get_context()
ctx.set_option("profile", 1)
mesh = UnitCubeMesh(N, N, N)
V = FunctionSpace(mesh, "CG", 2)
m = _lib.DeviceMesh.upload_p2(ctx, mesh.coordinates(), V.cell_nodes(), V.num_nodes())
nn = V.num_nodes()
for bs in (1, 3):
    A = _lib.DeviceMatrix.create(m, bs)
    if bs == 1:
        A.assemble_scalar(kscale=1.0, mass=1.0)
    else:
        A.assemble_elasticity(1.0, 1.5)
        xyz = V.node_coordinates()
        d = np.nonzero(xyz[:, 0] == 0)[0]
        dofs = (d[:, None] * 3 + np.arange(3)).ravel()
    s = A.sizes()
    b, x = _lib.DeviceVector(ctx, nn * bs), _lib.DeviceVector(ctx, nn * bs)
    b.fill(1.0)
    if bs == 3:
        A.apply_dirichlet(b, dofs, np.zeros(dofs.size), symmetric=True, x=x)
    gb = (s["nnzb"] * (8 * bs * bs + 4) + nn * (8 + 16 * bs)) / 1e9
    print("bs=%d rows=%d blocks/row=%.1f bytes/SpMV=%.3f GB" % (bs, s["nrows"], s["nnzb"] / nn, gb), flush=True)
    xr = _lib.DeviceVector.from_numpy(ctx, np.random.default_rng(0).standard_normal(nn * bs))
    yr, yt = _lib.DeviceVector(ctx, nn * bs), _lib.DeviceVector(ctx, nn * bs)
    ctx.set_option("spmv_mode", 1)
    A.spmv(xr, yr)
    ctx.set_option("spmv_mode", 0)
    yref = yr.numpy()
    combos = [(0, 0, 0, 0)] + ([(r, l, n, f) for f in (2, 1) for r, l in ((256, 2), (128, 2), (128, 4)) for n in (2, 3)] if bs == 1 else
                               [(r, l, n, f) for f in (2, 1) for r, l in ((256, 2), (128, 2), (128, 4)) for n in (2, 3)])
    for rows, lpr, nst, flat in combos:
        ctx.set_option("spmv_rows", rows); ctx.set_option("spmv_lpr", lpr); ctx.set_option("spmv_stages", nst); ctx.set_option("spmv_flat", flat)
        try:
            A.spmv(xr, yt)
            err = np.abs(yt.numpy() - yref).max() / np.abs(yref).max()
            best = None
            for _ in range(2):
                x.fill(0.0)
                info = A.solve(b, x, "cg", rtol=1e-30, maxit=40)
                ms = info["spmv_ms"] / max(info["iterations"], 1)
                best = ms if best is None else min(best, ms)
            print("  rows=%3d lpr=%d stages=%d %s : %.4f ms %7.1f GB/s   max rel diff vs plain kernel %.1e"
                  % (rows, lpr, nst, {0: "auto", 1: "flat", 2: "rows"}[flat], best, gb / best * 1e3, err), flush=True)
        except _lib.SolverError as e:
            print("  rows=%3d lpr=%d stages=%d flat=%d : %s" % (rows, lpr, nst, flat, str(e)[:80]), flush=True)
    ctx.set_option("spmv_rows", 0); ctx.set_option("spmv_lpr", 0); ctx.set_option("spmv_stages", 0); ctx.set_option("spmv_flat", 0)
    del A
