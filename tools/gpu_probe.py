"""Quick device probe: time the phases of the 3D heat problem at size N (not the bench; for tuning)."""
import sys
import time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fenicssolver_b200 import _lib


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    rtol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-12
    maxit = int(sys.argv[3]) if len(sys.argv) > 3 else 100000
    ctx = _lib.Context(0)
    def timed(label, fn, bytes_=None):
        ctx.sync(); t = time.perf_counter(); r = fn(); ctx.sync(); dt = time.perf_counter() - t
        extra = "  %.1f GB/s" % (bytes_ / dt / 1e9) if bytes_ else ""
        print("%-28s %9.3f ms%s" % (label, dt * 1e3, extra), flush=True)
        return r
    m = timed("mesh_box", lambda: _lib.DeviceMesh.box(ctx, (N, N, N), (0, 0, 0), (1, 1, 1)))
    _, _, nv, nc = m.sizes()
    A = timed("mat_create (symbolic)", lambda: _lib.DeviceMatrix.create(m, 1))
    s = A.sizes(); nnz = s["nnz"]
    print("N=%d nverts=%d ncells=%d nnz=%d" % (N, nv, nc, nnz))
    b_asm = 16 * nc + 24 * nv + 8 * nnz
    for mode in (1, 0, 1):
        ctx.set_option("asm_mode", mode)
        A.zero()
        timed("assemble laplace mode %d" % mode, lambda: A.assemble_scalar(kscale=20.0), b_asm)
    b = _lib.DeviceVector(ctx, nv); x = _lib.DeviceVector(ctx, nv)
    timed("assemble source", lambda: _lib.assemble_source(m, b, 1000.0))
    p = N + 1
    z0 = np.arange(p * p, dtype=np.int64); z1 = z0 + p * p * N
    dofs = np.concatenate([z0, z1]); vals = np.concatenate([np.full(z0.size, 350.0), np.full(z1.size, 300.0)])
    timed("dirichlet (symmetric)", lambda: A.apply_dirichlet(b, dofs, vals, True, x))
    y = _lib.DeviceVector(ctx, nv)
    b_spmv = 12 * nnz + 24 * nv
    def time_spmv(A, x, y, bytes_, label):
        combos = [(2, 256, 2, 2), (1, 256, 2, 2)] + [(0, r, l, s) for r in (256, 128) for l in (1, 2, 4, 8) for s in (2, 3)]
        for mode, rows, lpr, nst in combos:
            ctx.set_option("spmv_mode", mode); ctx.set_option("spmv_rows", rows)
            ctx.set_option("spmv_lpr", lpr); ctx.set_option("spmv_stages", nst)
            try:
                A.spmv(x, y)
            except _lib.SolverError:
                continue          # combination not instantiated for this block size
            ctx.sync(); t = time.perf_counter()
            for _ in range(20): A.spmv(x, y)
            ctx.sync(); dt = (time.perf_counter() - t) / 20
            print("%s spmv mode %d rows %3d lpr %d stages %d   %9.3f ms  %.1f GB/s" % (label, mode, rows, lpr, nst, dt * 1e3, bytes_ / dt / 1e9), flush=True)
        ctx.set_option("spmv_mode", 0); ctx.set_option("spmv_rows", 0); ctx.set_option("spmv_lpr", 0); ctx.set_option("spmv_stages", 0)
    time_spmv(A, x, y, b_spmv, "csr ")
    if os.environ.get("PROBE_ELASTICITY", "1") == "1" and N <= 160:
        A3 = timed("mat_create bs=3", lambda: _lib.DeviceMatrix.create(m, 3))
        s3 = A3.sizes()
        timed("assemble elasticity", lambda: A3.assemble_elasticity(7.9e10, 9.2e10), 16 * nc + 24 * nv + 8 * s3["nnz"])
        x3 = _lib.DeviceVector(ctx, 3 * nv); y3 = _lib.DeviceVector(ctx, 3 * nv); x3.fill(1.0)
        time_spmv(A3, x3, y3, s3["nnzb"] * 76 + nv * (8 + 24 + 24), "bsr3")
        del A3, x3, y3
    for prof in (0, 1):
        ctx.set_option("profile", prof)
        x.fill(0.0); A.apply_dirichlet(b, dofs, vals, True, x)
        t = time.perf_counter(); info = A.solve(b, x, "cg", rtol=rtol, maxit=maxit); dt = time.perf_counter() - t
        it = max(info["iterations"], 1)
        print("cg profile=%d: %s wall %.1f ms, %.3f ms/iter, iter GB/s %.1f, spmv avg %.3f ms -> %.1f GB/s" % (
            prof, info, dt * 1e3, info["solve_ms"] / it, (b_spmv + 88 * nv) / (info["solve_ms"] / it * 1e-3) / 1e9,
            info["spmv_ms"] / it, b_spmv / max(info["spmv_ms"] / it * 1e-3, 1e-12) / 1e9), flush=True)
    xh = x.numpy()
    zc = (np.arange(nv) // (p * p)) / N
    ex = 350 - 50 * zc + 1000 * zc * (1 - zc) / 40
    print("rel L2 vs analytic profile: %.3e" % (np.linalg.norm(xh - ex) / np.linalg.norm(ex)))
    print("launches", ctx.launch_count(), ctx.device_info())


if __name__ == "__main__":
    main()
