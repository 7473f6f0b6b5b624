"""A/B of the degree-1 scalar matrix assembly on one GPU: atomic scatter (asm_mode 1: one thread per cell, fp64 REDs at the position
map) against row gather (asm_mode 2: one thread per row, no atomics) and the per-warp combine plan (asm_mode 3: sorted contributions,
one RED per distinct slot of a warp), CUDA-event timed, at config C2's mesh size.
    python tools/asm_ab.py [N] [reps]
Prints ms per call for: zero-fill, Laplace, Laplace + mass + advection (config C4's form), and the matrix-free action."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from fenicssolver_b200 import _lib  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 5
stream = torch.cuda.Stream()
ctx = _lib.Context(0, stream=stream.cuda_stream)
m = _lib.DeviceMesh.box(ctx, (N, N, N), (0, 0, 0), (1, 1, 1))
A = _lib.DeviceMatrix.create(m, 1)
nv, nc = m.sizes()[2], m.sizes()[3]
x = _lib.DeviceVector.from_numpy(ctx, np.random.default_rng(0).standard_normal(nv))
y = _lib.DeviceVector(ctx, nv)
vel = np.array([0.0, 0.0, 1e-4])


def timed(fn):
    fn()
    ctx.sync()
    best = 1e30
    for _ in range(REPS):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        ctx.sync()
        best = min(best, a.elapsed_time(b))
    return best


print("N=%d: %d vertices, %d cells, nnz %d" % (N, nv, nc, A.sizes()["nnz"]), flush=True)
print("zero-fill of A: %.3f ms" % timed(A.zero), flush=True)
ref = {}
for mode in (1, 2, 3):
    ctx.set_option("asm_mode", mode)
    name = {1: "scatter (REDs)", 2: "row gather", 3: "warp combine plan"}[mode]
    if mode == 3:
        import time
        t0 = time.perf_counter()
        A.assemble_scalar(kscale=1.0)          # first call in this mode builds the plan (two sort passes over the mesh)
        ctx.sync()
        print("%-16s plan build + first assembly %.1f ms (once per matrix)" % (name, (time.perf_counter() - t0) * 1e3), flush=True)
    for label, kw in (("Laplace", dict(kscale=20.0)), ("Laplace+mass+advection", dict(kscale=0.3, mass=2.5e6, adv=4.2e6, vel=vel))):
        t_add = timed(lambda: A.assemble_scalar(**kw))
        t_set = timed(lambda: A.assemble_scalar(overwrite=True, **kw))
        A.assemble_scalar(overwrite=True, **kw)
        vals = A.download_csr()[2]
        if label in ref:
            d = np.abs(vals - ref[label]).max() / np.abs(ref[label]).max()
        else:
            ref[label], d = vals, 0.0
        print("%-16s %-24s add %.3f ms   set (zero + add) %.3f ms   max rel diff vs scatter %.1e" % (name, label, t_add, t_set, d), flush=True)
    y.fill(0.0)
    print("%-16s %-24s %.3f ms" % (name, "action y += A x", timed(lambda: _lib.apply_scalar(m, x, y, kscale=0.3, mass=2.5e6))), flush=True)
