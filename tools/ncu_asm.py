"""Minimal run for ncu: one P1 tetrahedron matrix assembly per asm_mode given on the command line.
    ncu --set full --clock-control none --import-source on -k regex:'k_scalar_form' -o out python tools/ncu_asm.py 256 1 3"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from fenicssolver_b200 import _lib  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
modes = [int(a) for a in sys.argv[2:]] or [1, 3]
ctx = _lib.Context(0)
m = _lib.DeviceMesh.box(ctx, (N, N, N), (0, 0, 0), (1, 1, 1))
A = _lib.DeviceMatrix.create(m, 1)
vel = np.array([0.0, 0.0, 1e-4])
for mode in modes:
    ctx.set_option("asm_mode", mode)
    for _ in range(2):
        A.zero()
        A.assemble_scalar(kscale=0.3, mass=2.5e6, adv=4.2e6, vel=vel)
    ctx.sync()
print("done")
