"""Diagnostics for the multigrid preconditioner: dampings per level, one V-cycle against a numpy restatement, iteration counts."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.sparse as sp
sys.path.insert(0, "/tmp")
from oracle import fem_oracle as fo
from fenicssolver_b200 import ScalarTransportSolver, _lib
from fenicssolver_b200.dolfin_compat import UnitCubeMesh
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import test_gpu_mg as tm

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
solver = ScalarTransportSolver.ScalarTransportSolver(tm.heat_settings(UnitCubeMesh(N, N, N), 'gmg'))
T = solver.solve()
print("info", solver.solve_info)
mg = solver._mg
print("omega", [mg.omega(l) for l in range(len(mg.matrices))])
for l, A in enumerate(mg.matrices):
    rp, ci, va = A.download_csr()
    n = rp.size - 1
    M = sp.csr_matrix((va, ci.astype(np.int64), rp), shape=(n, n))
    nn = int(round(n ** (1 / 3))) - 1
    c, t = fo.unit_cube_mesh(nn, nn, nn)
    fv, opp, _ = fo.exterior_facets(t); mid = c[fv].mean(axis=1)
    inlet = np.nonzero(c[:, 2] == 0)[0]
    Ao, bo = fo.heat_system(c, t, 20.0, [(inlet, 350.0)], source=1000.0, neumann=[(fv[mid[:, 0] == 0], 2000.0)], robin=[(fv[mid[:, 2] == 1], 400.0, 300.0)])
    print("level", l, "n", n, "matrix diff vs oracle", abs(M - Ao).max() / abs(Ao).max())
ctx = solver.device_space().ctx
n0 = mg.matrices[0].sizes()["nrows"]
rng = np.random.default_rng(0)
r = rng.standard_normal(n0)
c, t = fo.unit_cube_mesh(N, N, N)
r[c[:, 2] == 0] = 0.0
z = _lib.DeviceVector(ctx, n0)
mg.apply(_lib.DeviceVector.from_numpy(ctx, r), z, 2)
np.save("gpurun_out/mg_z.npy", z.numpy()); np.save("gpurun_out/mg_r.npy", r)
print("z norm", np.linalg.norm(z.numpy()))
