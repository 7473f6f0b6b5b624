"""Run under torchrun on >= 2 GPUs: distributed (z-slab) solve of the 3D heat problem through the solver
API, compared on rank 0 with a single-GPU solve of the same problem and with the exact profile."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import bench
    from fenicssolver_b200 import ScalarTransportSolver, _lib, backend
    s = bench.case_settings(N, distributed=True)
    s['solver_settings']['gather_result'] = True
    solver = ScalarTransportSolver.ScalarTransportSolver(s)
    t0 = time.perf_counter()
    T = solver.solve()
    dt = time.perf_counter() - t0
    xd = T.vector().get_local()
    info = solver.solve_info
    # same problem again through NCCL collectives instead of the peer-memory mailboxes / halo
    ctx = solver.device_space().ctx
    ctx.set_option("dist_p2p", 0)
    s2 = bench.case_settings(N, distributed=True)
    s2['solver_settings']['gather_result'] = True
    solver_nccl = ScalarTransportSolver.ScalarTransportSolver(s2)
    t0 = time.perf_counter()
    xn = solver_nccl.solve().vector().get_local()
    dtn = time.perf_counter() - t0
    ctx.set_option("dist_p2p", 1)
    if rank == 0:
        print("peer-memory path: %d iterations %.3fs | NCCL path: %d iterations %.3fs | rel diff %.2e"
              % (info["iterations"], dt, solver_nccl.solve_info["iterations"], dtn, np.linalg.norm(xd - xn) / np.linalg.norm(xn)), flush=True)
    ok = bool(np.linalg.norm(xd - xn) <= 1e-10 * np.linalg.norm(xn))
    if rank == 0:
        z = (np.arange((N + 1) ** 3) // ((N + 1) ** 2)) / N
        exact = 350 - 50 * z + 1000 * z * (1 - z) / 40
        err = np.linalg.norm(xd - exact) / np.linalg.norm(exact)
        # single-GPU solve on a separate context of the same device
        ctx1 = _lib.Context(local)
        s1 = bench.case_settings(N, distributed=False)
        solver1 = ScalarTransportSolver.ScalarTransportSolver(s1)
        solver1._space = backend.DeviceSpace(solver1.mesh, 1, ctx=ctx1)
        x1 = solver1.solve().vector().get_local()
        d = np.linalg.norm(xd - x1) / np.linalg.norm(x1)
        print("N=%d world=%d: iters dist=%d single=%d  rel_l2(dist vs exact)=%.2e  rel_l2(dist vs single)=%.2e  wall=%.3fs"
              % (N, world, info["iterations"], solver1.solve_info["iterations"], err, d, dt), flush=True)
        ok = ok and err < 1e-10 and d < 1e-10 and info["converged"] == 1
        # the concatenated owned row blocks reproduce the global CSR pattern exactly
    # transient run with gather_result=False: every rank keeps only its part of the solution on the device between the steps
    # (the T_prev term and the Krylov start vector of the next step read it there); compared with the gathered run
    def transient_settings(gather):
        st = bench.case_settings(N, distributed=True)
        st['solver_settings']['gather_result'] = gather
        st['solver_settings']['transient_settings'] = {'transient': True, 'starting_time': 0.0, 'time_step': 50.0, 'ending_time': 175.0}
        return st
    tr_local = ScalarTransportSolver.ScalarTransportSolver(transient_settings(False))
    tr_local.solve()
    mine_local = tr_local.local_result().copy()
    tr_gath = ScalarTransportSolver.ScalarTransportSolver(transient_settings(True))
    xg = tr_gath.solve().vector().get_local()
    spc = tr_local.device_space()
    ref_slice = xg[spc.v_off + spc.own_v0:spc.v_off + spc.own_v1]
    dtr = float(np.linalg.norm(mine_local - ref_slice) / np.linalg.norm(ref_slice))
    moved = float(np.abs(ref_slice - 293.0).max())
    t = torch.tensor([dtr, -moved], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("transient, %d steps, gather_result=False vs True: rel diff %.2e (field moved by up to %.1f)" % (tr_local.current_step, t[0].item(), -t[1].item()), flush=True)
    ok = ok and t[0].item() < 1e-12 and tr_local.current_step >= 3
    # multigrid-preconditioned CG with the fine level on z-slabs and the coarse hierarchy replicated: the same cycle as on one GPU, so
    # the same iteration count and the same solution (heat, and the elasticity cantilever through its default solve_amg path)
    from fenicssolver_b200 import LinearElasticitySolver
    Nm = N if N % 4 == 0 else 4 * (N // 4)
    sg = bench.case_settings(Nm, distributed=True)
    sg['solver_settings']['gather_result'] = True
    sg['solver_settings']['solver_parameters']['preconditioner'] = 'gmg'
    mgd = ScalarTransportSolver.ScalarTransportSolver(sg)
    xmg = mgd.solve().vector().get_local()
    el = bench.c3_settings(Nm, None)
    el['solver_settings']['distributed'] = True
    el['solver_settings']['gather_result'] = True
    eld = LinearElasticitySolver.LinearElasticitySolver(el)
    xel = eld.solve().vector().get_local()
    if rank == 0:
        s1g = bench.case_settings(Nm, distributed=False)
        s1g['solver_settings']['solver_parameters']['preconditioner'] = 'gmg'
        mg1 = ScalarTransportSolver.ScalarTransportSolver(s1g)
        mg1._space = backend.DeviceSpace(mg1.mesh, 1, ctx=ctx1)
        x1g = mg1.solve().vector().get_local()
        el1 = LinearElasticitySolver.LinearElasticitySolver(bench.c3_settings(Nm, None))
        el1._space = backend.DeviceSpace(el1.mesh, 3, ctx=ctx1, space=el1.function_space)
        x1e = el1.solve().vector().get_local()
        dg = float(np.linalg.norm(xmg - x1g) / np.linalg.norm(x1g))
        de = float(np.linalg.norm(xel - x1e) / np.linalg.norm(x1e))
        zc = (np.arange((Nm + 1) ** 3) // ((Nm + 1) ** 2)) / Nm
        eg = float(np.linalg.norm(xmg - (350 - 50 * zc + 1000 * zc * (1 - zc) / 40)) / np.linalg.norm(xmg))
        print("distributed multigrid-CG %d^3: heat %d iterations (%d levels; one GPU %d) rel diff %.2e, vs exact %.2e | elasticity (solve_amg default) %d "
              "iterations (one GPU %d) rel diff %.2e" % (Nm, mgd.solve_info["iterations"], mgd.solve_info.get("mg_levels", 0), mg1.solve_info["iterations"], dg, eg,
                                                         eld.solve_info["iterations"], el1.solve_info["iterations"], de), flush=True)
        ok = ok and dg < 1e-10 and de < 1e-9 and eg < 1e-10 and mgd.solve_info["converged"] == 1 and eld.solve_info["converged"] == 1
        ok = ok and abs(mgd.solve_info["iterations"] - mg1.solve_info["iterations"]) <= 1 and abs(eld.solve_info["iterations"] - el1.solve_info["iterations"]) <= 1
        ok = ok and eld.solve_info.get("mg_levels", 0) >= 2
    rp, ci, va = solver.device_space().A.download_csr()
    sp_ = solver.device_space()
    lo, hi = sp_.own_v0, sp_.own_v1
    mine = (np.diff(rp)[lo:hi], ci[rp[lo]:rp[hi]] + sp_.v_off, va[rp[lo]:rp[hi]])
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    if rank == 0:
        lens = np.concatenate([p[0] for p in parts]); cols = np.concatenate([p[1] for p in parts]); vals = np.concatenate([p[2] for p in parts])
        rp1, ci1, va1 = solver1.device_space().A.download_csr()
        same_pat = np.array_equal(np.concatenate([[0], np.cumsum(lens)]), rp1) and np.array_equal(cols, ci1)
        dv = np.abs(vals - va1).max() / np.abs(va1).max()
        print("global CSR from owned row blocks: pattern exact=%s  max rel value diff=%.2e" % (same_pat, dv), flush=True)
        ok = ok and same_pat and dv < 1e-13
        print("DIST_CHECK_OK" if ok else "DIST_CHECK_FAILED", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
