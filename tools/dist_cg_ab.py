"""Run under torchrun on N GPUs: time per CG iteration of the distributed solve of config C2's system (z-slabs) for the
peer-memory persistent kernel, the two-kernel peer-memory chain and the NCCL chain, over a fixed number of iterations.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/dist_cg_ab.py [size] [iterations]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import bench
    from fenicssolver_b200 import ScalarTransportSolver
    from fenicssolver_b200.SolverBase import collect_dirichlet
    solver = ScalarTransportSolver.ScalarTransportSolver(bench.case_settings(N, distributed=True))
    solver.init_solver()
    solver.current_step = 0
    F, bcs = solver.generate_form(0, None, None, solver.w_current, solver.w_prev)
    dofs, vals = collect_dirichlet(bcs, solver.mesh)
    space = solver.device_space()
    ctx = space.ctx
    ctx.set_option("profile", 1)
    x = space.vector()
    b, _ = F.assemble(space)
    space.apply_dirichlet(b, dofs, vals, symmetric=True, x=x)
    cases = [("persistent kernel, peer memory", {"cg_variant": 0, "dist_p2p": 1}),
             ("two-kernel chain, peer memory", {"cg_variant": 2, "dist_p2p": 1}),
             ("two-kernel chain, NCCL", {"cg_variant": 2, "dist_p2p": 0})]
    for extra in os.environ.get("FSB_AB_EXTRA", "").split(";"):
        if extra:
            name, kv = extra.split(":")
            cases.append((name, {k: int(v) for k, v in (p.split("=") for p in kv.split(","))}))
    for name, opts in cases:
        for k, v in opts.items():
            ctx.set_option(k, v)
        best = None
        for _ in range(3):
            x.fill(293.0)
            torch.cuda.synchronize()
            dist.barrier()
            info = space.solve(b, x, method="cg", rtol=1e-30, maxit=iters)
            t = torch.tensor([info["solve_ms"]], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item()) / max(info["iterations"], 1)
            best = ms if best is None else min(best, ms)
        if rank == 0:
            print("N=%d world=%d %-36s : %.1f us per iteration (%d iterations, max over ranks, best of 3)" % (N, world, name, best * 1e3, iters), flush=True)
        for k in opts:
            ctx.set_option(k, {"cg_variant": 0, "dist_p2p": 1}.get(k, 0))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
