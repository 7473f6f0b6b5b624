"""Where the host time of bench.py's end-to-end step goes: wall-clock marks around every phase of
construct -> solve -> get_local -> release, then a cProfile of one more step (rank 0).
    python tools/e2e_profile.py [N]                                                       # one GPU: plain array mesh
    python -m torch.distributed.run --nproc-per-node 2 ... tools/e2e_profile.py [N]        # slab-distributed: host arrays + box description"""
import cProfile
import io
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    from fenicssolver_b200 import ScalarTransportSolver, backend
    from fenicssolver_b200.dolfin_compat import Mesh, UnitCubeMesh
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = backend.get_context(local)
    hmesh = UnitCubeMesh(N, N, N)
    c, t = hmesh.coordinates(), hmesh.cells()
    pc = torch.empty(c.shape, dtype=torch.float64, pin_memory=True)
    pt = torch.empty(t.shape, dtype=torch.int32, pin_memory=True)
    pc.numpy()[:] = c
    pt.numpy()[:] = t
    if world == 1:
        hmesh = Mesh(pc.numpy(), pt.numpy(), cells_sorted=True)
    else:
        hmesh._coords, hmesh._cells = pc.numpy(), pt.numpy()
        hmesh.force_upload = True

    def step(marks=None):
        m = [("start", time.perf_counter())]
        hmesh._exterior = None
        for k in ("_dmesh", "_slab", "_boundary_geometry"):
            hmesh.__dict__.pop(k, None)
        m.append(("reset", time.perf_counter()))
        sv = ScalarTransportSolver.ScalarTransportSolver(bench.case_settings(N, mesh=hmesh, distributed=world > 1))
        m.append(("construct", time.perf_counter()))
        T = sv.solve()
        m.append(("solve()", time.perf_counter()))
        out = sv.local_result() if world > 1 else T.vector().get_local()
        m.append(("result", time.perf_counter()))
        tim = dict(sv.timings)
        del sv, T
        m.append(("release", time.perf_counter()))
        ctx.sync()
        m.append(("ctx.sync", time.perf_counter()))
        if world > 1:
            dist.barrier()
            m.append(("barrier", time.perf_counter()))
        if marks is not None:
            marks.append(([(k, (b - a) * 1e3) for (_, a), (k, b) in zip(m[:-1], m[1:])], tim))
        return out

    step()
    marks = []
    for _ in range(3):
        t0 = time.perf_counter()
        step(marks)
        if rank == 0:
            print("step total %.1f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
    for r in range(world):
        if r == rank:
            for mk, tim in marks:
                print("rank %d: " % rank + "  ".join("%s %.1f" % kv for kv in mk), "| timings(ms):", {k: round(v * 1e3, 1) for k, v in tim.items()}, flush=True)
        if world > 1:
            dist.barrier()
    pr = cProfile.Profile()
    pr.enable()
    step()
    pr.disable()
    if rank == 0:
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
        print(s.getvalue()[:7000], flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
