"""Where the host time of bench.py's end-to-end step goes: wall-clock marks around every phase of
construct -> solve -> get_local -> teardown, then a cProfile of one more step.  python tools/e2e_profile.py [N]"""
import cProfile
import gc
import io
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    import torch
    from fenicssolver_b200 import ScalarTransportSolver, backend
    from fenicssolver_b200.dolfin_compat import UnitCubeMesh
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    torch.cuda.set_device(0)
    ctx = backend.get_context(0)
    hmesh = UnitCubeMesh(N, N, N)
    c, t = hmesh.coordinates(), hmesh.cells()
    pc = torch.empty(c.shape, dtype=torch.float64, pin_memory=True)
    pt = torch.empty(t.shape, dtype=torch.int32, pin_memory=True)
    pc.numpy()[:] = c
    pt.numpy()[:] = t
    hmesh._coords, hmesh._cells = pc.numpy(), pt.numpy()
    hmesh.force_upload = True
    hmesh.exterior_facets()

    def step(marks=None):
        m = [("start", time.perf_counter())]
        sv = ScalarTransportSolver.ScalarTransportSolver(bench.case_settings(N, mesh=hmesh))
        m.append(("construct", time.perf_counter()))
        T = sv.solve()
        m.append(("solve()", time.perf_counter()))
        out = T.vector().get_local()
        m.append(("get_local", time.perf_counter()))
        tim = dict(sv.timings)
        del sv, T
        m.append(("del solver", time.perf_counter()))
        gc.collect()
        m.append(("gc.collect", time.perf_counter()))
        ctx.sync()
        m.append(("ctx.sync", time.perf_counter()))
        if marks is not None:
            marks.append(([(k, (b - a) * 1e3) for (_, a), (k, b) in zip(m[:-1], m[1:])], tim))
        return out

    step()
    marks = []
    for _ in range(3):
        t0 = time.perf_counter()
        step(marks)
        print("step total %.1f ms" % ((time.perf_counter() - t0) * 1e3))
    for mk, tim in marks:
        print("  ".join("%s %.1f" % kv for kv in mk), "| timings(ms):", {k: round(v * 1e3, 1) for k, v in tim.items()})
    pr = cProfile.Profile()
    pr.enable()
    step()
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
    print(s.getvalue()[:6000])


if __name__ == "__main__":
    main()
