"""Host profile of one warm config-C3 step (LinearElasticitySolver, 128^3, default solve_amg path): where the time outside the
Krylov solve goes.  python tools/c3_profile.py [N]"""
import cProfile
import io
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from fenicssolver_b200 import LinearElasticitySolver  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
sv = LinearElasticitySolver.LinearElasticitySolver(bench.c3_settings(N, None))
sv.solve()
sv.solve()
t0 = time.perf_counter()
sv.solve()
sv.device_space().ctx.sync()
print("warm step %.1f ms; timings(ms): %s; solve_info: %s" % ((time.perf_counter() - t0) * 1e3, {k: round(v * 1e3, 1) for k, v in sv.timings.items()}, sv.solve_info))
pr = cProfile.Profile()
pr.enable()
sv.solve()
sv.device_space().ctx.sync()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(25)
print(s.getvalue()[:6000])
