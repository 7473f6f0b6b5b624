"""Per-step wall-clock breakdown of config C4 (transient advection-diffusion, re-assembly every step)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tools.config_runs as cr  # noqa: E402


def main():
    from fenicssolver_b200 import ScalarTransportSolver
    acc = {}
    orig = ScalarTransportSolver.ScalarTransportSolver.solve_current_step

    def timed(self):
        t0 = time.perf_counter()
        orig(self)
        self.device_space().ctx.sync()
        dt = time.perf_counter() - t0
        acc.setdefault("step", []).append(dt)
        for k in ("assemble", "solve"):
            acc.setdefault(k, []).append(self.timings.get(k, 0.0))
        acc.setdefault("iters", []).append(self.solve_info["iterations"])
        acc.setdefault("solve_ms_dev", []).append(self.solve_info["solve_ms"])
    ScalarTransportSolver.ScalarTransportSolver.solve_current_step = timed
    cr.c4()
    for k, v in acc.items():
        v = np.array(v[3:])
        print("%-14s mean %.3f  min %.3f  max %.3f" % (k, v.mean() * (1e3 if k in ("step", "assemble", "solve") else 1), v.min() * (1e3 if k in ("step", "assemble", "solve") else 1), v.max() * (1e3 if k in ("step", "assemble", "solve") else 1)))


if __name__ == "__main__":
    main()
