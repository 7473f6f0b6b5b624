"""Build libfsb.so (the hand-written sm_100a CUDA library) in-tree with nvcc.

The shared object lands next to its sources (fenicssolver_b200/csrc/libfsb.so) so that it travels
with the repo snapshot to the GPU box; it is git-ignored.  `python -m fenicssolver_b200.build`.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(CSRC, "libfsb.so")
SOURCES = ["fsb_core.cu", "fsb_pattern.cu", "fsb_facets.cu", "fsb_assemble.cu", "fsb_assemble_p2.cu", "fsb_supg.cu", "fsb_spmv.cu", "fsb_squeeze.cu", "fsb_solve.cu", "fsb_cgp.cu", "fsb_mg.cu", "fsb_dist.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; cannot build libfsb.so")
    return exe


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every CUDA translation unit for sm_100a and link libfsb.so.  Returns its path."""
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, "fsb_internal.cuh"), os.path.join(CSRC, "fsb_device.cuh"), os.path.join(CSRC, "fsb_p1.cuh"), os.path.join(CSRC, "fsb_spmv_core.cuh"), os.path.join(CSRC, "fsb_cgp.cuh"), os.path.join(ROOT, "include", "fsb.h")]
    objs, jobs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for log in ex.map(run, jobs):
                if verbose and log:
                    print(log)
    if jobs or _stale(LIB, objs):
        run([nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl", "-lpthread"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
