"""ctypes loader + driver for oracle/fem_oracle_c.c (CPU baseline; TEST/BENCH INFRASTRUCTURE ONLY).

`heat_cube(N, ...)` runs the whole 3D-heat pipeline of config C2 on the host cores and reports the
time of each phase; bench.py uses it for `cpu_baseline` and `--impl reference`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libfem_oracle.so")
_lib = None


def build(force=False):
    """gcc -O3 -fopenmp; -march=native is dropped from the cached .so's name on purpose: the build
    container and the GPU box may differ, so the library is rebuilt there if loading fails."""
    src = os.path.join(HERE, "fem_oracle_c.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        cmd = ["gcc", "-O3", "-fopenmp", "-fPIC", "-std=c11", "-shared", "-o", LIB, src, "-lm"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle C build failed:\n" + r.stderr)
    return LIB


def load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(LIB)
        vp, i64, dbl = C.c_void_p, C.c_int64, C.c_double
        lib.fo_num_threads.restype = C.c_int
        lib.fo_box_mesh.argtypes = [vp, vp, vp, vp, vp]
        lib.fo_vertex_cells.argtypes = [i64, i64, vp, vp, vp]
        lib.fo_csr_pattern.restype = i64
        lib.fo_csr_pattern.argtypes = [i64, vp, vp, vp, vp, vp]
        lib.fo_assemble_heat.argtypes = [i64, vp, vp, dbl, dbl, vp, vp, vp, vp]
        lib.fo_apply_dirichlet_sym.argtypes = [i64, vp, vp, vp, vp, vp, vp]
        lib.fo_spmv.argtypes = [i64, vp, vp, vp, vp, vp]
        lib.fo_pcg_jacobi.restype = C.c_int
        lib.fo_pcg_jacobi.argtypes = [i64, vp, vp, vp, vp, vp, dbl, dbl, C.c_int, C.POINTER(dbl)]
        _lib = lib
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def num_threads():
    return load().fo_num_threads()


def use_all_cores():
    """Use every core this process may run on, whatever OMP_NUM_THREADS says (torchrun sets it to 1 in every rank)."""
    import os
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    load().fo_set_num_threads(int(n))
    return num_threads()


def box_mesh(n, p0=(0, 0, 0), p1=(1, 1, 1)):
    lib = load()
    n_ = np.asarray(n, dtype=np.int32)
    nv = int(np.prod(n_ + 1))
    nc = 6 * int(np.prod(n_))
    coords = np.empty((nv, 3))
    cells = np.empty((nc, 4), dtype=np.int32)
    lib.fo_box_mesh(_p(n_), _p(np.asarray(p0, dtype=np.float64)), _p(np.asarray(p1, dtype=np.float64)), _p(coords), _p(cells))
    return coords, cells


def csr_pattern(cells, nverts):
    lib = load()
    cells = np.ascontiguousarray(cells, dtype=np.int32)
    vptr = np.empty(nverts + 1, dtype=np.int64)
    v2c = np.empty(cells.size, dtype=np.int32)
    lib.fo_vertex_cells(nverts, cells.shape[0], _p(cells), _p(vptr), _p(v2c))
    rp = np.zeros(nverts + 1, dtype=np.int64)
    nnz = lib.fo_csr_pattern(nverts, _p(cells), _p(vptr), _p(v2c), _p(rp), None)
    ci = np.empty(nnz, dtype=np.int32)
    lib.fo_csr_pattern(nverts, _p(cells), _p(vptr), _p(v2c), _p(rp), _p(ci))
    return rp, ci


class HeatCube:
    """Config C2 at size N on the CPU: setup (mesh + pattern, untimed like the GPU arm's symbolic phase)
    then step() = assemble + Dirichlet + Jacobi-PCG, the timed unit."""

    def __init__(self, N, k=20.0, S=1000.0, T0=350.0, T1=300.0, T_init=293.0):
        self.lib = load()
        self.N, self.k, self.S, self.T_init = N, k, S, T_init
        t = time.perf_counter()
        self.coords, self.cells = box_mesh((N, N, N))
        self.nv = self.coords.shape[0]
        self.rp, self.ci = csr_pattern(self.cells, self.nv)
        self.t_setup = time.perf_counter() - t
        p = N + 1
        self.flag = np.zeros(self.nv, dtype=np.uint8)
        self.g = np.zeros(self.nv)
        self.flag[:p * p] = 1
        self.g[:p * p] = T0
        self.flag[-p * p:] = 1
        self.g[-p * p:] = T1
        self.vals = np.zeros(self.ci.size)
        self.b = np.zeros(self.nv)
        self.x = np.zeros(self.nv)

    def step(self, rtol=1e-12, maxit=100000):
        lib = self.lib
        t0 = time.perf_counter()
        self.vals[:] = 0.0
        self.b[:] = 0.0
        lib.fo_assemble_heat(self.cells.shape[0], _p(self.cells), _p(self.coords), self.k, self.S, _p(self.rp), _p(self.ci), _p(self.vals), _p(self.b))
        lib.fo_apply_dirichlet_sym(self.nv, _p(self.rp), _p(self.ci), _p(self.vals), _p(self.b), _p(self.flag), _p(self.g))
        t1 = time.perf_counter()
        self.x[:] = np.where(self.flag, self.g, self.T_init)      # initial field, Dirichlet values imposed
        rel = C.c_double()
        it = lib.fo_pcg_jacobi(self.nv, _p(self.rp), _p(self.ci), _p(self.vals), _p(self.b), _p(self.x), rtol, 0.0, maxit, C.byref(rel))
        t2 = time.perf_counter()
        return {"t_assemble": t1 - t0, "t_solve": t2 - t1, "iterations": it, "relres": rel.value}
