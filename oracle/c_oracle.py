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
        lib.fo_zero.argtypes = [vp, i64]
        lib.fo_pcg_jacobi_segment.restype = C.c_int
        lib.fo_pcg_jacobi_segment.argtypes = [i64, vp, vp, vp, vp, vp, vp, vp, dbl, dbl, C.c_int]
        lib.fo_assemble_scalar.argtypes = [i64, vp, vp, dbl, dbl, dbl, vp, vp, vp, vp]
        lib.fo_apply_scalar.argtypes = [i64, vp, vp, dbl, dbl, vp, vp]
        lib.fo_apply_dirichlet_nonsym.argtypes = [i64, vp, vp, vp, vp, vp, vp]
        lib.fo_bicgstab_jacobi.restype = C.c_int
        lib.fo_bicgstab_jacobi.argtypes = [i64, vp, vp, vp, vp, vp, dbl, dbl, C.c_int, C.POINTER(dbl)]
        lib.fo_assemble_elasticity.argtypes = [i64, vp, vp, dbl, dbl, vp, vp, vp, vp, vp]
        lib.fo_assemble_heat_p2.argtypes = [i64, vp, vp, vp, vp, dbl, dbl, vp, vp, vp, vp]
        lib.fo_mg_lambda_max.restype = dbl
        lib.fo_mg_lambda_max.argtypes = [i64, vp, vp, vp]
        lib.fo_mg_apply.argtypes = [C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int]
        lib.fo_mg_pcg.restype = C.c_int
        lib.fo_mg_pcg.argtypes = [C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp, dbl, C.c_int, C.c_int, C.c_int, C.POINTER(dbl)]
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
        self.vals = np.empty(self.ci.size)          # first touched (and zeroed every step) by fo_zero, in parallel
        self.b = np.empty(self.nv)
        self.x = np.empty(self.nv)
        for a in (self.vals, self.b, self.x):
            self.lib.fo_zero(_p(a), a.size)

    def step(self, rtol=1e-12, maxit=100000):
        lib = self.lib
        t0 = time.perf_counter()
        lib.fo_zero(_p(self.vals), self.vals.size)
        lib.fo_zero(_p(self.b), self.b.size)
        lib.fo_assemble_heat(self.cells.shape[0], _p(self.cells), _p(self.coords), self.k, self.S, _p(self.rp), _p(self.ci), _p(self.vals), _p(self.b))
        lib.fo_apply_dirichlet_sym(self.nv, _p(self.rp), _p(self.ci), _p(self.vals), _p(self.b), _p(self.flag), _p(self.g))
        t1 = time.perf_counter()
        self.x[:] = np.where(self.flag, self.g, self.T_init)      # initial field, Dirichlet values imposed
        rel = C.c_double()
        it = lib.fo_pcg_jacobi(self.nv, _p(self.rp), _p(self.ci), _p(self.vals), _p(self.b), _p(self.x), rtol, 0.0, maxit, C.byref(rel))
        t2 = time.perf_counter()
        return {"t_assemble": t1 - t0, "t_solve": t2 - t1, "iterations": it, "relres": rel.value}

    # ---- the same step cut into consecutive segments (bench.py --impl reference: the driver's K steps are K segments of ONE
    # complete step, every second of it measured, none extrapolated)
    def begin_step(self):
        """Segment 0's first part: zero, assemble, Dirichlet, start vector.  Returns the seconds it took."""
        lib = self.lib
        t0 = time.perf_counter()
        if getattr(self, "_work", None) is None:
            self._work = np.empty(4 * self.nv)
            lib.fo_zero(_p(self._work), self._work.size)        # parallel first touch
        lib.fo_zero(_p(self.vals), self.vals.size)
        lib.fo_zero(_p(self.b), self.b.size)
        lib.fo_assemble_heat(self.cells.shape[0], _p(self.cells), _p(self.coords), self.k, self.S, _p(self.rp), _p(self.ci), _p(self.vals), _p(self.b))
        lib.fo_apply_dirichlet_sym(self.nv, _p(self.rp), _p(self.ci), _p(self.vals), _p(self.b), _p(self.flag), _p(self.g))
        self.x[:] = np.where(self.flag, self.g, self.T_init)
        self._state = np.zeros(8)
        return time.perf_counter() - t0

    def solve_segment(self, iters, rtol=1e-12):
        """Up to `iters` more CG iterations of the step begun by begin_step().  Returns (seconds, converged, iterations so far)."""
        t0 = time.perf_counter()
        done = self.lib.fo_pcg_jacobi_segment(self.nv, _p(self.rp), _p(self.ci), _p(self.vals), _p(self.b), _p(self.x), _p(self._work),
                                              _p(self._state), rtol, 0.0, int(iters))
        return time.perf_counter() - t0, bool(done), int(self._state[4])

    def exact_profile(self):
        """Nodally exact solution of this problem (SURVEY 8c KAT 4): T(z) = T0 + (T1 - T0) z + S z (1 - z) / (2 k)."""
        z = self.coords[:, 2]
        return self.g[0] + (self.g[-1] - self.g[0]) * z + self.S * z * (1 - z) / (2 * self.k)


class MultigridLevels:
    """ctypes argument pack for fo_mg_apply / fo_mg_pcg: per-level CSR arrays, constrained-dof flags, vertices per axis and
    eigenvalue estimates (computed with fo_mg_lambda_max unless given)."""

    def __init__(self, levels, lmax=None):
        """levels: list of dicts with rp (int64), ci (int32), va (float64), bc (uint8 flags), dims (vertices per axis, 3 ints)."""
        lib = load()
        self.keep = levels
        nl = len(levels)
        self.nlevels = nl
        self.n = np.array([L["rp"].size - 1 for L in levels], dtype=np.int64)
        self.dims = np.array([list(L["dims"]) + [1] * (3 - len(L["dims"])) for L in levels], dtype=np.int32)
        ptrs = lambda key: (C.c_void_p * nl)(*[L[key].ctypes.data for L in levels])      # noqa: E731
        self.rp, self.ci, self.va, self.bc = ptrs("rp"), ptrs("ci"), ptrs("va"), ptrs("bc")
        if lmax is None:
            lmax = [lib.fo_mg_lambda_max(int(self.n[l]), _p(L["rp"]), _p(L["ci"]), _p(L["va"])) for l, L in enumerate(levels)]
        self.lmax = np.array(lmax, dtype=np.float64)

    def apply(self, r, nu=2, coarse_sweeps=24):
        z = np.empty_like(r)
        load().fo_mg_apply(self.nlevels, _p(self.n), _p(self.dims), self.rp, self.ci, self.va, self.bc, _p(self.lmax), _p(r), _p(z), nu, coarse_sweeps)
        return z

    def pcg(self, b, x, rtol=1e-12, maxit=1000, nu=2, coarse_sweeps=24):
        rel = C.c_double()
        it = load().fo_mg_pcg(self.nlevels, _p(self.n), _p(self.dims), self.rp, self.ci, self.va, self.bc, _p(self.lmax), _p(b), _p(x),
                              rtol, maxit, nu, coarse_sweeps, C.byref(rel))
        return it, rel.value


class HeatCubeMG:
    """Config C2 on the CPU with CG preconditioned by the geometric multigrid of csrc/fsb_mg.cu: one HeatCube per level (N, N/2, ...
    while even and >= 4); step() = assemble + Dirichlet on every level, then the multigrid PCG.  The eigenvalue estimates are made
    once (first step) and reused, as the GPU path does."""

    def __init__(self, N, **kw):
        sizes = [N]
        while sizes[-1] % 2 == 0 and sizes[-1] >= 4:
            sizes.append(sizes[-1] // 2)
        t = time.perf_counter()
        self.cubes = [HeatCube(n, **kw) for n in sizes]
        self.t_setup = time.perf_counter() - t
        self.lmax = None

    def step(self, rtol=1e-12, maxit=1000):
        lib = load()
        t0 = time.perf_counter()
        for h in self.cubes:
            lib.fo_zero(_p(h.vals), h.vals.size)
            lib.fo_zero(_p(h.b), h.b.size)
            lib.fo_assemble_heat(h.cells.shape[0], _p(h.cells), _p(h.coords), h.k, h.S, _p(h.rp), _p(h.ci), _p(h.vals), _p(h.b))
            lib.fo_apply_dirichlet_sym(h.nv, _p(h.rp), _p(h.ci), _p(h.vals), _p(h.b), _p(h.flag), _p(h.g))
        levels = [{"rp": h.rp, "ci": h.ci, "va": h.vals, "bc": h.flag, "dims": (h.N + 1,) * 3} for h in self.cubes]
        mg = MultigridLevels(levels, self.lmax)
        self.lmax = mg.lmax
        t1 = time.perf_counter()
        f = self.cubes[0]
        f.x[:] = np.where(f.flag, f.g, f.T_init)
        it, rel = mg.pcg(f.b, f.x, rtol=rtol, maxit=maxit)
        t2 = time.perf_counter()
        return {"iterations": it, "relres": rel, "t_assemble": t1 - t0, "t_solve": t2 - t1, "x": f.x, "levels": len(levels)}


def expand_pattern(rp, ci, ncomp):
    """Scalar CSR pattern of a vector space with dof = ncomp * node + component from the node pattern (columns stay sorted)."""
    lens = np.diff(rp)
    rp3 = np.zeros(ncomp * lens.size + 1, dtype=np.int64)
    np.cumsum(np.repeat(lens * ncomp, ncomp), out=rp3[1:])
    ci3 = np.empty(int(rp3[-1]), dtype=np.int32)
    cols = (ci.astype(np.int64)[:, None] * ncomp + np.arange(ncomp)).astype(np.int32)          # [nnz, ncomp]
    for r in range(lens.size):                      # small meshes only; the sized path below builds it block-wise
        blk = cols[rp[r]:rp[r + 1]].ravel()
        for i in range(ncomp):
            ci3[rp3[ncomp * r + i]:rp3[ncomp * r + i + 1]] = blk
    return rp3, ci3


def expand_pattern_fast(rp, ci, ncomp):
    """The same without a Python loop over rows (vectorised; used at bench sizes)."""
    lens = np.diff(rp)
    n = lens.size
    rp3 = np.zeros(ncomp * n + 1, dtype=np.int64)
    np.cumsum(np.repeat(lens * ncomp, ncomp), out=rp3[1:])
    cols = (ci.astype(np.int64)[:, None] * ncomp + np.arange(ncomp)).astype(np.int32).reshape(-1)     # row-major: entry k -> ncomp columns
    # every node row r contributes its block of ncomp*len columns ncomp times in a row
    starts = rp[:-1] * ncomp
    out = np.empty(int(rp3[-1]), dtype=np.int32)
    row_of = np.repeat(np.arange(n), lens * ncomp)                    # node row of each expanded column entry
    offs = np.arange(cols.size) - np.repeat(starts, lens * ncomp)     # position inside the node row's expanded block
    for i in range(ncomp):
        out[rp3[ncomp * row_of + i] + offs] = cols
    return rp3, out


class ElasticityCube:
    """Config C3 on the CPU: unit cube N^3 P1, 3 dofs per node, clamp on x = 0, body force (0, 0, -rho g) with the reference's
    load sign (a(u, v) = -L(v)); step() = assemble + symmetric Dirichlet + Jacobi-PCG on the scalar CSR."""

    def __init__(self, N, E=2e11, nu=0.27, rho=7800.0):
        self.lib = load()
        self.N = N
        self.mu, self.lam = E / (2 * (1 + nu)), E * nu / ((1 + nu) * (1 - 2 * nu))
        self.f = np.array([0.0, 0.0, -rho * 9.81]) * -1.0          # reference sign flip (LinearElasticitySolver.py:242-243)
        t = time.perf_counter()
        self.coords, self.cells = box_mesh((N, N, N))
        self.nv = self.coords.shape[0]
        rp, ci = csr_pattern(self.cells, self.nv)
        self.rp, self.ci = expand_pattern_fast(rp, ci, 3)
        self.n = 3 * self.nv
        self.t_setup = time.perf_counter() - t
        clamp = np.nonzero(self.coords[:, 0] == 0.0)[0]
        self.flag = np.zeros(self.n, dtype=np.uint8)
        self.flag[(3 * clamp[:, None] + np.arange(3)).ravel()] = 1
        self.g = np.zeros(self.n)
        self.vals = np.empty(self.ci.size)
        self.b = np.empty(self.n)
        self.x = np.empty(self.n)
        for a in (self.vals, self.b, self.x):
            self.lib.fo_zero(_p(a), a.size)

    def step(self, rtol=1e-12, maxit=100000):
        lib = self.lib
        t0 = time.perf_counter()
        lib.fo_zero(_p(self.vals), self.vals.size)
        lib.fo_zero(_p(self.b), self.b.size)
        lib.fo_assemble_elasticity(self.cells.shape[0], _p(self.cells), _p(self.coords), self.mu, self.lam, _p(self.f), _p(self.rp), _p(self.ci),
                                   _p(self.vals), _p(self.b))
        lib.fo_apply_dirichlet_sym(self.n, _p(self.rp), _p(self.ci), _p(self.vals), _p(self.b), _p(self.flag), _p(self.g))
        t1 = time.perf_counter()
        lib.fo_zero(_p(self.x), self.x.size)
        rel = C.c_double()
        it = lib.fo_pcg_jacobi(self.n, _p(self.rp), _p(self.ci), _p(self.vals), _p(self.b), _p(self.x), rtol, 0.0, maxit, C.byref(rel))
        t2 = time.perf_counter()
        return {"t_assemble": t1 - t0, "t_solve": t2 - t1, "iterations": it, "relres": rel.value}


class TransientCube:
    """Config C4 on the CPU: 3D transient advection-diffusion on the unit cube N^3 P1, Crank-Nicolson (theta = 0.5) with the
    convection term fully implicit as the reference has it (ScalarTransportSolver.py:292-293, 311), matrix re-assembled every step,
    plain bc.apply + Jacobi-BiCGStab.  step() advances one time step; T holds the current field."""

    def __init__(self, N, k=0.6, rho=1000.0, cp=4200.0, T_hot=360.0, T_cold=300.0, T_init=300.0):
        self.lib = load()
        self.N, self.k = N, k
        self.c = rho * cp
        h = 1.0 / N
        self.dt = self.c * h * h / k
        self.vel = np.array([0.0, 0.0, 2 * k / (self.c * h) * 0.5])          # cell Peclet number 0.5
        self.coords, self.cells = box_mesh((N, N, N))
        self.nv = self.coords.shape[0]
        self.rp, self.ci = csr_pattern(self.cells, self.nv)
        p = (N + 1) ** 2
        self.flag = np.zeros(self.nv, dtype=np.uint8)
        self.g = np.zeros(self.nv)
        self.flag[:p] = 1; self.g[:p] = T_hot
        self.flag[-p:] = 1; self.g[-p:] = T_cold
        self.vals = np.empty(self.ci.size)
        self.b = np.empty(self.nv)
        self.T = np.empty(self.nv)
        for a in (self.vals, self.b, self.T):
            self.lib.fo_zero(_p(a), a.size)
        self.T[:] = T_init
        self.iterations = []

    def step(self, rtol=1e-12, maxit=100000, theta=0.5):
        lib = self.lib
        nc = self.cells.shape[0]
        t0 = time.perf_counter()
        lib.fo_zero(_p(self.vals), self.vals.size)
        lib.fo_zero(_p(self.b), self.b.size)
        lib.fo_assemble_scalar(nc, _p(self.cells), _p(self.coords), theta * self.k, self.c / self.dt, self.c, _p(self.vel), _p(self.rp), _p(self.ci), _p(self.vals))
        lib.fo_apply_scalar(nc, _p(self.cells), _p(self.coords), -(1.0 - theta) * self.k, self.c / self.dt, _p(self.T), _p(self.b))
        lib.fo_apply_dirichlet_nonsym(self.nv, _p(self.rp), _p(self.ci), _p(self.vals), _p(self.b), _p(self.flag), _p(self.g))
        t1 = time.perf_counter()
        self.T[self.flag.astype(bool)] = self.g[self.flag.astype(bool)]      # start vector: previous field with the Dirichlet values imposed
        rel = C.c_double()
        it = lib.fo_bicgstab_jacobi(self.nv, _p(self.rp), _p(self.ci), _p(self.vals), _p(self.b), _p(self.T), rtol, 0.0, maxit, C.byref(rel))
        t2 = time.perf_counter()
        self.iterations.append(it)
        return {"t_assemble": t1 - t0, "t_solve": t2 - t1, "iterations": it, "relres": rel.value}


class HeatCubeP2:
    """The degree-2 heat problem of the bench's `p2` block on the CPU: unit cube N^3, P2 tetrahedra, Dirichlet 350 / 300 on z = 0 / 1, source S;
    set-up (edge numbering, node table, pattern: numpy oracle, untimed like the GPU arm's symbolic phase), then step() = C/OpenMP assembly
    (fo_assemble_heat_p2 with the numpy oracle's exact reference tensors) + symmetric Dirichlet + Jacobi-PCG."""

    def __init__(self, N, k=20.0, S=1000.0, T0=350.0, T1=300.0, T_init=293.0):
        from . import fem_oracle_p2 as p2
        self.lib = load()
        self.N, self.k, self.S, self.T_init = N, k, S, T_init
        t = time.perf_counter()
        coords, cells = box_mesh((N, N, N))
        self.cell_nodes, self.node_coords, _ = p2.p2_dofmap(coords, cells)
        self.cell_nodes = np.ascontiguousarray(self.cell_nodes, dtype=np.int32)
        self.coords = coords
        self.n = self.node_coords.shape[0]
        rp, ci = p2.csr_pattern(self.cell_nodes, self.n)
        self.rp, self.ci = np.ascontiguousarray(rp, dtype=np.int64), np.ascontiguousarray(ci, dtype=np.int32)
        R, _, _, F = p2.reference_tensors(3)
        self.R, self.F = np.ascontiguousarray(R, dtype=np.float64), np.ascontiguousarray(F, dtype=np.float64)
        self.t_setup = time.perf_counter() - t
        z = self.node_coords[:, 2]
        self.flag = ((z == 0.0) | (z == 1.0)).astype(np.uint8)
        self.g = np.where(z == 0.0, T0, np.where(z == 1.0, T1, 0.0))
        self.vals = np.empty(self.ci.size)
        self.b = np.empty(self.n)
        self.x = np.empty(self.n)
        for a in (self.vals, self.b, self.x):
            self.lib.fo_zero(_p(a), a.size)

    def step(self, rtol=1e-12, maxit=100000):
        lib = self.lib
        t0 = time.perf_counter()
        lib.fo_zero(_p(self.vals), self.vals.size)
        lib.fo_zero(_p(self.b), self.b.size)
        lib.fo_assemble_heat_p2(self.cell_nodes.shape[0], _p(self.cell_nodes), _p(self.coords), _p(self.R), _p(self.F), self.k, self.S,
                                _p(self.rp), _p(self.ci), _p(self.vals), _p(self.b))
        lib.fo_apply_dirichlet_sym(self.n, _p(self.rp), _p(self.ci), _p(self.vals), _p(self.b), _p(self.flag), _p(self.g))
        t1 = time.perf_counter()
        self.x[:] = np.where(self.flag, self.g, self.T_init)
        rel = C.c_double()
        it = lib.fo_pcg_jacobi(self.n, _p(self.rp), _p(self.ci), _p(self.vals), _p(self.b), _p(self.x), rtol, 0.0, maxit, C.byref(rel))
        t2 = time.perf_counter()
        z = self.node_coords[:, 2]
        exact = self.g[np.argmin(z)] + (self.g[np.argmax(z)] - self.g[np.argmin(z)]) * z + self.S * z * (1 - z) / (2 * self.k)
        err = float(np.linalg.norm(self.x - exact) / np.linalg.norm(exact))
        return {"t_assemble": t1 - t0, "t_solve": t2 - t1, "iterations": it, "relres": rel.value, "rel_l2_vs_exact": err}
