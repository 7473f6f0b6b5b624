"""ctypes binding of libfsb.so (include/fsb.h) and thin RAII wrappers over its handles.

The product path has no CPU fallback: if the shared library is missing or a call fails, SolverError
is raised (the reference's error type, /root/reference/FenicsSolver/SolverBase.py:61).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np


class SolverError(Exception):
    pass


_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libfsb.so")

c_i32, c_i64, c_dbl, c_vp = C.c_int32, C.c_int64, C.c_double, C.c_void_p
P = C.POINTER


class SolveInfo(C.Structure):
    _fields_ = [("iterations", c_i32), ("converged", c_i32), ("rnorm", c_dbl), ("bnorm", c_dbl),
                ("solve_ms", c_dbl), ("spmv_ms", c_dbl), ("operand_nnzb", c_i64)]


# name -> (restype, argtypes); must list every symbol include/fsb.h declares (tests check this)
SIGNATURES = {
    "fsb_init": (C.c_int, [C.c_int, c_vp, P(c_vp)]),
    "fsb_destroy": (None, [c_vp]),
    "fsb_last_error": (C.c_char_p, [c_vp]),
    "fsb_sync": (C.c_int, [c_vp]),
    "fsb_device_info": (C.c_int, [c_vp, P(c_i32), P(c_i64), P(c_i64)]),
    "fsb_set_option": (C.c_int, [c_vp, C.c_char_p, c_i64]),
    "fsb_launch_count": (c_i64, [c_vp]),
    "fsb_mesh_upload": (C.c_int, [c_vp, c_i32, c_i32, c_i64, c_vp, c_i64, c_vp, P(c_vp)]),
    "fsb_mesh_upload_part": (C.c_int, [c_vp, c_i32, c_i32, c_i64, c_vp, c_i64, c_vp, c_i64, P(c_vp)]),
    "fsb_mesh_box": (C.c_int, [c_vp, c_i32, c_vp, c_vp, c_vp, c_i32, c_i32, P(c_vp)]),
    "fsb_mesh_upload_p2": (C.c_int, [c_vp, c_i32, c_i32, c_i64, c_vp, c_i64, c_vp, c_i64, P(c_vp)]),
    "fsb_mesh_sizes": (C.c_int, [c_vp, P(c_i32), P(c_i32), P(c_i64), P(c_i64)]),
    "fsb_mesh_download": (C.c_int, [c_vp, c_vp, c_vp]),
    "fsb_mesh_exterior_facets": (C.c_int, [c_vp, P(c_i64), P(c_i64)]),
    "fsb_mesh_exterior_facets_get": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp]),
    "fsb_mesh_boundary_geometry": (C.c_int, [c_vp, P(c_i64), c_vp, c_vp, c_vp, c_vp]),
    "fsb_mesh_destroy": (None, [c_vp]),
    "fsb_vec_create": (C.c_int, [c_vp, c_i64, P(c_vp)]),
    "fsb_vec_fill": (C.c_int, [c_vp, c_dbl]),
    "fsb_vec_upload": (C.c_int, [c_vp, c_vp, c_i64]),
    "fsb_vec_download": (C.c_int, [c_vp, c_vp, c_i64]),
    "fsb_host_alloc": (C.c_int, [c_vp, c_i64, P(c_vp)]),
    "fsb_host_free": (C.c_int, [c_vp, c_vp]),
    "fsb_vec_copy": (C.c_int, [c_vp, c_vp]),
    "fsb_vec_axpy": (C.c_int, [c_vp, c_dbl, c_vp]),
    "fsb_vec_add_entries": (C.c_int, [c_vp, c_i64, c_vp, c_vp]),
    "fsb_vec_size": (C.c_int, [c_vp, P(c_i64)]),
    "fsb_vec_ptr": (c_vp, [c_vp]),
    "fsb_vec_destroy": (None, [c_vp]),
    "fsb_mat_create": (C.c_int, [c_vp, c_i32, P(c_vp)]),
    "fsb_mat_from_csr": (C.c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, P(c_vp)]),
    "fsb_mat_sizes": (C.c_int, [c_vp, P(c_i64), P(c_i64), P(c_i32), P(c_i64)]),
    "fsb_mat_download_csr": (C.c_int, [c_vp, c_vp, c_vp, c_vp]),
    "fsb_mat_zero": (C.c_int, [c_vp]),
    "fsb_mat_set_owned_rows": (C.c_int, [c_vp, c_i64, c_i64]),
    "fsb_mat_destroy": (None, [c_vp]),
    "fsb_assemble_scalar": (C.c_int, [c_vp, c_vp, c_dbl, c_vp, c_dbl, c_dbl, c_vp]),
    "fsb_assemble_scalar_set": (C.c_int, [c_vp, c_vp, c_dbl, c_vp, c_dbl, c_dbl, c_vp]),
    "fsb_apply_scalar": (C.c_int, [c_vp, c_vp, c_vp, c_dbl, c_vp, c_dbl, c_dbl, c_vp]),
    "fsb_assemble_elasticity": (C.c_int, [c_vp, c_vp, c_dbl, c_dbl]),
    "fsb_assemble_source": (C.c_int, [c_vp, c_vp, c_i32, c_vp, c_dbl, c_vp, c_i32]),
    "fsb_assemble_source_nodal": (C.c_int, [c_vp, c_vp, c_i32, c_vp, c_dbl]),
    "fsb_assemble_facet_load": (C.c_int, [c_vp, c_vp, c_i32, c_i64, c_vp, c_vp, c_i32, c_vp, c_dbl]),
    "fsb_assemble_facet_mass": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_dbl]),
    "fsb_facet_area": (C.c_int, [c_vp, c_i64, c_vp, P(c_dbl)]),
    "fsb_assemble_thermal_load": (C.c_int, [c_vp, c_vp, c_dbl, c_vp, c_dbl, c_dbl, c_dbl]),
    "fsb_assemble_von_mises_load": (C.c_int, [c_vp, c_vp, c_dbl, c_dbl, c_vp]),
    "fsb_assemble_facet_radiation": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_dbl, c_dbl, c_dbl]),
    "fsb_assemble_scalar_nonlinear_k": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_dbl, c_dbl]),
    "fsb_assemble_advection_nodal": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_dbl]),
    "fsb_assemble_scalar_supg": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_dbl, c_dbl, c_vp, c_dbl]),
    "fsb_assemble_source_supg": (C.c_int, [c_vp, c_vp, c_dbl, c_vp, c_dbl, c_vp, c_i32]),
    "fsb_assemble_facet_supg": (C.c_int, [c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_dbl, c_dbl, c_vp, c_dbl]),
    "fsb_apply_dirichlet": (C.c_int, [c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_i32]),
    "fsb_spmv": (C.c_int, [c_vp, c_vp, c_vp]),
    "fsb_dot": (C.c_int, [c_vp, c_vp, P(c_dbl)]),
    "fsb_solve_cg": (C.c_int, [c_vp, c_vp, c_vp, c_dbl, c_dbl, c_i32, c_i32, P(SolveInfo)]),
    "fsb_solve_bicgstab": (C.c_int, [c_vp, c_vp, c_vp, c_dbl, c_dbl, c_i32, c_i32, P(SolveInfo)]),
    "fsb_mg_create": (C.c_int, [c_vp, c_i32, c_vp, c_vp, c_i32, c_vp, P(c_vp)]),
    "fsb_mg_create_slab": (C.c_int, [c_vp, c_i32, c_vp, c_vp, c_i32, c_vp, c_i32, c_i32, c_i32, P(c_vp)]),
    "fsb_mg_omega": (C.c_int, [c_vp, c_i32, P(c_dbl)]),
    "fsb_mg_apply": (C.c_int, [c_vp, c_vp, c_vp, c_i32]),
    "fsb_solve_cg_mg": (C.c_int, [c_vp, c_vp, c_vp, c_dbl, c_dbl, c_i32, c_i32, P(SolveInfo)]),
    "fsb_mg_destroy": (None, [c_vp]),
    "fsb_dist_unique_id": (C.c_int, [c_vp]),
    "fsb_dist_init": (C.c_int, [c_vp, c_i32, c_i32, c_vp]),
    "fsb_dist_set_slab": (C.c_int, [c_vp, c_i32, c_i32, c_i64]),
    "fsb_dist_set_halo": (C.c_int, [c_vp, c_i64, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "fsb_dist_halo": (C.c_int, [c_vp]),
    "fsb_dist_allreduce_max": (C.c_int, [c_vp, P(c_dbl)]),
}

_lib = None


def load_library(path=None):
    """Load libfsb.so and attach the prototypes.  Raises SolverError when it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise SolverError("CUDA library %s is missing: run `python -m fenicssolver_b200.build` "
                          "(there is no CPU fallback)" % path)
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _np(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(c_vp)


class Context:
    """One per GPU (fsb_ctx).  `stream` is an optional cudaStream_t handle (int)."""

    def __init__(self, device=0, stream=None):
        self.lib = load_library()
        h = c_vp()
        rc = self.lib.fsb_init(int(device), c_vp(stream) if stream else None, C.byref(h))
        if rc != 0:
            raise SolverError("fsb_init(device=%d) failed with status %d (is a CUDA GPU visible?)" % (device, rc))
        self.h = h
        self.device = device
        self.rank, self.nranks = 0, 1

    def check(self, rc):
        if rc != 0:
            msg = self.lib.fsb_last_error(self.h)
            raise SolverError("libfsb error %d: %s" % (rc, msg.decode() if msg else "?"))

    def close(self):
        if getattr(self, "h", None):
            self.lib.fsb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self.check(self.lib.fsb_sync(self.h))

    def set_option(self, name, value):
        self.check(self.lib.fsb_set_option(self.h, name.encode(), int(value)))

    def launch_count(self):
        return int(self.lib.fsb_launch_count(self.h))

    def device_info(self):
        sm, fr, tot = c_i32(), c_i64(), c_i64()
        self.check(self.lib.fsb_device_info(self.h, C.byref(sm), C.byref(fr), C.byref(tot)))
        return {"sm_count": sm.value, "free_bytes": fr.value, "total_bytes": tot.value}

    # distributed -------------------------------------------------------------------------------
    def dist_unique_id(self):
        buf = (C.c_char * 128)()
        rc = self.lib.fsb_dist_unique_id(buf)
        if rc != 0:
            raise SolverError("fsb_dist_unique_id failed (%d): NCCL not loadable" % rc)
        return bytes(buf)

    def dist_init(self, rank, nranks, uid):
        buf = (C.c_char * 128).from_buffer_copy(uid)
        self.check(self.lib.fsb_dist_init(self.h, rank, nranks, buf))
        self.rank, self.nranks = rank, nranks

    def dist_set_slab(self, ghost_lo, ghost_hi, owned_planes):
        self.check(self.lib.fsb_dist_set_slab(self.h, ghost_lo, ghost_hi, owned_planes))

    def dist_set_halo(self, n_owned, n_local, neigh_rank, send_ptr, send_idx, recv_off, recv_cnt):
        nr = _np(neigh_rank, np.int32)
        sp_, si, ro, rc_ = _np(send_ptr, np.int64), _np(send_idx, np.int64), _np(recv_off, np.int64), _np(recv_cnt, np.int64)
        self.check(self.lib.fsb_dist_set_halo(self.h, int(n_owned), int(n_local), nr.size, _ptr(nr), _ptr(sp_), _ptr(si), _ptr(ro), _ptr(rc_)))

    def allreduce_max(self, value):
        v = c_dbl(value)
        self.check(self.lib.fsb_dist_allreduce_max(self.h, C.byref(v)))
        return v.value


class _Handle:
    _destroy = None

    def __init__(self, ctx, h):
        self.ctx, self.h = ctx, h

    def close(self):
        if getattr(self, "h", None) and self.ctx.h:
            getattr(self.ctx.lib, self._destroy)(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceMesh(_Handle):
    _destroy = "fsb_mesh_destroy"

    @classmethod
    def upload(cls, ctx, coords, cells, vertex_offset=0):
        """`vertex_offset`: `cells` holds global ids of a mesh whose vertices [offset, offset + len(coords)) are `coords` (a slab of
        a larger mesh, passed as slices of the global arrays); the ids are made local on the device."""
        coords = _np(coords, np.float64)
        cells = _np(cells, np.int32)
        h = c_vp()
        gdim, tdim = coords.shape[1], cells.shape[1] - 1
        ctx.check(ctx.lib.fsb_mesh_upload_part(ctx.h, gdim, tdim, coords.shape[0], _ptr(coords), cells.shape[0], _ptr(cells),
                                               int(vertex_offset), C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def upload_p2(cls, ctx, coords, cell_nodes, nnodes, nverts=None):
        """Degree-2 node layout: cell_nodes[nc][6|10] = sorted vertices then edge nodes (UFC order).  `nverts`: how many
        leading rows of `coords` to upload (default: all; pass nnodes with coordinates for every node when the vertex
        nodes do not come first, as in a partitioned space)."""
        coords = _np(coords, np.float64)
        cell_nodes = _np(cell_nodes, np.int32)
        gdim = coords.shape[1]
        h = c_vp()
        ctx.check(ctx.lib.fsb_mesh_upload_p2(ctx.h, gdim, gdim, int(nverts) if nverts is not None else coords.shape[0], _ptr(coords),
                                             cell_nodes.shape[0], _ptr(cell_nodes), int(nnodes), C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def box(cls, ctx, n, p0, p1, layer0=0, layer1=None):
        tdim = len(n)
        n_ = _np(n, np.int32)
        p0_, p1_ = _np(p0, np.float64), _np(p1, np.float64)
        if layer1 is None:
            layer1 = int(n[-1])
        h = c_vp()
        ctx.check(ctx.lib.fsb_mesh_box(ctx.h, tdim, _ptr(n_), _ptr(p0_), _ptr(p1_), layer0, layer1, C.byref(h)))
        return cls(ctx, h)

    def sizes(self):
        g, t, nv, nc = c_i32(), c_i32(), c_i64(), c_i64()
        self.ctx.check(self.ctx.lib.fsb_mesh_sizes(self.h, C.byref(g), C.byref(t), C.byref(nv), C.byref(nc)))
        return g.value, t.value, nv.value, nc.value

    def exterior_facets(self, ids=True):
        """K1 on the device: (fverts[nbf, tdim], opp[nbf], cell[nbf], facet_id[nbf]) of the facets held by exactly one cell, in
        lexicographic order of the vertex tuples; facet_id is dolfin's global facet index (`ids=False`: None — the ranking pass
        is only run when the ids are asked for)."""
        nbf, nf = c_i64(), c_i64()
        self.ctx.check(self.ctx.lib.fsb_mesh_exterior_facets(self.h, C.byref(nbf), C.byref(nf)))
        _, t, _, _ = self.sizes()
        fv = np.empty((nbf.value, t), dtype=np.int32)
        opp = np.empty(nbf.value, dtype=np.int32)
        cell = np.empty(nbf.value, dtype=np.int32)
        fid = np.empty(nbf.value, dtype=np.int64) if ids else None
        self.ctx.check(self.ctx.lib.fsb_mesh_exterior_facets_get(self.h, _ptr(fv), _ptr(opp), _ptr(cell), _ptr(fid)))
        self.num_facets = nf.value
        return fv, opp, cell, fid

    def exterior_facet_ids(self):
        nbf = c_i64()
        self.ctx.check(self.ctx.lib.fsb_mesh_exterior_facets(self.h, C.byref(nbf), None))
        fid = np.empty(nbf.value, dtype=np.int64)
        self.ctx.check(self.ctx.lib.fsb_mesh_exterior_facets_get(self.h, None, None, None, _ptr(fid)))
        return fid

    def boundary_geometry(self):
        """(bverts[nbv], finv[nbf, tdim], bxyz[nbv, gdim], mid[nbf, gdim]): distinct boundary vertices, the exterior facets in terms of
        them, their coordinates and the facet midpoints — what SubDomain.mark evaluates its predicate on."""
        nbv, nbf = c_i64(), c_i64()
        self.ctx.check(self.ctx.lib.fsb_mesh_exterior_facets(self.h, C.byref(nbf), None))
        self.ctx.check(self.ctx.lib.fsb_mesh_boundary_geometry(self.h, C.byref(nbv), None, None, None, None))
        g, t, _, _ = self.sizes()
        bv = np.empty(nbv.value, dtype=np.int32)
        finv = np.empty((nbf.value, t), dtype=np.int32)
        bxyz = np.empty((nbv.value, g), dtype=np.float64)
        mid = np.empty((nbf.value, g), dtype=np.float64)
        self.ctx.check(self.ctx.lib.fsb_mesh_boundary_geometry(self.h, None, _ptr(bv), _ptr(finv), _ptr(bxyz), _ptr(mid)))
        return bv, finv, bxyz, mid

    def download(self):
        g, t, nv, nc = self.sizes()
        xyz = np.empty((nv, g), dtype=np.float64)
        cells = np.empty((nc, t + 1), dtype=np.int32)
        self.ctx.check(self.ctx.lib.fsb_mesh_download(self.h, _ptr(xyz), _ptr(cells)))
        return xyz, cells


class _PinnedBlock:
    """One fsb_host_alloc block; released (back to the library's pool) with the last array that views it."""

    def __init__(self, ctx, nbytes):
        self.ctx, self.ptr = ctx, c_vp()
        ctx.check(ctx.lib.fsb_host_alloc(ctx.h, int(nbytes), C.byref(self.ptr)))

    def __del__(self):
        try:
            if self.ptr and self.ctx.h is not None:
                self.ctx.lib.fsb_host_free(self.ctx.h, self.ptr)
        except Exception:
            pass
        self.ptr = None


def pinned_empty(ctx, n, dtype=np.float64):
    """np.empty(n, dtype) in page-locked memory owned by the context's pool."""
    nbytes = int(n) * np.dtype(dtype).itemsize
    blk = _PinnedBlock(ctx, max(nbytes, 8))
    buf = (C.c_char * max(nbytes, 8)).from_address(blk.ptr.value)
    buf._fsb_block = blk                       # the ctypes buffer is the numpy array's base: keeps the block alive
    return np.frombuffer(buf, dtype=dtype, count=int(n))


class DeviceVector(_Handle):
    _destroy = "fsb_vec_destroy"

    def __init__(self, ctx, n=None, h=None):
        if h is None:
            h = c_vp()
            ctx.check(ctx.lib.fsb_vec_create(ctx.h, int(n), C.byref(h)))
        super().__init__(ctx, h)
        self.n = int(n)

    @classmethod
    def from_numpy(cls, ctx, a):
        a = _np(a, np.float64).ravel()
        v = cls(ctx, a.size)
        v.upload(a)
        return v

    def upload(self, a):
        a = _np(a, np.float64).ravel()
        self.ctx.check(self.ctx.lib.fsb_vec_upload(self.h, _ptr(a), a.size))

    def numpy(self):
        """Host copy.  Large vectors land in a page-locked block of the library's pool (one direct DMA, no page faults on a fresh
        array); the block returns to the pool when the array is garbage-collected."""
        out = pinned_empty(self.ctx, self.n) if self.n >= (1 << 20) else np.empty(self.n, dtype=np.float64)
        self.ctx.check(self.ctx.lib.fsb_vec_download(self.h, _ptr(out), self.n))
        return out

    def fill(self, value):
        self.ctx.check(self.ctx.lib.fsb_vec_fill(self.h, float(value)))

    def copy_from(self, other):
        self.ctx.check(self.ctx.lib.fsb_vec_copy(self.h, other.h))

    def axpy(self, a, x):
        self.ctx.check(self.ctx.lib.fsb_vec_axpy(self.h, float(a), x.h))

    def add_entries(self, idx, vals):
        i, v = _np(idx, np.int64).ravel(), _np(vals, np.float64).ravel()
        self.ctx.check(self.ctx.lib.fsb_vec_add_entries(self.h, i.size, _ptr(i), _ptr(v)))

    def dot(self, other):
        r = c_dbl()
        self.ctx.check(self.ctx.lib.fsb_dot(self.h, other.h, C.byref(r)))
        return r.value

    def halo(self):
        self.ctx.check(self.ctx.lib.fsb_dist_halo(self.h))

    def ptr(self):
        return self.ctx.lib.fsb_vec_ptr(self.h)


class DeviceMatrix(_Handle):
    _destroy = "fsb_mat_destroy"

    @classmethod
    def create(cls, mesh, ncomp=1):
        h = c_vp()
        mesh.ctx.check(mesh.ctx.lib.fsb_mat_create(mesh.h, ncomp, C.byref(h)))
        m = cls(mesh.ctx, h)
        m.mesh = mesh
        return m

    @classmethod
    def from_csr(cls, ctx, row_ptr, col_idx, vals):
        rp, ci, va = _np(row_ptr, np.int64), _np(col_idx, np.int32), _np(vals, np.float64)
        h = c_vp()
        ctx.check(ctx.lib.fsb_mat_from_csr(ctx.h, rp.size - 1, _ptr(rp), _ptr(ci), _ptr(va), C.byref(h)))
        m = cls(ctx, h)
        m.mesh = None
        return m

    def sizes(self):
        nrows, nnz, bs, nnzb = c_i64(), c_i64(), c_i32(), c_i64()
        self.ctx.check(self.ctx.lib.fsb_mat_sizes(self.h, C.byref(nrows), C.byref(nnz), C.byref(bs), C.byref(nnzb)))
        return {"nrows": nrows.value, "nnz": nnz.value, "bs": bs.value, "nnzb": nnzb.value}

    def download_csr(self, values=True):
        s = self.sizes()
        rp = np.empty(s["nrows"] + 1, dtype=np.int64)
        ci = np.empty(s["nnz"], dtype=np.int32)
        va = np.empty(s["nnz"], dtype=np.float64) if values else None
        self.ctx.check(self.ctx.lib.fsb_mat_download_csr(self.h, _ptr(rp), _ptr(ci), _ptr(va)))
        return rp, ci, va

    def zero(self):
        self.ctx.check(self.ctx.lib.fsb_mat_zero(self.h))

    def set_owned_rows(self, r0, r1):
        self.ctx.check(self.ctx.lib.fsb_mat_set_owned_rows(self.h, int(r0), int(r1)))

    # assembly ------------------------------------------------------------------------------------
    def assemble_scalar(self, kscale=1.0, ktensor=None, mass=0.0, adv=0.0, vel=None, overwrite=False):
        """A += the scalar form; `overwrite`: A = the form (zero() + assemble in one call; the row-gather kernel skips the zero-fill)."""
        kt = None if ktensor is None else _np(ktensor, np.float64)
        ve = None if vel is None else _np(vel, np.float64)
        fn = self.ctx.lib.fsb_assemble_scalar_set if overwrite else self.ctx.lib.fsb_assemble_scalar
        self.ctx.check(fn(self.mesh.h, self.h, float(kscale), _ptr(kt), float(mass), float(adv), _ptr(ve)))

    def assemble_elasticity(self, mu, lmbda):
        self.ctx.check(self.ctx.lib.fsb_assemble_elasticity(self.mesh.h, self.h, float(mu), float(lmbda)))

    def assemble_facet_mass(self, fverts, h):
        fv = _np(fverts, np.int32)
        self.ctx.check(self.ctx.lib.fsb_assemble_facet_mass(self.mesh.h, self.h, fv.shape[0], _ptr(fv), float(h)))

    def apply_dirichlet(self, b, dofs, vals, symmetric=True, x=None):
        d = _np(dofs, np.int64)
        v = _np(np.broadcast_to(np.asarray(vals, dtype=np.float64), d.shape), np.float64)
        self.ctx.check(self.ctx.lib.fsb_apply_dirichlet(self.h, b.h, x.h if x is not None else None, d.size, _ptr(d), _ptr(v), int(bool(symmetric))))

    # Krylov --------------------------------------------------------------------------------------
    def spmv(self, x, y):
        self.ctx.check(self.ctx.lib.fsb_spmv(self.h, x.h, y.h))

    def solve(self, b, x, method="cg", rtol=1e-12, atol=0.0, maxit=10000, precond="jacobi"):
        info = SolveInfo()
        fn = {"cg": self.ctx.lib.fsb_solve_cg, "bicgstab": self.ctx.lib.fsb_solve_bicgstab}[method]
        pc = {"none": 0, None: 0, "jacobi": 1}[precond]
        self.ctx.check(fn(self.h, b.h, x.h, float(rtol), float(atol), int(maxit), pc, C.byref(info)))
        return {"iterations": info.iterations, "converged": info.converged, "rnorm": info.rnorm, "bnorm": info.bnorm,
                "solve_ms": info.solve_ms, "spmv_ms": info.spmv_ms, "operand_nnzb": info.operand_nnzb}


class Multigrid(_Handle):
    """Geometric multigrid hierarchy over assembled level matrices (fine first) on nested box meshes."""
    _destroy = "fsb_mg_destroy"

    def __init__(self, ctx, matrices, ncells, tdim, omega=None, slab=None):
        """`omega`: per-level dampings from an earlier hierarchy on the same levels (skips the eigenvalue estimates).
        `slab`: (layer0, owned_z0, owned_z1) of a slab-distributed fine level — matrices[0] is this rank's slab, the coarser
        matrices are whole levels replicated on every rank."""
        self.matrices = list(matrices)              # keep the level matrices alive
        arr = (c_vp * len(matrices))(*[m.h for m in matrices])
        nc = np.zeros((len(matrices), 3), dtype=np.int32)
        for l, n in enumerate(ncells):
            nc[l, :len(n)] = n
        h = c_vp()
        om = None if omega is None else _np(omega, np.float64)
        if slab is None:
            ctx.check(ctx.lib.fsb_mg_create(ctx.h, len(matrices), arr, _ptr(nc), int(tdim), _ptr(om), C.byref(h)))
        else:
            ctx.check(ctx.lib.fsb_mg_create_slab(ctx.h, len(matrices), arr, _ptr(nc), int(tdim), _ptr(om), int(slab[0]), int(slab[1]), int(slab[2]), C.byref(h)))
        super().__init__(ctx, h)

    def apply(self, r, z, nu=2):
        self.ctx.check(self.ctx.lib.fsb_mg_apply(self.h, r.h, z.h, int(nu)))

    def omega(self, level):
        w = c_dbl()
        self.ctx.check(self.ctx.lib.fsb_mg_omega(self.h, int(level), C.byref(w)))
        return w.value

    def solve(self, b, x, rtol=1e-12, atol=0.0, maxit=1000, nu=2):
        info = SolveInfo()
        self.ctx.check(self.ctx.lib.fsb_solve_cg_mg(self.h, b.h, x.h, float(rtol), float(atol), int(maxit), int(nu), C.byref(info)))
        return {"iterations": info.iterations, "converged": info.converged, "rnorm": info.rnorm, "bnorm": info.bnorm,
                "solve_ms": info.solve_ms, "spmv_ms": info.spmv_ms, "operand_nnzb": info.operand_nnzb}


def apply_scalar(mesh, x, y, kscale=1.0, ktensor=None, mass=0.0, adv=0.0, vel=None):
    kt = None if ktensor is None else _np(ktensor, np.float64)
    ve = None if vel is None else _np(vel, np.float64)
    mesh.ctx.check(mesh.ctx.lib.fsb_apply_scalar(mesh.h, x.h, y.h, float(kscale), _ptr(kt), float(mass), float(adv), _ptr(ve)))


def assemble_source(mesh, b, S, ncomp=1, scale=1.0, cell_tags=None, tag=0):
    s = _np(np.broadcast_to(np.asarray(S, dtype=np.float64), (ncomp,)), np.float64)
    tags = None if cell_tags is None else _np(cell_tags, np.int32)
    mesh.ctx.check(mesh.ctx.lib.fsb_assemble_source(mesh.h, b.h, ncomp, _ptr(s), float(scale), _ptr(tags), int(tag)))


def assemble_source_nodal(mesh, b, S, ncomp=1, scale=1.0):
    mesh.ctx.check(mesh.ctx.lib.fsb_assemble_source_nodal(mesh.h, b.h, ncomp, S.h, float(scale)))


def assemble_facet_load(mesh, b, fverts, g, ncomp=1, scale=1.0, opp=None, normal=False):
    fv = _np(fverts, np.int32)
    gv = _np(np.atleast_1d(np.asarray(g, dtype=np.float64)), np.float64)
    op = None if opp is None else _np(opp, np.int32)
    mesh.ctx.check(mesh.ctx.lib.fsb_assemble_facet_load(mesh.h, b.h, ncomp, fv.shape[0], _ptr(fv), _ptr(op), 1 if normal else 0, _ptr(gv), float(scale)))


def facet_area(mesh, fverts):
    fv = _np(fverts, np.int32)
    a = c_dbl()
    mesh.ctx.check(mesh.ctx.lib.fsb_facet_area(mesh.h, fv.shape[0], _ptr(fv), C.byref(a)))
    return a.value


def assemble_thermal_load(mesh, b, beta, T=None, T_const=0.0, T_ref=0.0, scale=1.0):
    """b += scale * beta * int (T - T_ref) div(v) dx; T: DeviceVector of nodal temperatures or None (constant T_const)."""
    mesh.ctx.check(mesh.ctx.lib.fsb_assemble_thermal_load(mesh.h, b.h, float(beta), T.h if T is not None else None,
                                                          float(T_const), float(T_ref), float(scale)))


def assemble_von_mises_load(mesh, u, mu, lmbda, b):
    mesh.ctx.check(mesh.ctx.lib.fsb_assemble_von_mises_load(mesh.h, u.h, float(mu), float(lmbda), b.h))


def assemble_facet_radiation(mesh, A, r, T, fverts, m, T_ambient, rscale=1.0):
    """A += int 4 m T^3 u v ds, r += rscale * int m (T^4 - Ta^4) v ds over the facets (A or r may be None)."""
    fv = _np(fverts, np.int32)
    mesh.ctx.check(mesh.ctx.lib.fsb_assemble_facet_radiation(mesh.h, A.h if A is not None else None, r.h if r is not None else None,
                                                             T.h, fv.shape[0], _ptr(fv), float(m), float(T_ambient), float(rscale)))


def assemble_scalar_nonlinear_k(mesh, A, r, T, k, dk, scale=1.0, rscale=1.0):
    """Newton terms of int k(T) grad T . grad v: A += Jacobian, r += rscale * residual (A or r may be None)."""
    mesh.ctx.check(mesh.ctx.lib.fsb_assemble_scalar_nonlinear_k(mesh.h, A.h if A is not None else None, r.h if r is not None else None,
                                                                T.h, k.h, dk.h, float(scale), float(rscale)))


# SUPG extra terms (test function q + tau vel.grad q): see include/fsb.h
def assemble_scalar_supg(mesh, A, vel, pe, mass=0.0, adv=0.0, x=None, y=None):
    ve = _np(vel, np.float64)
    mesh.ctx.check(mesh.ctx.lib.fsb_assemble_scalar_supg(mesh.h, A.h if A is not None else None, x.h if x is not None else None,
                                                         y.h if y is not None else None, float(mass), float(adv), _ptr(ve), float(pe)))


def assemble_source_supg(mesh, b, S, vel, pe, cell_tags=None, tag=0):
    ve = _np(vel, np.float64)
    tags = None if cell_tags is None else _np(cell_tags, np.int32)
    mesh.ctx.check(mesh.ctx.lib.fsb_assemble_source_supg(mesh.h, b.h, float(S), _ptr(ve), float(pe), _ptr(tags), int(tag)))


def assemble_facet_supg(mesh, A, b, fverts, opp, vel, pe, g=0.0, h=0.0):
    fv, op, ve = _np(fverts, np.int32), _np(opp, np.int32), _np(vel, np.float64)
    mesh.ctx.check(mesh.ctx.lib.fsb_assemble_facet_supg(mesh.h, A.h if A is not None else None, b.h if b is not None else None,
                                                        fv.shape[0], _ptr(fv), _ptr(op), float(g), float(h), _ptr(ve), float(pe)))


def assemble_advection_nodal(mesh, A, vel, scale=1.0, x=None, y=None):
    """A += scale * int (v_h.grad u) v dx with v_h the P1 interpolant of the nodal velocity DeviceVector `vel` (or y += .. x)."""
    mesh.ctx.check(mesh.ctx.lib.fsb_assemble_advection_nodal(mesh.h, A.h if A is not None else None, x.h if x is not None else None,
                                                             y.h if y is not None else None, vel.h, float(scale)))
