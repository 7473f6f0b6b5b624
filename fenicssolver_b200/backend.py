"""Device side of a P1 space: mesh upload/generation, pattern, assembly calls, Dirichlet, Krylov.

This is the layer SolverBase.solve_linear_problem / solve_amg delegate to instead of dolfin's
assemble / assemble_system / LinearVariationalSolver / PETScKrylovSolver
(/root/reference/FenicsSolver/SolverBase.py:592-672).  It only sequences libfsb calls; there is no
numerical fallback on the host.

Distributed runs (one process per GPU): a box mesh is cut into z-slabs of vertex planes, each rank
generates its slab plus one ghost plane per neighbour on its own device, assembles without
communication (owner computes) and solves with halo exchange + all-reduced dot products.  Any other
mesh (file meshes, user arrays) and every degree-2 space is split by recursive coordinate bisection of its nodes
(partition.NodePartition): owned nodes first, ghosts appended, general send lists for the halo.
"""
from __future__ import annotations

import os
import time
import weakref

import numpy as np

from . import _lib
from ._lib import SolverError
from .partition import NodePartition, rcb_partition

_contexts = {}


def get_context(device=None, stream=None):
    """One libfsb context per device and process (LOCAL_RANK selects the device by default).
    `stream`: a cudaStream_t handle (e.g. torch.cuda.Stream().cuda_stream) for the first creation."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    ctx = _contexts.get(device)
    if ctx is None or ctx.h is None:
        ctx = _lib.Context(device, stream=stream)
        _contexts[device] = ctx
    return ctx


class Comm:
    """Process group description for the z-slab decomposition.  `bootstrap` broadcasts the NCCL
    unique id (torch.distributed is used only for that and for host-side gathers)."""

    def __init__(self, rank=0, nranks=1, bootstrap=None):
        self.rank, self.nranks, self.bootstrap = rank, nranks, bootstrap

    @classmethod
    def from_torch(cls):
        import torch.distributed as dist
        if not dist.is_available() or not dist.is_initialized():
            return cls()

        def bcast(obj):
            box = [obj]
            dist.broadcast_object_list(box, src=0)
            return box[0]
        return cls(dist.get_rank(), dist.get_world_size(), bcast)


def slab_partition(nplanes, nranks):
    """Contiguous blocks of vertex planes, as even as possible: [(p0, p1), ...] per rank."""
    base, rem = divmod(nplanes, nranks)
    out, p = [], 0
    for r in range(nranks):
        k = base + (1 if r < rem else 0)
        out.append((p, p + k))
        p += k
    return out


class DeviceSpace:
    """P1 space with `ncomp` components on a mesh, resident on one GPU (or one slab of it)."""

    def __init__(self, mesh, ncomp=1, ctx=None, comm=None, space=None):
        self.mesh, self.ncomp = mesh, ncomp
        self.fs = space                                 # FunctionSpace (None: P1 on the mesh)
        self.degree = getattr(space, "degree", 1)
        self.ctx = ctx or get_context()
        self.comm = comm or Comm()
        self.timings = {}
        t0 = time.perf_counter()
        self.v_off = 0                                  # global vertex id of local vertex 0
        self.ghost_lo = self.ghost_hi = 0
        nv_global = mesh.num_vertices() if self.degree == 1 else space.num_nodes()
        self.part = None                                # NodePartition of a general (non-slab) distributed space
        general = self.comm.nranks > 1 and (self.degree == 2 or not mesh.box or getattr(mesh, "force_general_partition", False))
        if general:
            # host integer work on the replicated mesh: node partition, local numbering, halo lists
            cell_nodes = mesh.cells() if self.degree == 1 else space.cell_nodes()
            node_xyz = mesh.coordinates() if self.degree == 1 else space.node_coordinates()
            owner = rcb_partition(node_xyz, self.comm.nranks)
            self.part = NodePartition(cell_nodes, owner, self.comm.rank, self.comm.nranks)
            pt = self.part
            if self.degree == 1:
                self.dmesh = _lib.DeviceMesh.upload(self.ctx, node_xyz[pt.l2g], np.sort(pt.cell_nodes_local, axis=1))
            else:
                # coordinates for every local node, so the vertex nodes of a cell may carry any local id
                self.dmesh = _lib.DeviceMesh.upload_p2(self.ctx, node_xyz[pt.l2g], pt.cell_nodes_local, pt.n_local, nverts=pt.n_local)
            if self.ctx.nranks == 1:
                uid = self.ctx.dist_unique_id() if self.comm.rank == 0 else None
                uid = self.comm.bootstrap(uid)
                self.ctx.dist_init(self.comm.rank, self.comm.nranks, uid)
            self.activate(force=True)
        elif self.degree == 2:
            # degree-2 node layout: host integer work (edge numbering), then one upload
            self.dmesh = _lib.DeviceMesh.upload_p2(self.ctx, mesh.coordinates(), space.cell_nodes(), space.num_nodes())
        elif self.comm.nranks > 1:
            if getattr(mesh, "distributed", None) != (self.comm.rank, self.comm.nranks):
                mesh.distributed = (self.comm.rank, self.comm.nranks)
                mesh._exterior = None
                mesh.__dict__.pop("_boundary_geometry", None)
            self.dmesh, lay = mesh.slab_device_mesh(self.ctx)          # made once; the boundary search (K1) ran on the same slab
            self.plane = lay["plane"]
            self.ghost_lo, self.ghost_hi = lay["ghost_lo"], lay["ghost_hi"]
            self.owned_planes = lay["owned_planes"]
            self.v_off = lay["v_off"]
            if self.ctx.nranks == 1:
                uid = self.ctx.dist_unique_id() if self.comm.rank == 0 else None
                uid = self.comm.bootstrap(uid)
                self.ctx.dist_init(self.comm.rank, self.comm.nranks, uid)
            self.activate(force=True)
        else:
            self.dmesh = mesh.device_mesh(self.ctx)         # generated (box) or uploaded once; shared with the boundary search (K1)
        _, _, self.nv_local, self.nc_local = self.dmesh.sizes()
        if self.degree == 2:
            self.nv_local = space.num_nodes()           # rows are P2 nodes (vertices + edges)
        if self.part is not None:
            self.nv_local = self.part.n_local
            self.own_v0, self.own_v1 = 0, self.part.n_owned
        else:
            self.own_v0 = self.ghost_lo * getattr(self, "plane", 0)
            self.own_v1 = self.own_v0 + (self.owned_planes * self.plane if self.comm.nranks > 1 else self.nv_local)
        self.nv_global = nv_global
        self.ctx.sync()
        self.timings["mesh"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        self.A = _lib.DeviceMatrix.create(self.dmesh, ncomp)          # symbolic phase (K2)
        if self.comm.nranks > 1:
            self.A.set_owned_rows(self.own_v0, self.own_v1)
        self.ctx.sync()
        self.timings["symbolic"] = time.perf_counter() - t0

    def activate(self, force=False):
        """The halo layout (slab or send lists) is state of the context: declare this space's layout if another
        distributed space on the same context was the last one to do so."""
        owner = getattr(self.ctx, "_layout_owner", None)
        if self.comm.nranks == 1 or (not force and owner is not None and owner() is self):
            return
        if self.part is not None:
            pt = self.part
            self.ctx.dist_set_halo(pt.n_owned, pt.n_local, pt.neighbours, pt.send_ptr, pt.send_idx, pt.recv_off, pt.recv_cnt)
        else:
            self.ctx.dist_set_slab(self.ghost_lo, self.ghost_hi, self.owned_planes)
        self.ctx._layout_owner = weakref.ref(self)       # weak: the context must not keep a released space alive

    # ---- index helpers ----------------------------------------------------------------------------
    @property
    def ndof_local(self):
        return self.nv_local * self.ncomp

    @property
    def ndof_global(self):
        return self.nv_global * self.ncomp

    def local_vertices(self, gverts):
        """Global vertex ids -> local ids, dropping those outside this rank's planes."""
        g = np.asarray(gverts, dtype=np.int64)
        if self.part is not None:
            l, ok = self.part.to_local(g)
            return l[ok]
        l = g - self.v_off
        return l[(l >= 0) & (l < self.nv_local)]

    def local_facets(self, fverts, opp=None):
        """Facets (global vertex lists) -> what the facet kernels take: local vertex lists (P1) or the facets'
        P2 node lists (vertices then edges), plus the opposite vertices when given."""
        if self.part is not None:
            fn = self.fs.facet_nodes(fverts) if self.degree == 2 else np.asarray(fverts, dtype=np.int64)
            nfn = fn.shape[1] if fn.ndim == 2 else 0
            if fn.shape[0] == 0:
                return np.zeros((0, nfn), dtype=np.int32), (None if opp is None else np.zeros(0, dtype=np.int32))
            l, ok = self.part.to_local(fn)
            keep = ok.all(axis=1)
            if opp is None:
                return (l[keep] if self.degree == 2 else np.sort(l[keep], axis=1)).astype(np.int32), None
            lo, oko = self.part.to_local(opp)
            keep &= oko
            return (l[keep] if self.degree == 2 else np.sort(l[keep], axis=1)).astype(np.int32), lo[keep].astype(np.int32)
        if self.degree == 2:
            fn = self.fs.facet_nodes(fverts)
            return fn, (None if opp is None else np.asarray(opp, dtype=np.int32))
        fv = np.asarray(fverts, dtype=np.int64) - self.v_off
        keep = np.all((fv >= 0) & (fv < self.nv_local), axis=1) if fv.size else np.zeros(0, bool)
        if opp is None:
            return fv[keep].astype(np.int32), None
        op = np.asarray(opp, dtype=np.int64) - self.v_off
        keep &= (op >= 0) & (op < self.nv_local)
        return fv[keep].astype(np.int32), op[keep].astype(np.int32)

    def local_dofs(self, gdofs, gvals):
        """Global dof ids/values -> local, dropping those outside this rank's planes."""
        g = np.asarray(gdofs, dtype=np.int64)
        v = np.broadcast_to(np.asarray(gvals, dtype=np.float64), g.shape)
        if self.part is not None:
            ln, ok = self.part.to_local(g // self.ncomp)
            return (ln * self.ncomp + g % self.ncomp)[ok], v[ok]
        l = g - self.v_off * self.ncomp
        keep = (l >= 0) & (l < self.ndof_local)
        return l[keep], v[keep]

    def local_cell_tags(self, tags):
        """Global per-cell array -> the cells this rank holds."""
        if self.part is not None:
            return np.asarray(tags)[self.part.cells_global]
        if self.comm.nranks > 1:
            per_layer = self.mesh.num_cells() // self.mesh.box["n"][-1]
            return np.asarray(tags)[self.v_off // self.plane * per_layer:][:self.nc_local]
        return tags

    def local_coordinates(self):
        if self.part is not None:
            xyz = self.mesh.coordinates() if self.degree == 1 else self.fs.node_coordinates()
            return xyz[self.part.l2g]
        if self.comm.nranks == 1:
            return self.mesh.coordinates()
        return self.mesh.coordinates()[self.v_off:self.v_off + self.nv_local]

    # ---- vectors ----------------------------------------------------------------------------------
    def vector(self, fill=None):
        v = _lib.DeviceVector(self.ctx, self.ndof_local)
        if fill is not None and fill != 0.0:
            v.fill(fill)
        return v

    def vector_from_global(self, values):
        a = np.asarray(values, dtype=np.float64).ravel()
        if self.part is not None:
            return _lib.DeviceVector.from_numpy(self.ctx, a.reshape(-1, self.ncomp)[self.part.l2g].ravel())
        lo = self.v_off * self.ncomp
        return _lib.DeviceVector.from_numpy(self.ctx, a[lo:lo + self.ndof_local])

    def local_nodal(self, values, ncomp=1):
        """Device vector with this rank's part (owned + ghost nodes) of a GLOBAL nodal field with `ncomp` values per node — a
        field on another space over the same nodes, e.g. a velocity or temperature field entering a form."""
        a = np.asarray(values, dtype=np.float64).reshape(-1, ncomp)
        if a.shape[0] != self.nv_global:
            raise SolverError("a nodal field needs one value per node: got %d rows for %d nodes" % (a.shape[0], self.nv_global))
        if self.part is not None:
            a = a[self.part.l2g]
        elif self.comm.nranks > 1:
            a = a[self.v_off:self.v_off + self.nv_local]
        return _lib.DeviceVector.from_numpy(self.ctx, a.ravel())

    def scratch_vector(self, key):
        """A zeroed device vector owned by the space and reused between calls (the right-hand side of every
        time step): cudaMalloc/cudaFree synchronise the device and are kept out of the step."""
        cache = self.__dict__.setdefault("_scratch", {})
        v = cache.get(key)
        if v is None:
            v = cache[key] = _lib.DeviceVector(self.ctx, self.ndof_local)
        else:
            v.fill(0.0)
        return v

    def vector_from_function(self, fn):
        """Device copy of a Function: a known-uniform field is filled on the device (no H2D copy)."""
        u = fn.uniform_value()
        if u is not None:
            return self.vector(fill=u)
        dev = fn.device_vector()
        if dev is not None and getattr(fn, "_local", False) and dev.n == self.ndof_local:
            self.activate()
            dev.halo()                                  # a solve leaves the ghosts of its solution vector stale
            return dev
        return self.vector_from_global(fn.array())

    def owned_values(self, dvec):
        a = dvec.numpy()
        return a[self.own_v0 * self.ncomp:self.own_v1 * self.ncomp]

    def gather_global(self, dvec):
        """Solution in global vertex order on every rank (host).  Single GPU: one D2H copy."""
        mine = self.owned_values(dvec)
        if self.comm.nranks == 1:
            return mine
        import torch.distributed as dist
        parts = [None] * self.comm.nranks
        if self.part is not None:
            dist.all_gather_object(parts, (self.part.owned, mine))
            out = np.empty((self.nv_global, self.ncomp))
            for ids, vals in parts:
                out[ids] = vals.reshape(-1, self.ncomp)
            return out.ravel()
        dist.all_gather_object(parts, mine)
        return np.concatenate(parts)

    # ---- Dirichlet + solve ------------------------------------------------------------------------
    def apply_dirichlet(self, b, gdofs, gvals, symmetric, x=None):
        self.activate()
        d, v = self.local_dofs(gdofs, gvals)
        self.A.apply_dirichlet(b, d, v, symmetric=symmetric, x=x)

    def solve(self, b, x, method="cg", rtol=1e-12, atol=0.0, maxit=100000, precond="jacobi"):
        self.activate()
        info = self.A.solve(b, x, method=method, rtol=rtol, atol=atol, maxit=maxit, precond=precond)
        return info
