"""Plain-data stand-ins for the dolfin objects FenicsSolver user code passes in.

Reference scripts do `from dolfin import *` and hand Mesh / FunctionSpace / Constant / SubDomain
objects to the solvers (e.g. /root/reference/examples/test_heat_transfer.py:34-45,59-89).  The same
scripts run against this package with `from fenicssolver_b200.dolfin_compat import *`: the classes
below carry only data (numpy arrays, numbers, predicates); all arithmetic happens in libfsb.

Host-side integer work that dolfin does at mesh construction lives here too (K1 in SURVEY 2.2):
facet numbering (global facet id = lexicographic rank of the sorted vertex tuple), exterior-facet
detection, SubDomain.mark.
"""
from __future__ import annotations

import math
import re

import numpy as np

from ._lib import SolverError

DOLFIN_EPS = 3.0e-16
__all__ = ["DOLFIN_EPS", "near", "Point", "Constant", "Expression", "Mesh", "UnitSquareMesh", "RectangleMesh",
           "UnitCubeMesh", "BoxMesh", "SubDomain", "AutoSubDomain", "MeshFunction", "FacetMarkers", "FunctionSpace",
           "VectorFunctionSpace", "Function", "DirichletBC", "PointSource", "SolverError", "plot", "interactive", "set_log_level",
           "File", "boundary_flux", "CRITICAL", "ERROR", "WARNING", "INFO", "PROGRESS", "DEBUG", "parameters"]

# log levels and no-op stand-ins for the dolfin calls the reference's example scripts make around the solver
# (set_log_level(ERROR), plot(...), interactive(): examples/test_linear_elasticity.py:31,141; test_heat_transfer.py:172-192)
CRITICAL, ERROR, WARNING, INFO, PROGRESS, DEBUG = 50, 40, 30, 20, 16, 10
parameters = {"linear_algebra_backend": "libfsb", "mesh_partitioner": "rcb", "form_compiler": {"optimize": True}}


def set_log_level(level):
    import logging
    logging.getLogger("fenicssolver_b200").setLevel(int(level))


def plot(*args, **kwargs):
    """Batch-safe no-op (dolfin.plot); SolverBase.plot() draws 2-D scalar results when plotting is interactive."""
    return None


def interactive():
    return None


def near(a, b, eps=DOLFIN_EPS):
    """dolfin.near: |a-b| <= eps (absolute).  Vectorised over numpy arrays."""
    return np.abs(np.asarray(a) - b) <= eps


class Point:
    def __init__(self, *xyz):
        self.x = np.array(xyz, dtype=np.float64)

    def __getitem__(self, i):
        return self.x[i]

    def __len__(self):
        return self.x.size


class Constant:
    """dolfin.Constant: a number or a tuple of numbers."""

    def __init__(self, value):
        self._v = np.array(value, dtype=np.float64)

    def values(self):
        return np.atleast_1d(self._v).copy()

    @property
    def ufl_shape(self):
        return self._v.shape

    def __float__(self):
        return float(self._v)

    def __repr__(self):
        return "Constant(%s)" % (self._v.tolist(),)


_EXPR_NAMES = {k: getattr(np, k) for k in ("sin", "cos", "tan", "exp", "log", "sqrt", "fabs", "floor", "ceil", "arctan2")}
_EXPR_NAMES.update({"pow": np.power, "abs": np.abs, "pi": math.pi, "atan2": np.arctan2, "DOLFIN_PI": math.pi, "where": np.where, "logical_and": np.logical_and, "logical_or": np.logical_or, "logical_not": np.logical_not,
                    "fmin": np.minimum, "fmax": np.maximum, "min": np.minimum, "max": np.maximum, "sinh": np.sinh, "cosh": np.cosh,
                    "tanh": np.tanh, "asin": np.arcsin, "acos": np.arccos, "atan": np.arctan, "erf": None})
_EXPR_NAMES.pop("erf")


def _cpp_to_python(src):
    """C++ expression syntax that Python lacks: `cond ? a : b` (right-associative, any nesting) becomes where(cond, a, b);
    `a && b`, `a || b` become logical_and / logical_or calls (C++ precedence: weaker than the comparisons), `!x` becomes ~x."""
    def split_ternary(e):
        depth = 0
        q = -1
        for i, ch in enumerate(e):
            if ch in "([":
                depth += 1
            elif ch in ")]":
                depth -= 1
            elif ch == "?" and depth == 0:
                q = i
                break
        if q < 0:
            return None
        depth, nest = 0, 0
        for j in range(q + 1, len(e)):
            ch = e[j]
            if ch in "([":
                depth += 1
            elif ch in ")]":
                depth -= 1
            elif depth == 0 and ch == "?":
                nest += 1
            elif depth == 0 and ch == ":":
                if nest == 0:
                    return e[:q], e[q + 1:j], e[j + 1:]
                nest -= 1
        raise SolverError("unbalanced ?: in Expression %r" % e)

    def conv(e):
        # innermost parenthesised groups first, so a ternary inside f(...) is handled on its own
        out, i = "", 0
        while i < len(e):
            if e[i] == "(":
                depth, j = 1, i + 1
                while j < len(e) and depth:
                    depth += e[j] == "("
                    depth -= e[j] == ")"
                    j += 1
                out += "(" + ", ".join(conv(a) for a in _split_args(e[i + 1:j - 1])) + ")"
                i = j
            else:
                out += e[i]
                i += 1
        parts = split_ternary(out)
        if parts is not None:
            c, a, b = parts
            return "where(%s, %s, %s)" % (conv(c).strip(), conv(a).strip(), conv(b).strip())
        # || binds weaker than &&, both weaker than the comparisons: functions instead of Python's tightly binding | and &
        for op, fn in (("||", "logical_or"), ("&&", "logical_and")):
            pieces = _split_top(out, op)
            if len(pieces) > 1:
                acc = conv(pieces[0]).strip()
                for nxt in pieces[1:]:
                    acc = "%s(%s, %s)" % (fn, acc, conv(nxt).strip())
                return acc
        return out

    return conv(re.sub(r"!(?!=)", " _not_ ", src))


def _validate_expression_ast(py_src, names):
    """Only arithmetic on whitelisted names: no attribute access, no dunder names, no lambdas/comprehensions/strings.  Returns the
    compiled code.  `_not_ x` (from C++ `!x`) was rewritten to a logical_not call before parsing; integer-literal division follows
    C++ (both operands integer literals -> truncating division), everything else is floating point as in a compiled Expression."""
    import ast
    py_src = re.sub(r"_not_\s*(\w+(\[[^\]]*\])?|\([^()]*\))", r"logical_not(\1)", py_src)
    if "_not_" in py_src:
        raise SolverError("unsupported use of ! in Expression")
    try:
        tree = ast.parse(py_src.strip(), mode="eval")
    except SyntaxError as ex:
        raise SolverError("cannot parse Expression %r: %s" % (py_src, ex))
    allowed = (ast.Expression, ast.BinOp, ast.UnaryOp, ast.Compare, ast.Call, ast.Name, ast.Load, ast.Constant, ast.Subscript, ast.Tuple,
               ast.Add, ast.Sub, ast.Mult, ast.Div, ast.Pow, ast.Mod, ast.USub, ast.UAdd, ast.Invert, ast.Lt, ast.LtE, ast.Gt, ast.GtE,
               ast.Eq, ast.NotEq, ast.BitAnd, ast.BitOr, ast.FloorDiv)

    class IntDiv(ast.NodeTransformer):
        def visit_UnaryOp(self, node):
            self.generic_visit(node)
            if isinstance(node.op, ast.USub) and isinstance(node.operand, ast.Constant) and type(node.operand.value) is int:
                return ast.copy_location(ast.Constant(-node.operand.value), node)
            return node

        def visit_BinOp(self, node):
            self.generic_visit(node)
            if (isinstance(node.op, ast.Div) and isinstance(node.left, ast.Constant) and isinstance(node.right, ast.Constant)
                    and type(node.left.value) is int and type(node.right.value) is int and node.right.value != 0):
                q = abs(node.left.value) // abs(node.right.value)          # C++ truncates toward zero
                return ast.copy_location(ast.Constant(q if (node.left.value < 0) == (node.right.value < 0) else -q), node)
            return node

    for node in ast.walk(tree):
        if not isinstance(node, allowed):
            raise SolverError("unsupported construct %s in Expression %r" % (type(node).__name__, py_src))
        if isinstance(node, ast.Name) and (node.id.startswith("_") or node.id not in names):
            raise SolverError("unknown name %r in Expression %r" % (node.id, py_src))
        if isinstance(node, ast.Call) and not isinstance(node.func, ast.Name):
            raise SolverError("only calls of the math functions are allowed in Expression %r" % py_src)
        if isinstance(node, ast.Constant) and not isinstance(node.value, (int, float)):
            raise SolverError("only numeric literals are allowed in Expression %r" % py_src)
    tree = ast.fix_missing_locations(IntDiv().visit(tree))
    return compile(tree, "<Expression>", "eval")


def _split_top(e, op):
    """Split at top-level occurrences of the two-character operator `op`."""
    out, depth, cur, i = [], 0, "", 0
    while i < len(e):
        ch = e[i]
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if depth == 0 and e.startswith(op, i):
            out.append(cur)
            cur = ""
            i += len(op)
            continue
        cur += ch
        i += 1
    out.append(cur)
    return out


def _split_args(e):
    """Split at top-level commas."""
    args, depth, cur = [], 0, ""
    for ch in e:
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == "," and depth == 0:
            args.append(cur)
            cur = ""
        else:
            cur += ch
    args.append(cur)
    return args


class Expression:
    """dolfin.Expression for C++-style strings in x[0..2] (and keyword parameters), evaluated at the
    vertices, i.e. interpolated into P1 as SolverBase.py:310-314,364,387 do."""

    def __init__(self, code, degree=1, **params):
        self.code = code
        self.degree = degree
        self.params = params

    def _eval_one(self, src, x):
        env = dict(_EXPR_NAMES)
        env.update(self.params)
        env["x"] = x
        if not re.fullmatch(r"[\w\s\.\+\-\*/\(\)\[\],<>=!&|?:]*", src):
            raise SolverError("unsupported characters in Expression %r" % src)
        code = _validate_expression_ast(_cpp_to_python(src), env)
        val = eval(code, {"__builtins__": {}}, env)  # noqa: S307  (validated: arithmetic on whitelisted names only)
        return np.broadcast_to(np.asarray(val, dtype=np.float64), x[0].shape).copy()

    def __call__(self, coords):
        """coords [n, gdim] -> values [n] (scalar expression) or [n, k]."""
        x = [coords[:, i] for i in range(coords.shape[1])]
        if isinstance(self.code, (tuple, list)):
            return np.stack([self._eval_one(str(c), x) for c in self.code], axis=1)
        return self._eval_one(str(self.code), x)


# ------------------------------------------------------------------------------------------ meshes
_HEX_TETS = np.array([(0, 1, 3, 7), (0, 1, 5, 7), (0, 4, 5, 7), (0, 2, 3, 7), (0, 4, 6, 7), (0, 2, 6, 7)])


class Mesh:
    """A simplex mesh.  Either host arrays (file / user supplied) or a dolfin-layout box description
    that is generated directly on the device (UnitCubeMesh(256,256,256) never crosses PCIe)."""

    def __init__(self, coordinates=None, cells=None, box=None, cells_sorted=False):
        """`cells_sorted`: the caller guarantees int32, C-contiguous cells with ascending vertex ids per cell (what mesh.order()
        gives); the arrays are then adopted as they are (e.g. pinned host memory) instead of sorted into a copy."""
        if isinstance(coordinates, str):
            c, t = read_dolfin_xml_mesh(coordinates)
            coordinates, cells = c, t
        self.box = box                      # dict(n=(..), p0=(..), p1=(..)) or None
        self._coords = None if coordinates is None else np.ascontiguousarray(coordinates, dtype=np.float64)
        if cells is not None and cells_sorted and isinstance(cells, np.ndarray) and cells.dtype == np.int32 and cells.flags.c_contiguous:
            self._cells = cells
        else:
            self._cells = None if cells is None else np.sort(np.ascontiguousarray(cells, dtype=np.int32), axis=1)
        if box is None and (self._coords is None or self._cells is None):
            raise SolverError("Mesh needs coordinates and cells, or a box description")
        self._exterior = None
        self._facet_table = None

    # geometry()/topology() give the two numbers the reference reads (SolverBase.py:153-154)
    class _Dim:
        def __init__(self, d):
            self._d = d

        def dim(self):
            return self._d

    def geometry(self):
        return Mesh._Dim(self.gdim)

    def topology(self):
        return Mesh._Dim(self.tdim)

    @property
    def gdim(self):
        return len(self.box["n"]) if self.box else self._coords.shape[1]

    @property
    def tdim(self):
        return len(self.box["n"]) if self.box else self._cells.shape[1] - 1

    def num_vertices(self):
        if self.box:
            return int(np.prod([k + 1 for k in self.box["n"]]))
        return self._coords.shape[0]

    def num_cells(self):
        if self.box:
            return int(np.prod(self.box["n"])) * (6 if self.tdim == 3 else 2)
        return self._cells.shape[0]

    def coordinates(self):
        if self._coords is None:
            n, p0, p1 = self.box["n"], self.box["p0"], self.box["p1"]
            axes = [p0[i] + np.arange(n[i] + 1) * (p1[i] - p0[i]) / n[i] for i in range(len(n))]
            grids = np.meshgrid(*axes[::-1], indexing="ij")          # last axis slowest, x fastest
            self._coords = np.stack([g.ravel() for g in grids[::-1]], axis=1)
        return self._coords

    def cells(self):
        if self._cells is None:
            self._cells = box_cells(self.box["n"])
        return self._cells

    def vertex_coordinates(self, ids):
        """Coordinates of the vertices `ids` ([k, gdim]); a generated box computes them from the ids (the same expression
        coordinates() evaluates, bit for bit) instead of materialising every vertex on the host."""
        ids = np.asarray(ids, dtype=np.int64)
        if self._coords is not None or not self.box:
            return self.coordinates()[ids]
        n, p0, p1 = self.box["n"], self.box["p0"], self.box["p1"]
        out = np.empty((ids.size, len(n)))
        rem = ids
        for i in range(len(n)):
            idx = rem % (n[i] + 1)
            rem = rem // (n[i] + 1)
            out[:, i] = p0[i] + idx * (p1[i] - p0[i]) / n[i]
        return out

    # ---- facets -------------------------------------------------------------------------------
    def device_mesh(self, ctx=None):
        """The whole mesh on this process's GPU — generated there from a box description, uploaded from the host arrays
        otherwise — created once and shared by the boundary search (K1) and the single-GPU DeviceSpace."""
        from . import _lib, backend
        ctx = ctx or backend.get_context()
        dm = self.__dict__.get("_dmesh")
        if dm is None or dm.h is None or dm.ctx is not ctx:
            if self.box and not getattr(self, "force_upload", False):
                dm = _lib.DeviceMesh.box(ctx, self.box["n"], self.box["p0"], self.box["p1"])
            else:
                dm = _lib.DeviceMesh.upload(ctx, self.coordinates(), self.cells())
            self._dmesh = dm
        return dm

    def slab_device_mesh(self, ctx=None):
        """Slab-distributed run on a box layout: this rank's z-slab (its owned vertex planes plus one ghost plane per neighbour)
        on its GPU — generated there, or uploaded from the host arrays when the mesh came as arrays with a box description
        (force_upload).  Returns (DeviceMesh, layout dict); made once per mesh and shared by the boundary search and the
        DeviceSpace.  Vertex ids on the device are local: global id - v_off."""
        from . import _lib, backend
        ctx = ctx or backend.get_context()
        rank, nranks = self.distributed
        hit = self.__dict__.get("_slab")
        if hit is not None and hit[0].h is not None and hit[0].ctx is ctx and hit[1]["rank"] == rank and hit[1]["nranks"] == nranks:
            return hit
        n = self.box["n"]
        nlast = n[-1]
        if nlast + 1 < nranks:
            raise SolverError("more ranks than vertex planes")
        plane = int(np.prod([k + 1 for k in n[:-1]]))
        zp0, zp1 = backend.slab_partition(nlast + 1, nranks)[rank]
        layer0, layer1 = max(zp0 - 1, 0), min(zp1, nlast)
        lay = dict(rank=rank, nranks=nranks, plane=plane, layer0=layer0, layer1=layer1, ghost_lo=int(zp0 > 0), ghost_hi=int(zp1 <= nlast),
                   owned_planes=zp1 - zp0, v_off=layer0 * plane, nlast=nlast)
        if getattr(self, "force_upload", False):
            per_layer = self.num_cells() // nlast
            nvl = (layer1 - layer0 + 1) * plane
            dm = _lib.DeviceMesh.upload(ctx, self.coordinates()[lay["v_off"]:lay["v_off"] + nvl],
                                        self.cells()[per_layer * layer0:per_layer * layer1], vertex_offset=lay["v_off"])
        else:
            dm = _lib.DeviceMesh.box(ctx, n, self.box["p0"], self.box["p1"], layer0, layer1)
        self._slab = (dm, lay)
        return self._slab

    def exterior_facets(self):
        """(fverts[nbf, tdim], opposite_vertex[nbf]) of the exterior facets, in lexicographic order of the vertex tuples.
        Found on the device (libfsb K1, csrc/fsb_facets.cu).  One GPU: the whole mesh.  Slab-distributed box: the search runs
        on this rank's slab and the list holds the boundary facets of that slab only (global vertex ids; the facets of the two
        cut planes, interior to the whole mesh, are dropped) — every consumer restricts facet lists to the rank's vertices
        anyway (DeviceSpace.local_facets).  Host routes that remain: a generated box when this process has no GPU at all
        (direct enumeration of the box surface: host-only mesh inspection) and the RCB-distributed array mesh (np.unique table
        of the replicated host mesh)."""
        if self._exterior is None:
            dist = getattr(self, "distributed", None)
            if self.box and not _device_present():
                self._exterior = box_exterior_facets(self.box["n"])
            elif self.box and dist and getattr(self, "slab_partition", True) and not getattr(self, "force_general_partition", False):
                dm, lay = self.slab_device_mesh()
                fv, opp, cell, _fid = dm.exterior_facets(ids=False)
                zl = fv // lay["plane"]
                nplanes = lay["layer1"] - lay["layer0"] + 1
                cut = np.zeros(fv.shape[0], dtype=bool)
                if lay["layer0"] > 0:
                    cut |= np.all(zl == 0, axis=1)
                if lay["layer1"] < lay["nlast"]:
                    cut |= np.all(zl == nplanes - 1, axis=1)
                keep = ~cut
                self._exterior = ((fv[keep] + lay["v_off"]).astype(np.int32), (opp[keep] + lay["v_off"]).astype(np.int32))
                self._exterior_keep = keep
                self._exterior_cells = cell[keep]
            elif self.box and dist:
                # RCB-partitioned space on a generated box (degree 2): every rank needs the global list; direct enumeration
                self._exterior = box_exterior_facets(self.box["n"])
                self._exterior_cells = None
            elif dist:
                facets, cf, count = self.facet_table()
                ci, li = np.nonzero(count[cf] == 1)
                fid = cf[ci, li]
                order = np.argsort(fid, kind="stable")
                self._exterior = (facets[fid[order]].astype(np.int32), self.cells()[ci[order], li[order]].astype(np.int32), fid[order])
                self._exterior_cells = None
            else:
                fv, opp, cell, _fid = self.device_mesh().exterior_facets(ids=False)
                self._exterior = (fv, opp, "device")          # the ids are ranked on the device when first asked for
                self._exterior_cells = cell
                self._exterior_keep = None
        return self._exterior[0], self._exterior[1]

    def boundary_geometry(self):
        """(finv[nbf, tdim], pts[nbv, gdim], mid[nbf, gdim]) for the facets of exterior_facets(): the facets' vertices as indices into
        the distinct boundary vertices, those vertices' coordinates and the facet midpoints (what SubDomain.mark evaluates on)."""
        geom = self.__dict__.get("_boundary_geometry")
        if geom is None:
            fverts, _ = self.exterior_facets()
            dm = None
            if getattr(self, "_exterior_cells", None) is not None:
                dm = self._slab[0] if getattr(self, "_exterior_keep", None) is not None else self.__dict__.get("_dmesh")
            if dm is not None and dm.h is not None:
                _bv, inv, pts, mid = dm.boundary_geometry()          # made on the device next to the facet search (K1)
                keep = getattr(self, "_exterior_keep", None)
                if keep is not None:
                    inv, mid = inv[keep], mid[keep]
                geom = (inv, pts, mid)
            else:                                                    # host-enumerated box surface / RCB table
                flag = np.zeros(self.num_vertices(), dtype=bool)
                flag[fverts.ravel()] = True
                uv = np.flatnonzero(flag)                            # distinct boundary vertices, ascending
                rank = np.empty(flag.size, dtype=np.int32)
                rank[uv] = np.arange(uv.size, dtype=np.int32)
                inv = rank[fverts]
                pts = self.vertex_coordinates(uv)
                mid = pts[inv[:, 0]].copy()
                for j in range(1, inv.shape[1]):
                    mid += pts[inv[:, j]]
                geom = (inv, pts, mid / inv.shape[1])
            self._boundary_geometry = geom
        return geom

    def locate_point(self, p, tol=1e-12):
        """(cell index, barycentric coordinates) of a cell containing point p (the first one in cell order, as a
        linear search would find it); SolverError if p is outside the mesh."""
        x = np.asarray(p.x if isinstance(p, Point) else p, dtype=np.float64)[:self.gdim]
        c, t = self.coordinates(), self.cells()
        d = self.tdim
        best = None
        for s0 in range(0, t.shape[0], 1 << 20):                     # chunks bound the temporary memory
            tc = t[s0:s0 + (1 << 20)]
            X = c[tc]                                                # [n, d+1, d]
            J = np.transpose(X[:, 1:, :] - X[:, :1, :], (0, 2, 1))
            lam = np.linalg.solve(J, np.broadcast_to(x - X[:, 0, :], (tc.shape[0], d))[..., None])[..., 0]
            lam = np.concatenate([1.0 - lam.sum(axis=1, keepdims=True), lam], axis=1)
            inside = np.nonzero(lam.min(axis=1) >= -tol)[0]
            if inside.size:
                best = (s0 + int(inside[0]), lam[inside[0]])
                break
        if best is None:
            raise SolverError("point %s is outside the mesh" % (x,))
        return best

    def exterior_facet_ids(self):
        """dolfin facet indices of the exterior facets (rank of the sorted vertex tuple among all distinct facets)."""
        self.exterior_facets()
        if len(self._exterior) < 3:
            raise SolverError("global facet numbering is not materialised on this route (slab-distributed box / no GPU)")
        if isinstance(self._exterior[2], str):
            self._exterior = (self._exterior[0], self._exterior[1], self.device_mesh().exterior_facet_ids())
        return self._exterior[2]

    def facet_table(self):
        """All facets in dolfin numbering: local facet i is opposite local vertex i of the sorted cell;
        global facet id = lexicographic rank of the sorted facet vertex tuple (verified against
        data/mesh_facet_region.xml).  Returns (facets, cell_facets, count)."""
        if self._facet_table is None:
            cells = self.cells()
            nc, nl = cells.shape
            allf = np.stack([np.delete(cells, i, axis=1) for i in range(nl)], axis=1).reshape(-1, nl - 1)
            facets, inv, count = np.unique(allf, axis=0, return_inverse=True, return_counts=True)
            self._facet_table = (facets, inv.reshape(nc, nl), count)
        return self._facet_table


_DEVICE_PRESENT = None


def _device_present():
    """True when this process can create a libfsb context (a CUDA device is visible); probed once."""
    global _DEVICE_PRESENT
    if _DEVICE_PRESENT is None:
        try:
            from . import backend
            backend.get_context()
            _DEVICE_PRESENT = True
        except Exception:
            _DEVICE_PRESENT = False
    return _DEVICE_PRESENT


def box_cells(n):
    """Cells of the dolfin box layout, sorted per cell (BoxMesh: six tets per hex on the v0-v7
    diagonal; RectangleMesh 'right': (v0,v1,v3),(v0,v2,v3))."""
    if len(n) == 2:
        nx, ny = n
        cx, cy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
        v0 = (cy * (nx + 1) + cx).ravel()
        v1, v2, v3 = v0 + 1, v0 + nx + 1, v0 + nx + 2
        cells = np.stack([np.stack([v0, v1, v3], 1), np.stack([v0, v2, v3], 1)], axis=1).reshape(-1, 3)
        return cells.astype(np.int32)
    nx, ny, nz = n
    px, py = nx + 1, (nx + 1) * (ny + 1)
    cz, cy, cx = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    v0 = (cx + cy * px + cz * py).ravel()
    hv = np.stack([v0, v0 + 1, v0 + px, v0 + px + 1, v0 + py, v0 + py + 1, v0 + py + px, v0 + py + px + 1], axis=1)
    return hv[:, _HEX_TETS].reshape(-1, 4).astype(np.int32)


def _box_face_templates(d):
    """For one box of the dolfin layout (local vertex k: bit 0 = x, bit 1 = y, bit 2 = z): the simplex facets lying in each of
    its 2 d faces, as {(axis, side): [(facet local vertices ascending, opposite local vertex), ...]}."""
    tets = _HEX_TETS if d == 3 else np.array([(0, 1, 3), (0, 2, 3)])
    out = {}
    for tet in tets:
        for i in range(d + 1):
            f = tuple(int(v) for k, v in enumerate(tet) if k != i)
            for axis in range(d):
                bits = {(v >> axis) & 1 for v in f}
                if len(bits) == 1:
                    out.setdefault((axis, bits.pop()), []).append((f, int(tet[i])))
    return out


def box_exterior_facets(n):
    """Exterior facets of the box layout by direct enumeration: every boundary face of a boundary box contributes the simplex
    facets its template (one box) has in that face — index arithmetic on the known layout, O(surface), no search.  Returns
    (fverts, opposite vertex) in lexicographic order of the vertex tuples."""
    d = len(n)
    strides = np.cumprod([1] + [k + 1 for k in n[:-1]]).astype(np.int64)
    off = np.array([sum(((k >> a) & 1) * int(strides[a]) for a in range(d)) for k in range(1 << d)], dtype=np.int64)
    tmpl = _box_face_templates(d)
    fv_all, opp_all = [], []
    for axis in range(d):
        other = [a for a in range(d) if a != axis]
        grids = np.meshgrid(*[np.arange(n[a], dtype=np.int64) for a in other[::-1]], indexing="ij")
        base_other = sum(g.ravel() * strides[a] for g, a in zip(grids[::-1], other)) if other else np.zeros(1, dtype=np.int64)
        for side in (0, 1):
            v0 = base_other + (0 if side == 0 else (n[axis] - 1)) * strides[axis]
            for f, o in tmpl[(axis, side)]:
                fv_all.append(v0[:, None] + off[list(f)][None, :])
                opp_all.append(v0 + off[o])
    fv = np.concatenate(fv_all)
    opp = np.concatenate(opp_all)
    # lexicographic order through one 64-bit key per facet when the ids fit (3 x 21 bits), else a lexsort
    nv = int(np.prod([k + 1 for k in n]))
    if nv < (1 << 21):
        key = fv[:, 0]
        for j in range(1, d):
            key = (key << 21) | fv[:, j]
        order = np.argsort(key, kind="stable")
    else:
        order = np.lexsort(fv.T[::-1])
    return fv[order].astype(np.int32), opp[order].astype(np.int32)


def UnitSquareMesh(nx, ny, diagonal="right"):
    if diagonal != "right":
        raise SolverError("only the default 'right' diagonal is implemented")
    return Mesh(box=dict(n=(int(nx), int(ny)), p0=(0.0, 0.0), p1=(1.0, 1.0)))


def RectangleMesh(p0, p1, nx, ny, diagonal="right"):
    if diagonal != "right":
        raise SolverError("only the default 'right' diagonal is implemented")
    return Mesh(box=dict(n=(int(nx), int(ny)), p0=(float(p0[0]), float(p0[1])), p1=(float(p1[0]), float(p1[1]))))


def UnitCubeMesh(nx, ny, nz):
    return Mesh(box=dict(n=(int(nx), int(ny), int(nz)), p0=(0.0, 0.0, 0.0), p1=(1.0, 1.0, 1.0)))


def BoxMesh(p0, p1, nx, ny, nz):
    return Mesh(box=dict(n=(int(nx), int(ny), int(nz)), p0=tuple(float(p0[i]) for i in range(3)), p1=tuple(float(p1[i]) for i in range(3))))


def read_dolfin_xml_mesh(path):
    """dolfin-XML mesh (Mesh(filename), SolverBase.py:224).  Cells are sorted per cell as mesh.order() does."""
    txt = open(path, "r").read()
    m = re.search(r'<mesh[^>]*celltype="(\w+)"[^>]*dim="(\d+)"', txt)
    if not m:
        raise SolverError("%s is not a dolfin-XML mesh" % path)
    celltype, dim = m.group(1), int(m.group(2))
    if celltype not in ("tetrahedron", "triangle"):
        raise SolverError("cell type %s is not supported" % celltype)
    nv = int(re.search(r'<vertices size="(\d+)"', txt).group(1))
    nc = int(re.search(r'<cells size="(\d+)"', txt).group(1))
    names = ["x", "y", "z"][:dim]
    vre = re.compile(r'<vertex index="(\d+)"' + "".join(r'\s+%s="([^"]+)"' % k for k in names))
    varr = np.array([[float(g) for g in mm.groups()] for mm in vre.finditer(txt)])
    coords = np.zeros((nv, dim))
    coords[varr[:, 0].astype(np.int64)] = varr[:, 1:]
    nl = 4 if celltype == "tetrahedron" else 3
    cre = re.compile(r'<%s index="(\d+)"' % celltype + "".join(r'\s+v%d="(\d+)"' % k for k in range(nl)))
    carr = np.array([[int(g) for g in mm.groups()] for mm in cre.finditer(txt)], dtype=np.int64)
    cells = np.zeros((nc, nl), dtype=np.int32)
    cells[carr[:, 0]] = carr[:, 1:]
    return coords, np.sort(cells, axis=1)


def read_xdmf_mesh(path):
    """XDMF mesh as dolfin's XDMFFile writes it (SolverBase.py:246-252 `XDMFFile(...).read(mesh, True)`), for files whose heavy data is
    inline (`Format="XML"`, dolfin's `XDMFFile.Encoding.ASCII`).  Files that point into an HDF5 container cannot be read here: the
    image has no HDF5 library."""
    import xml.etree.ElementTree as ET
    root = ET.parse(path).getroot()
    grid = root.find(".//Grid")
    if grid is None:
        raise SolverError("%s holds no XDMF Grid" % path)
    topo, geom = grid.find("Topology"), grid.find("Geometry")
    if topo is None or geom is None:
        raise SolverError("%s: Grid without Topology/Geometry" % path)
    kind = (topo.get("TopologyType") or topo.get("Type") or "").lower()
    nl = {"tetrahedron": 4, "triangle": 3}.get(kind)
    if nl is None:
        raise SolverError("XDMF topology %r is not supported (Triangle / Tetrahedron)" % kind)

    def data(node, dtype):
        item = node.find("DataItem")
        if item is None:
            raise SolverError("%s: missing DataItem" % path)
        if (item.get("Format") or "XML").upper() != "XML":
            raise SolverError("%s stores its arrays in HDF5 (Format=%r); no HDF5 library is available here: re-save with "
                              "XDMFFile.Encoding.ASCII or convert to dolfin-XML" % (path, item.get("Format")))
        dims = [int(k) for k in (item.get("Dimensions") or "").split()]
        a = np.array((item.text or "").split(), dtype=dtype)
        return a.reshape(dims) if dims and int(np.prod(dims)) == a.size else a
    cells = data(topo, np.int64).reshape(-1, nl)
    coords = data(geom, np.float64)
    gdim = 3 if (geom.get("GeometryType") or "XYZ").upper() == "XYZ" else 2
    coords = coords.reshape(-1, gdim)
    if nl == 3 and gdim == 3 and np.all(coords[:, 2] == coords[0, 2]):
        coords = coords[:, :2]                     # planar triangles written with a z column
    if cells.min() < 0 or cells.max() >= coords.shape[0]:
        raise SolverError("%s: connectivity out of range" % path)
    return coords, np.sort(cells.astype(np.int32), axis=1)


def read_mesh_function_xml(path):
    """dolfin-XML MeshFunction (SolverBase.py:229,236) -> (dim, int values[size])."""
    txt = open(path, "r").read()
    m = re.search(r'<mesh_function type="\w+" dim="(\d+)" size="(\d+)"', txt)
    if not m:
        raise SolverError("%s is not a dolfin-XML mesh function" % path)
    dim, size = int(m.group(1)), int(m.group(2))
    arr = np.array([[int(a), int(b)] for a, b in re.findall(r'<entity index="(\d+)" value="(-?\d+)"', txt)], dtype=np.int64)
    vals = np.zeros(size, dtype=np.int64)
    if arr.size:
        vals[arr[:, 0]] = arr[:, 1]
    return dim, vals


# ------------------------------------------------------------------------------------------ subdomains / markers
class SubDomain:
    """dolfin.SubDomain: override inside(x, on_boundary)."""

    def inside(self, x, on_boundary):
        raise NotImplementedError

    def _call(self, x, on_boundary):
        return self.inside(x, on_boundary)

    def mark(self, markers, value):
        markers.mark_subdomain(self, value)


class AutoSubDomain(SubDomain):
    """dolfin.AutoSubDomain(lambda x: ...) or (lambda x, on_boundary: ...)."""

    def __init__(self, fn):
        self.fn = fn
        try:
            import inspect
            self.nargs = len(inspect.signature(fn).parameters)
        except (TypeError, ValueError):
            self.nargs = 1

    def inside(self, x, on_boundary):
        return self.fn(x) if self.nargs == 1 else self.fn(x, on_boundary)


def _evaluate_predicate(sub, pts, on_boundary=True):
    """inside() over points [n, gdim]: vectorised call with x[i] = coordinate arrays, falling back to a
    per-point loop when the predicate is not array-friendly."""
    if not isinstance(sub, SubDomain):
        if callable(sub):
            sub = AutoSubDomain(sub)
        else:
            raise SolverError("a boundary must be a SubDomain or a predicate, got %r" % type(sub))
    n = pts.shape[0]
    try:
        r = sub._call([pts[:, i] for i in range(pts.shape[1])], on_boundary)
        r = np.asarray(r)
        if r.shape == (n,):
            return r.astype(bool)
        if r.shape == ():
            return np.full(n, bool(r))
    except (ValueError, TypeError):
        pass
    return np.array([bool(sub._call(pts[i], on_boundary)) for i in range(n)], dtype=bool)


class FacetMarkers:
    """MeshFunction('size_t', mesh, tdim-1) restricted to the exterior facets (deviation from dolfin:
    interior facets are never marked; every boundary in the reference's cases lies on the surface)."""

    def __init__(self, mesh):
        self.mesh = mesh
        self.fverts, self.opp = mesh.exterior_facets()
        self._cache = {}
        self.version = 0          # bumped by every marking call; keys the per-marker caches (transient loops ask each step)
        self.values = np.zeros(self.fverts.shape[0], dtype=np.int64)

    @property
    def values(self):
        return self._values

    @values.setter
    def values(self, v):
        self._values = v
        self.touch()

    def touch(self):
        """Call after editing `values` in place (set_all / SubDomain.mark do it themselves)."""
        self.version += 1
        self._cache.clear()

    def set_all(self, v):
        self._values[:] = v
        self.touch()

    def mark_subdomain(self, sub, value):
        """SubDomain.mark: a facet is marked iff all its vertices and its midpoint are inside
        (dolfin's default check_midpoint=True; SolverBase.py:281-282)."""
        geom = self.mesh.boundary_geometry()
        inv, pts, mid = geom
        ok = _evaluate_predicate(sub, mid)
        pv = _evaluate_predicate(sub, pts)
        for j in range(inv.shape[1]):                     # all vertices inside: one 1-D gather per facet vertex
            ok &= pv[inv[:, j]]
        self._values[ok] = value
        self.touch()

    def facets(self, marker):
        hit = self._cache.get(marker)
        if hit is None:
            sel = self._values == marker
            hit = self._cache[marker] = (self.fverts[sel], self.opp[sel])
        return hit

    def vertices(self, marker):
        return np.unique(self.fverts[self.values == marker])

    def array(self):
        return self.values


class MeshFunction:
    """Cell markers (MeshFunction('size_t', mesh, tdim)); facet markers are FacetMarkers."""

    def __new__(cls, kind, mesh, dim_or_file, value=0):
        if isinstance(dim_or_file, str):
            dim, vals = read_mesh_function_xml(dim_or_file)
        else:
            dim, vals = int(dim_or_file), None
        if dim == mesh.tdim - 1:
            fm = FacetMarkers(mesh)
            if vals is not None:
                fm.values = vals[mesh.exterior_facet_ids()]
            return fm
        obj = super().__new__(cls)
        obj.mesh, obj.dim = mesh, dim
        obj.values = vals if vals is not None else np.full(mesh.num_cells() if dim == mesh.tdim else mesh.num_vertices(), value, dtype=np.int64)
        return obj

    def set_all(self, v):
        self.values[:] = v

    def mark_subdomain(self, sub, value):
        """SubDomain.mark on a cell (or vertex) function: an entity is marked iff all its vertices and its midpoint are inside
        (dolfin's default check_midpoint=True); on_boundary is False for cells."""
        c = self.mesh.coordinates()
        if self.dim == 0:
            self.values[_evaluate_predicate(sub, c, on_boundary=False)] = value
            return
        t = self.mesh.cells()
        inside_v = _evaluate_predicate(sub, c, on_boundary=False)
        mid = c[t].mean(axis=1)
        ok = inside_v[t].all(axis=1) & _evaluate_predicate(sub, mid, on_boundary=False)
        self.values[ok] = value

    def array(self):
        return self.values


# ------------------------------------------------------------------------------------------ spaces / functions
class _UflElement:
    def __init__(self, degree):
        self._degree = degree

    def degree(self):
        return self._degree


_UFC_EDGES = {1: ((0, 1),), 2: ((1, 2), (0, 2), (0, 1)), 3: ((2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1))}


class FunctionSpace:
    """P1 or P2 Lagrange on a mesh; ncomp = 1 (scalar) or dim (vector).

    Node numbering: vertices in vertex order, then (degree 2) one node per edge, edges numbered
    by the lexicographic rank of their sorted vertex pair; vector dof = ncomp*node + component.  Local node
    order of a cell: its vertices sorted ascending, then its edges in UFC order."""

    def __init__(self, mesh, family="CG", degree=1, ncomp=1, constrained_domain=None):
        if family not in ("CG", "Lagrange", "P"):
            raise SolverError("element family %r is not supported (CG/Lagrange only)" % family)
        if degree not in (1, 2):
            raise SolverError("fe_degree %r is not implemented: P1 and P2 run on the device path" % degree)
        if constrained_domain is not None:
            raise SolverError("periodic boundaries are not implemented")
        self._mesh, self.ncomp, self.degree = mesh, ncomp, degree
        self._ufl_element = _UflElement(degree)
        self._edges = self._cell_nodes = self._node_coords = None

    def mesh(self):
        return self._mesh

    def num_nodes(self):
        if self.degree == 1:
            return self._mesh.num_vertices()
        return self._mesh.num_vertices() + self.edges().shape[0]

    def dim(self):
        return self.num_nodes() * self.ncomp

    # ---- degree 2: edge nodes (host integer work, K1) ----------------------------------------------
    def edges(self):
        if self._edges is None:
            cells = self._mesh.cells().astype(np.int64)
            nv = self._mesh.num_vertices()
            loc = _UFC_EDGES[self._mesh.tdim]
            keys = np.stack([cells[:, a] * nv + cells[:, b] for a, b in loc], axis=1)       # cells are sorted: a < b
            uniq, inv = np.unique(keys.ravel(), return_inverse=True)
            self._edge_keys = uniq
            self._edges = np.stack([uniq // nv, uniq % nv], axis=1)
            self._cell_nodes = np.hstack([cells, nv + inv.reshape(keys.shape)]).astype(np.int32)
        return self._edges

    def cell_nodes(self):
        if self.degree == 1:
            return self._mesh.cells()
        self.edges()
        return self._cell_nodes

    def node_coordinates_at(self, nodes):
        """Coordinates of the given nodes only (degree 1 on a generated box: from the ids, no host copy of the whole mesh)."""
        if self.degree == 1:
            return self._mesh.vertex_coordinates(nodes)
        return self.node_coordinates()[np.asarray(nodes, dtype=np.int64)]

    def node_coordinates(self):
        c = self._mesh.coordinates()
        if self.degree == 1:
            return c
        if self._node_coords is None:
            e = self.edges()
            self._node_coords = np.vstack([c, 0.5 * (c[e[:, 0]] + c[e[:, 1]])])
        return self._node_coords

    def point_weights(self, p):
        """(nodes, basis values) of the cell containing point p: what PointSource adds to the right-hand side."""
        cell, lam = self._mesh.locate_point(p)
        nodes = self.cell_nodes()[cell].astype(np.int64)
        if self.degree == 1:
            return nodes, lam
        phi = [l * (2.0 * l - 1.0) for l in lam] + [4.0 * lam[a] * lam[b] for a, b in _UFC_EDGES[self._mesh.tdim]]
        return nodes, np.array(phi)

    def facet_nodes(self, fverts):
        """Nodes of facets given by their (sorted) vertices: the vertices, then (degree 2) the facet's edges."""
        fv = np.asarray(fverts, dtype=np.int64)
        if self.degree == 1 or fv.shape[0] == 0:
            return fv.astype(np.int32) if self.degree == 1 else np.zeros((0, fv.shape[1] * (fv.shape[1] + 1) // 2), dtype=np.int32)
        self.edges()
        nv = self._mesh.num_vertices()
        fv = np.sort(fv, axis=1)
        loc = _UFC_EDGES[fv.shape[1] - 1]
        keys = np.stack([fv[:, a] * nv + fv[:, b] for a, b in loc], axis=1)
        pos = np.searchsorted(self._edge_keys, keys)
        if not np.array_equal(self._edge_keys[np.minimum(pos, self._edge_keys.size - 1)], keys):
            raise SolverError("facet edge not found in the mesh edge table")
        return np.hstack([fv, nv + pos]).astype(np.int32)


def VectorFunctionSpace(mesh, family="CG", degree=1, dim=None, constrained_domain=None):
    return FunctionSpace(mesh, family, degree, ncomp=dim or mesh.gdim, constrained_domain=constrained_domain)


class _Vector:
    def __init__(self, fn):
        self._fn = fn

    def get_local(self):
        """A copy of the values (dolfin semantics).  A device-resident result is downloaded straight into the
        returned array: one D2H copy, no second host copy."""
        fn = self._fn
        if not fn._host_valid and fn._dev is not None:
            return fn._dev.numpy()
        return fn.array().copy()

    def array(self):
        return self._fn.array()

    def __getitem__(self, i):
        return self._fn.array()[i]

    def __setitem__(self, i, v):
        a = self._fn.array()
        a[i] = v
        self._fn.assign_array(a)

    def size(self):
        return self._fn.function_space.dim()

    def norm(self, kind="l2"):
        a = self._fn.array()
        kind = kind.lower()
        if kind == "l2":
            return float(np.linalg.norm(a))
        if kind == "l1":
            return float(np.abs(a).sum())
        if kind == "linf":
            return float(np.abs(a).max()) if a.size else 0.0
        raise SolverError("unknown vector norm %r (l1, l2, linf)" % kind)

    def set_local(self, values):
        self._fn.assign_array(np.asarray(values, dtype=np.float64))

    def apply(self, mode="insert"):
        return None                       # dolfin finalises assembly / ghost updates here; nothing to do

    def max(self):
        return float(self._fn.array().max())

    def min(self):
        return float(self._fn.array().min())

    def sum(self):
        return float(self._fn.array().sum())

    def __len__(self):
        return self.size()


class Function:
    """Nodal P1 field.  Values live on the device after a solve and are downloaded lazily."""

    def __init__(self, V, values=None, fill=None):
        """`fill`: a uniform value (scalar space) kept symbolic until someone asks for the array, so a
        constant initial field costs neither host memory traffic nor an H2D copy."""
        self.function_space = V
        self._host = None if values is None else np.ascontiguousarray(values, dtype=np.float64).ravel().copy()
        self._dev = None          # _lib.DeviceVector holding the same values (None: host only)
        self._host_valid = True
        self._fill = None
        if self._host is None:
            self._fill = 0.0 if fill is None else float(fill)
        self._name = "f"

    def uniform_value(self):
        """The value if the field is a known constant (and nothing else has been stored), else None."""
        return self._fill if (self._host is None and self._dev is None) else None

    def array(self):
        if not self._host_valid:
            if getattr(self, "_local", False):
                raise SolverError("this Function holds one rank's part of a distributed solution (solver_settings['gather_result'] = False): "
                                  "the global nodal array was never assembled; use solver.local_result() or gather_result=True")
            self._host = self._dev.numpy()       # the one D2H copy of a solve result, on demand
            self._host_valid = True
        elif self._host is None:
            self._host = np.full(self.function_space.dim(), self._fill)
        return self._host

    def assign_array(self, a):
        self._host = np.ascontiguousarray(a, dtype=np.float64).ravel().copy()
        self._host_valid = True
        self._dev = None
        self._local = False

    def set_device(self, dev, local=False):
        """Adopt a device vector as the current value (host copy becomes stale).  `local`: the vector is one rank's part
        (owned + ghost nodes) of a distributed field, not the global nodal array."""
        self._dev = dev
        self._host_valid = False
        self._host = None
        self._local = bool(local)

    def device_vector(self):
        return self._dev

    def assign(self, other):
        if isinstance(other, Function):
            if other.uniform_value() is not None:
                self._host, self._dev, self._host_valid, self._fill = None, None, True, other._fill
                self._local = False
            elif other._dev is not None and not other._host_valid:
                # device-resident value: copy on the device, no PCIe round trip per time step
                from ._lib import DeviceVector
                if self._dev is None or self._dev is other._dev or self._dev.n != other._dev.n:
                    self._dev = DeviceVector(other._dev.ctx, other._dev.n)
                self._dev.copy_from(other._dev)
                self._host_valid = False
                self._host = None
                self._local = getattr(other, "_local", False)
            else:
                self.assign_array(other.array())
        else:
            raise SolverError("Function.assign needs another Function")

    def vector(self):
        return _Vector(self)

    def compute_vertex_values(self, mesh=None):
        a = self.array()
        nc = self.function_space.ncomp
        nv = self.function_space.mesh().num_vertices()               # P2: the vertex nodes come first
        return a[:nv].copy() if nc == 1 else a.reshape(-1, nc)[:nv].T.reshape(-1)     # dolfin returns component-major

    @property
    def values(self):
        a = self.array()
        nc = self.function_space.ncomp
        return a if nc == 1 else a.reshape(-1, nc)

    def copy(self, deepcopy=True):
        if self.uniform_value() is not None:
            return Function(self.function_space, fill=self._fill)
        return Function(self.function_space, self.array())

    def rename(self, name, label=""):
        self._name = name

    def __call__(self, *x):
        """u(x, y[, z]) or u(Point) / u((x, y, z)): the value at a point (a number, or one per component), from the basis functions
        of the cell that contains it; SolverError outside the mesh."""
        p = x[0] if len(x) == 1 else x
        nodes, w = self.function_space.point_weights(p)
        nc = self.function_space.ncomp
        a = self.array()
        if nc == 1:
            return float(np.dot(w, a[nodes]))
        return np.dot(w, a.reshape(-1, nc)[nodes])


class File:
    """File("name.pvd") << u  /  << (u, t): ParaView output as dolfin's File writes it (SolverBase.py:570-577): one `.vtu` piece per
    write, the `.pvd` collection lists them.  Also `.vtu` alone."""

    def __init__(self, filename, encoding=None):
        self.filename, self.series = filename, []

    def __lshift__(self, what):
        from .SolverBase import write_pvd, write_vtu
        u, t = what if isinstance(what, tuple) else (what, float(len(self.series)))
        mesh = u.function_space.mesh()
        if self.filename.endswith(".pvd"):
            import os.path
            piece = "%s%06d.vtu" % (self.filename[:-4], len(self.series))
            write_vtu(piece, mesh, u.values, getattr(u, "_name", "f"))
            self.series.append((float(t), os.path.basename(piece)))
            write_pvd(self.filename, self.series)
        elif self.filename.endswith(".vtu"):
            write_vtu(self.filename, mesh, u.values, getattr(u, "_name", "f"))
        else:
            raise SolverError("File supports .pvd and .vtu")
        return self


def boundary_flux(u, markers, marker_id, coefficient=1.0):
    """assemble(coefficient * dot(grad(u), n) * ds(marker_id)) for a scalar degree-1 Function: what the reference's examples print after
    the solve (examples/test_heat_transfer.py:181-190 post_process, test_electrostatics.py:126-135).  Boundary-only host work: the
    gradient of the cell behind each marked facet, its outward unit normal and the facet measure."""
    V = u.function_space
    if V.ncomp != 1 or V.degree != 1:
        raise SolverError("boundary_flux is implemented for scalar degree-1 functions")
    mesh = V.mesh()
    c = mesh.coordinates()
    fverts, opp = markers.facets(marker_id)
    if len(fverts) == 0:
        return 0.0
    fverts, opp = np.asarray(fverts, dtype=np.int64), np.asarray(opp, dtype=np.int64)
    d = fverts.shape[1]
    cells = np.hstack([fverts, opp[:, None]])
    X = c[cells]
    J = np.transpose(X[:, 1:, :] - X[:, :1, :], (0, 2, 1))
    Jinv = np.linalg.inv(J)
    G = np.concatenate([-Jinv.sum(axis=1, keepdims=True), Jinv], axis=1)          # gradients of the barycentric coordinates
    grad = np.einsum("fa,fai->fi", u.array()[cells], G)
    Xf = c[fverts]
    if d == 2:
        tvec = Xf[:, 1] - Xf[:, 0]
        n = np.stack([tvec[:, 1], -tvec[:, 0]], axis=1)
        meas = np.linalg.norm(tvec, axis=1)
    else:
        n = np.cross(Xf[:, 1] - Xf[:, 0], Xf[:, 2] - Xf[:, 0])
        meas = 0.5 * np.linalg.norm(n, axis=1)
    n = n / np.linalg.norm(n, axis=1, keepdims=True)
    n[np.einsum("ij,ij->i", n, c[opp] - Xf[:, 0]) > 0] *= -1
    return float(coefficient * np.sum(meas * np.einsum("fi,fi->f", grad, n)))


class PointSource:
    """PointSource(V, point, magnitude): adds magnitude * phi_a(point) to the right-hand side
    (ScalarTransportSolver.py:150-158; applied with bc.apply(b), SolverBase.py:598-602)."""

    def __init__(self, V, p, magnitude=1.0):
        self.V, self.point, self.magnitude = V, p, float(magnitude)

    def entries(self):
        nodes, w = self.V.point_weights(self.point)
        return nodes, self.magnitude * np.asarray(w)


class DirichletBC:
    """DirichletBC(V, value, markers, id) (topological): the dofs of every facet carrying the marker.
    `component` restricts a vector space to V.sub(component)."""

    def __init__(self, V, value, markers, marker_id, component=None):
        self.V, self.value, self.markers, self.marker_id, self.component = V, value, markers, marker_id, component

    def dofs_and_values(self, coords):
        # `coords` are the node coordinates of V (vertices, then edge midpoints for degree 2)
        cache = self.V.__dict__.setdefault("_bc_nodes", {})
        key = (id(self.markers), self.marker_id, getattr(self.markers, "version", None))
        verts = cache.get(key)
        if verts is None:
            if len(cache) > 64:
                cache.clear()
            verts = cache[key] = np.unique(self.V.facet_nodes(self.markers.facets(self.marker_id)[0])).astype(np.int64)
        nc = self.V.ncomp
        comps = range(nc) if self.component is None else [self.component]
        comps = list(comps)
        val = self.value
        if isinstance(val, Constant):
            val = val.values()
            val = val[0] if val.size == 1 else val
        at_nodes = isinstance(val, Expression)
        if at_nodes:
            val = val(coords[verts])                # evaluated at the constrained nodes only (coords may be a lazy view)
        elif isinstance(val, Function):
            val = val.values
        val = np.asarray(val, dtype=np.float64)
        nv = verts.size if at_nodes else coords.shape[0]
        pick = (lambda a: a) if at_nodes else (lambda a: a[verts])
        dofs, vals = [], []
        for k, c in enumerate(comps):
            dofs.append(verts * nc + c)
            if val.ndim == 0:                                        # one constant
                v = np.full(verts.size, float(val))
            elif val.ndim == 1 and len(comps) > 1 and val.size == len(comps) and not at_nodes:   # constant vector
                v = np.full(verts.size, val[k])
            elif val.ndim == 1 and val.size == nv:                   # nodal scalar field
                v = pick(val)
            elif val.ndim == 2 and val.shape[0] == nv:               # nodal vector field
                v = pick(val)[:, c if self.component is None else 0]
            else:
                raise SolverError("cannot interpret a Dirichlet value of shape %r" % (val.shape,))
            vals.append(v)
        return np.concatenate(dofs), np.concatenate(vals)
