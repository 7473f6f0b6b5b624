"""CPU oracle: a numpy/scipy restatement of the FenicsSolver hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``fenicssolver_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline leg use it, and only as the checker.

PARITY UNPINNED at the dolfin boundary: the reference (qingfengxia/FenicsSolver)
delegates all arithmetic to dolfin/FFC/PETSc (FEniCS 2019.1 from the Ubuntu PPA,
versions unpinned, ``/root/reference/.travis.yml:30-40``), none of which is
installable here, and its own tests assert no numbers.  What pins this oracle
instead (``tests/test_oracle_kat.py``):

* the shipped fixture ``data/mesh.xml`` + ``data/TestHeatTransfer.json``
  (exact discrete answer T = 350 - 2.5 z),
* the facet-numbering rule checked against ``data/mesh_facet_region.xml``,
* analytic patch tests / closed forms on dolfin-layout cube meshes.

Restated reference call sites (file:line under /root/reference/FenicsSolver):

* weak forms            ScalarTransportSolver.py:228-311, LinearElasticitySolver.py:62-69,206-245
* boundary conditions   ScalarTransportSolver.py:142-211, LinearElasticitySolver.py:99-204
* assemble / solve      SolverBase.py:592-613 (assemble, bc.apply, LU), :643-672 (assemble_system, CG)

DoF numbering: vertex order (scalar dof = v, vector dof = dim*v + c); CSR columns
sorted ascending; structural zeros kept (dolfin's pattern is topology based).
"""
from __future__ import annotations

import math
import re

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

# --------------------------------------------------------------------------- meshes


def rectangle_mesh(x0, y0, x1, y1, nx, ny):
    """dolfin RectangleMesh, diagonal "right": quad (v0,v1,v2,v3) -> (v0,v1,v3),(v0,v2,v3)."""
    ix, iy = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), indexing="xy")
    coords = np.stack([x0 + ix.ravel() * (x1 - x0) / nx, y0 + iy.ravel() * (y1 - y0) / ny], axis=1)
    cx, cy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    v0 = (cy * (nx + 1) + cx).ravel()
    v1, v2, v3 = v0 + 1, v0 + nx + 1, v0 + nx + 2
    cells = np.empty((v0.size, 2, 3), dtype=np.int64)
    cells[:, 0] = np.stack([v0, v1, v3], axis=1)
    cells[:, 1] = np.stack([v0, v2, v3], axis=1)
    cells = np.sort(cells.reshape(-1, 3), axis=1)
    return coords.astype(np.float64), cells.astype(np.int32)


def unit_square_mesh(nx, ny):
    return rectangle_mesh(0.0, 0.0, 1.0, 1.0, nx, ny)


_HEX_TETS = ((0, 1, 3, 7), (0, 1, 7, 5), (0, 5, 7, 4), (0, 3, 2, 7), (0, 6, 4, 7), (0, 2, 6, 7))


def box_mesh(p0, p1, nx, ny, nz):
    """dolfin BoxMesh layout: vertex id = ix + iy(nx+1) + iz(nx+1)(ny+1); six tets per hex
    sharing the v0-v7 diagonal, cell id = 6*hex + k, each cell sorted ascending."""
    px, py = nx + 1, (nx + 1) * (ny + 1)
    iz, iy, ix = np.meshgrid(np.arange(nz + 1), np.arange(ny + 1), np.arange(nx + 1), indexing="ij")
    coords = np.stack([p0[0] + ix.ravel() * (p1[0] - p0[0]) / nx,
                       p0[1] + iy.ravel() * (p1[1] - p0[1]) / ny,
                       p0[2] + iz.ravel() * (p1[2] - p0[2]) / nz], axis=1)
    cz, cy, cx = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    v0 = (cx + cy * px + cz * py).ravel()
    hv = np.stack([v0, v0 + 1, v0 + px, v0 + px + 1, v0 + py, v0 + py + 1, v0 + py + px, v0 + py + px + 1], axis=1)
    cells = np.stack([hv[:, list(t)] for t in _HEX_TETS], axis=1).reshape(-1, 4)
    cells = np.sort(cells, axis=1)
    return coords.astype(np.float64), cells.astype(np.int32)


def unit_cube_mesh(nx, ny, nz):
    return box_mesh((0.0, 0.0, 0.0), (1.0, 1.0, 1.0), nx, ny, nz)


def read_dolfin_xml_mesh(path):
    """dolfin-XML mesh reader (SolverBase.py:223-226 `Mesh(filename)`); cells sorted per cell
    as dolfin's mesh.order() does."""
    txt = open(path, "r").read()
    dim = int(re.search(r'<mesh[^>]*dim="(\d+)"', txt).group(1))
    celltype = re.search(r'celltype="(\w+)"', txt).group(1)
    nv = int(re.search(r'<vertices size="(\d+)"', txt).group(1))
    nc = int(re.search(r'<cells size="(\d+)"', txt).group(1))
    coords = np.zeros((nv, dim))
    names = ["x", "y", "z"][:dim]
    vre = re.compile(r'<vertex index="(\d+)"' + "".join(r'\s+%s="([^"]+)"' % n for n in names))
    for m in vre.finditer(txt):
        coords[int(m.group(1))] = [float(m.group(2 + k)) for k in range(dim)]
    tag = {"tetrahedron": 4, "triangle": 3}[celltype]
    cre = re.compile(r'<%s index="(\d+)"' % celltype + "".join(r'\s+v%d="(\d+)"' % k for k in range(tag)))
    cells = np.zeros((nc, tag), dtype=np.int32)
    for m in cre.finditer(txt):
        cells[int(m.group(1))] = [int(m.group(2 + k)) for k in range(tag)]
    return coords, np.sort(cells, axis=1)


def read_mesh_function_xml(path):
    """dolfin-XML MeshFunction reader (SolverBase.py:229,236): returns (dim, values[size])."""
    txt = open(path, "r").read()
    m = re.search(r'<mesh_function type="\w+" dim="(\d+)" size="(\d+)"', txt)
    dim, size = int(m.group(1)), int(m.group(2))
    vals = np.zeros(size, dtype=np.int64)
    for m in re.finditer(r'<entity index="(\d+)" value="(\d+)"', txt):
        vals[int(m.group(1))] = int(m.group(2))
    return dim, vals


def facet_table(cells):
    """All facets with dolfin's numbering: local facet i is opposite local vertex i (of the
    sorted cell); global facet id = lexicographic rank of the sorted facet vertex tuple
    (verified against data/mesh_facet_region.xml).

    Returns (facets[nf, d], cell_facets[nc, d+1], count[nf]) ; count==1 <=> exterior facet."""
    nc, nl = cells.shape
    allf = np.stack([np.delete(cells, i, axis=1) for i in range(nl)], axis=1).reshape(-1, nl - 1)
    facets, inv, count = np.unique(allf, axis=0, return_inverse=True, return_counts=True)
    return facets, inv.reshape(nc, nl), count


def exterior_facets(cells):
    """Exterior facets as (verts[nbf, d], opposite_vertex[nbf], facet_id[nbf])."""
    nc, nl = cells.shape
    facets, cf, count = facet_table(cells)
    ext = count[cf] == 1                         # [nc, nl]
    ci, li = np.nonzero(ext)
    fid = cf[ci, li]
    order = np.argsort(fid, kind="stable")
    ci, li, fid = ci[order], li[order], fid[order]
    return facets[fid].astype(np.int32), cells[ci, li].astype(np.int32), fid


# --------------------------------------------------------------------------- geometry


def p1_geometry(coords, cells):
    """Per cell: |T| and the constant P1 gradients G[c, a, :] (UFC affine map, |det J|)."""
    d = cells.shape[1] - 1
    X = coords[cells]                            # [nc, d+1, d]
    J = np.transpose(X[:, 1:, :] - X[:, :1, :], (0, 2, 1))     # columns x_i - x_0
    detJ = np.linalg.det(J)
    vol = np.abs(detJ) / math.factorial(d)
    Jinv = np.linalg.inv(J)                      # rows = grad of lambda_1..d
    G = np.empty((cells.shape[0], d + 1, d))
    G[:, 1:, :] = Jinv
    G[:, 0, :] = -Jinv.sum(axis=1)
    return vol, G


def facet_measure(coords, fverts, opp=None):
    """|F| for each facet (edge length in 2D, triangle area in 3D) and, when `opp` is given,
    the outward unit normal (pointing away from the opposite vertex)."""
    X = coords[fverts]
    if fverts.shape[1] == 2:
        t = X[:, 1] - X[:, 0]
        meas = np.linalg.norm(t, axis=1)
        n = np.stack([t[:, 1], -t[:, 0]], axis=1)
    else:
        n = np.cross(X[:, 1] - X[:, 0], X[:, 2] - X[:, 0])
        meas = 0.5 * np.linalg.norm(n, axis=1)
    if opp is None:
        return meas
    n = n / np.linalg.norm(n, axis=1, keepdims=True)
    flip = np.einsum("ij,ij->i", n, coords[opp] - X[:, 0]) > 0
    n[flip] *= -1
    return meas, n


# --------------------------------------------------------------------------- pattern + assembly


def csr_pattern(cells, nverts, ncomp=1):
    """K2: topology-based CSR pattern (int64 row_ptr, int32 col_idx sorted, structural zeros kept).
    Vector spaces use interleaved dofs ncomp*v + c."""
    nl = cells.shape[1]
    I = np.repeat(cells, nl, axis=1).ravel()
    Jc = np.tile(cells, (1, nl)).ravel()
    P = sp.coo_matrix((np.ones(I.size, dtype=np.int8), (I, Jc)), shape=(nverts, nverts)).tocsr()
    P.sort_indices()
    if ncomp > 1:
        P = sp.kron(P, np.ones((ncomp, ncomp), dtype=np.int8), format="csr")
        P.sort_indices()
    return P.indptr.astype(np.int64), P.indices.astype(np.int32)


def _scatter(cells, Ke, nverts, ncomp=1):
    """Sum local matrices Ke[nc, nl*ncomp, nl*ncomp] into a CSR with the canonical pattern."""
    nl = cells.shape[1]
    dofs = (cells[:, :, None].astype(np.int64) * ncomp + np.arange(ncomp)).reshape(cells.shape[0], nl * ncomp)
    n = dofs.shape[1]
    I = np.repeat(dofs, n, axis=1).ravel()
    Jc = np.tile(dofs, (1, n)).ravel()
    A = sp.coo_matrix((Ke.reshape(-1), (I, Jc)), shape=(nverts * ncomp,) * 2).tocsr()
    A.sort_indices()
    return A


def _as_cell_tensor(k, nc, d):
    """Normalise a conductivity to per-cell tensors [nc, d, d] (scalar / dxd / per-cell)."""
    k = np.asarray(k, dtype=np.float64)
    if k.ndim == 0:
        return np.broadcast_to(np.eye(d) * k, (nc, d, d))
    if k.shape == (d, d):
        return np.broadcast_to(k, (nc, d, d))
    if k.shape == (nc,):
        return k[:, None, None] * np.eye(d)
    if k.shape == (nc, d, d):
        return k
    raise ValueError("conductivity shape %r" % (k.shape,))


def local_laplace(coords, cells, k=1.0):
    """K_e = |T| G k G^T   (ScalarTransportSolver.py:284-285, inner(k grad T, grad q) dx)."""
    vol, G = p1_geometry(coords, cells)
    kt = _as_cell_tensor(k, cells.shape[0], coords.shape[1])
    return vol[:, None, None] * np.einsum("cai,cij,cbj->cab", G, kt, G)


def local_mass(coords, cells, c=1.0):
    """M_e = c |T| /((d+1)(d+2)) (1 + delta_ab)   (ScalarTransportSolver.py:292 mass term)."""
    vol, _ = p1_geometry(coords, cells)
    nl = cells.shape[1]
    ref = (np.ones((nl, nl)) + np.eye(nl)) / (nl * (nl + 1))
    return (c * vol)[:, None, None] * ref


def local_advection(coords, cells, vel, c=1.0):
    """C_e[a,b] = c * int phi_a (v . grad phi_b)   (ScalarTransportSolver.py:311).
    vel: constant [d] or nodal [nverts, d] (P1 interpolant)."""
    vol, G = p1_geometry(coords, cells)
    nl = cells.shape[1]
    vel = np.asarray(vel, dtype=np.float64)
    if vel.ndim == 1:
        vg = np.einsum("j,cbj->cb", vel, G)                       # v . grad phi_b
        return (c * vol / nl)[:, None, None] * np.broadcast_to(vg[:, None, :], (cells.shape[0], nl, nl))
    ref = (np.ones((nl, nl)) + np.eye(nl)) / (nl * (nl + 1))      # int phi_a phi_c / |T|
    w = np.einsum("ae,cej->caj", ref, vel[cells])                 # sum_e M_ae v_e
    return (c * vol)[:, None, None] * np.einsum("caj,cbj->cab", w, G)


def local_advection_nodal(coords, cells, vel_nodes, c=1.0):
    """C_e[a,b] = c int (v_h . grad phi_b) phi_a with v_h the P1 interpolant of the nodal velocities
    (ScalarTransportSolver.py:130-139, 311): c |T|/((d+1)(d+2)) sum_k (1 + delta_ak) v_k . G_b."""
    vol, G = p1_geometry(coords, cells)
    nl = cells.shape[1]
    V = np.asarray(vel_nodes, dtype=np.float64).reshape(coords.shape[0], -1)[cells]      # [nc, nl, d]
    wsum = V.sum(axis=1)[:, None, :] + V                                                  # sum_k (1 + delta_ak) v_k
    return (c * vol / (nl * (nl + 1)))[:, None, None] * np.einsum("cai,cbi->cab", wsum, G)


def local_elasticity(coords, cells, mu, lmbda):
    """K_e[(a,i),(b,j)] = |T| ( mu (G_a.G_b d_ij + G_a[j] G_b[i]) + lambda G_a[i] G_b[j] )
    from inner(sigma(u), grad(v)) dx, sigma = 2 mu sym(grad u) + lambda div(u) I
    (LinearElasticitySolver.py:62-69, 215)."""
    vol, G = p1_geometry(coords, cells)
    nc, nl, d = G.shape
    gg = np.einsum("cai,cbi->cab", G, G)
    K = (mu * (np.einsum("cab,ij->caibj", gg, np.eye(d)) + np.einsum("caj,cbi->caibj", G, G))
         + lmbda * np.einsum("cai,cbj->caibj", G, G))
    return vol[:, None, None] * K.reshape(nc, nl * d, nl * d)


def assemble_matrix(cells, Ke, nverts, ncomp=1):
    return _scatter(cells, Ke, nverts, ncomp)


def assemble_source(coords, cells, S, ncomp=1, cell_mask=None):
    """b_a = int S phi_a dx: constant S (scalar or [ncomp]) -> |T|/(d+1) S; nodal S [nverts(,ncomp)]
    -> M_e S  (ScalarTransportSolver.py:213-226; LinearElasticitySolver.py:227-228)."""
    vol, _ = p1_geometry(coords, cells)
    if cell_mask is not None:
        vol = vol * cell_mask
    nl = cells.shape[1]
    nverts = coords.shape[0]
    S = np.asarray(S, dtype=np.float64)
    b = np.zeros((nverts, ncomp))
    if S.ndim == 0 or S.shape == (ncomp,):
        Sv = np.broadcast_to(S, (ncomp,))
        w = np.zeros(nverts)
        np.add.at(w, cells.ravel(), np.repeat(vol / nl, nl))
        b = w[:, None] * Sv[None, :]
    else:
        Sn = S.reshape(nverts, ncomp)
        ref = (np.ones((nl, nl)) + np.eye(nl)) / (nl * (nl + 1))
        be = vol[:, None, None] * np.einsum("ae,cek->cak", ref, Sn[cells])
        for k in range(ncomp):
            np.add.at(b[:, k], cells.ravel(), be[:, :, k].ravel())
    return b.reshape(-1)


def assemble_facet_load(coords, fverts, g, nverts, ncomp=1):
    """b_a += int_F g phi_a ds = |F|/d * g  for constant g (scalar, [ncomp] or per-facet [nf, ncomp])
    (ScalarTransportSolver.py:179-208; LinearElasticitySolver.py:180-196)."""
    meas = facet_measure(coords, fverts)
    d = fverts.shape[1]
    g = np.asarray(g, dtype=np.float64)
    gf = np.broadcast_to(g.reshape(-1, ncomp) if g.ndim else g, (fverts.shape[0], ncomp))
    b = np.zeros((nverts, ncomp))
    contrib = (meas / d)[:, None] * gf
    for k in range(ncomp):
        np.add.at(b[:, k], fverts.ravel(), np.repeat(contrib[:, k], d))
    return b.reshape(-1)


def local_facet_mass(coords, fverts, h=1.0):
    """h |F| /(d(d+1)) (1 + delta_ab): the HTC/Robin boundary matrix (ScalarTransportSolver.py:201-208)."""
    meas = facet_measure(coords, fverts)
    d = fverts.shape[1]
    ref = (np.ones((d, d)) + np.eye(d)) / (d * (d + 1))
    return (h * meas)[:, None, None] * ref


def boundary_area(coords, fverts):
    """assemble(Constant(1)*ds(id))  (LinearElasticitySolver.py:171)."""
    return float(facet_measure(coords, fverts).sum())


# --------------------------------------------------------------------------- Dirichlet + solve


def apply_dirichlet(A, b, dofs, values, symmetric):
    """DirichletBC.apply(A,b): zero row, unit diagonal, b=g (SolverBase.py:598-602, 608);
    symmetric=True restates assemble_system (SolverBase.py:644): also b -= A[:,bc] g, zero column.
    The sparsity pattern is kept (explicit zeros)."""
    A = A.tocsr().copy()
    b = np.array(b, dtype=np.float64, copy=True)
    n = A.shape[0]
    isbc = np.zeros(n, dtype=bool)
    isbc[dofs] = True
    g = np.zeros(n)
    g[dofs] = values
    rows = np.repeat(np.arange(n), np.diff(A.indptr))
    cols = A.indices
    if symmetric:
        colbc = isbc[cols] & ~isbc[rows]
        np.subtract.at(b, rows[colbc], A.data[colbc] * g[cols[colbc]])
        A.data[colbc] = 0.0
    A.data[isbc[rows]] = 0.0
    A.data[isbc[rows] & (rows == cols)] = 1.0
    b[isbc] = g[isbc]
    return A, b


def solve_direct(A, b):
    """The reference's scalar path: sparse LU (SolverBase.py:608-612 with dolfin defaults)."""
    return spla.spsolve(A.tocsc(), b)


def pcg_jacobi(A, b, x0=None, rtol=1e-12, atol=0.0, maxit=10000):
    """Jacobi-preconditioned CG, the exact recurrence the CUDA path runs (textbook PCG; convergence on
    the preconditioned recurrence residual ||M^-1 r||_2 <= max(rtol*||M^-1 b||_2, atol), PETSc's default
    KSP norm, which is insensitive to row scaling).  Returns (x, iterations, relres)."""
    A = A.tocsr()
    dinv = 1.0 / A.diagonal()
    x = np.zeros_like(b) if x0 is None else np.array(x0, dtype=np.float64, copy=True)
    r = b - A @ x
    bnorm = np.linalg.norm(dinv * b)
    tol = max(rtol * bnorm, atol)
    z = dinv * r
    p = z.copy()
    rz = r @ z
    rn = np.linalg.norm(z)
    it = 0
    while rn > tol and it < maxit:
        q = A @ p
        alpha = rz / (p @ q)
        x += alpha * p
        r -= alpha * q
        z = dinv * r
        rz_new = r @ z
        rn = np.linalg.norm(z)
        p = z + (rz_new / rz) * p
        rz = rz_new
        it += 1
    return x, it, rn / bnorm if bnorm > 0 else rn


def bicgstab_jacobi(A, b, x0=None, rtol=1e-12, atol=0.0, maxit=10000):
    """Right-Jacobi-preconditioned BiCGStab (van der Vorst), r0_hat = r0; convergence on ||M^-1 r||_2 as in
    pcg_jacobi.  Returns (x, it, relres)."""
    A = A.tocsr()
    dinv = 1.0 / A.diagonal()
    x = np.zeros_like(b) if x0 is None else np.array(x0, dtype=np.float64, copy=True)
    r = b - A @ x
    rhat = r.copy()
    bnorm = np.linalg.norm(dinv * b)
    tol = max(rtol * bnorm, atol)
    rho = alpha = omega = 1.0
    v = np.zeros_like(b)
    p = np.zeros_like(b)
    rn = np.linalg.norm(dinv * r)
    it = 0
    while rn > tol and it < maxit:
        rho_new = rhat @ r
        beta = (rho_new / rho) * (alpha / omega)
        p = r + beta * (p - omega * v)
        ph = dinv * p
        v = A @ ph
        alpha = rho_new / (rhat @ v)
        s = r - alpha * v
        sh = dinv * s
        t = A @ sh
        omega = (t @ s) / (t @ t)
        x += alpha * ph + omega * sh
        r = s - omega * t
        rho = rho_new
        rn = np.linalg.norm(dinv * r)
        it += 1
    return x, it, rn / bnorm if bnorm > 0 else rn


# --------------------------------------------------------------------------- thermal stress, von Mises, radiation


def thermal_load(coords, cells, beta, T, T_ref):
    """b[(a,i)] = int beta (T_h - T_ref) d(phi_a)/dx_i dx: the load of F -= inner(stress_t, grad(v))*dx with
    stress_t = E/(1-2nu) * tec * (T - T_ref) * I  (LinearElasticitySolver.py:78-85, 232-238).  T: number or
    nodal array (P1 interpolant, integrated exactly: |T| * mean of the vertex values)."""
    vol, G = p1_geometry(coords, cells)
    nv, d = coords.shape
    Tn = np.broadcast_to(np.asarray(T, dtype=np.float64), (nv,))
    dT = (Tn[cells] - T_ref).mean(axis=1)
    be = (beta * vol * dT)[:, None, None] * G                 # [nc, nl, d]
    b = np.zeros((nv, d))
    for i in range(d):
        np.add.at(b[:, i], cells.ravel(), be[:, :, i].ravel())
    return b.reshape(-1)


def von_mises_cells(coords, cells, u, mu, lmbda):
    """sqrt(3/2 s:s), s = sigma - (1/3) tr(sigma) I per cell for a P1 displacement (LinearElasticitySolver.py:71-73;
    the 1/3 holds in 2D as well, as written)."""
    vol, G = p1_geometry(coords, cells)
    d = coords.shape[1]
    U = np.asarray(u, dtype=np.float64).reshape(-1, d)[cells]          # [nc, nl, d]
    H = np.einsum("cai,cak->cik", U, G)                                # grad u
    eps = 0.5 * (H + np.transpose(H, (0, 2, 1)))
    sig = 2 * mu * eps + lmbda * np.trace(H, axis1=1, axis2=2)[:, None, None] * np.eye(d)
    s = sig - np.trace(sig, axis1=1, axis2=2)[:, None, None] / 3.0 * np.eye(d)
    return np.sqrt(1.5 * np.einsum("cij,cij->c", s, s))


def von_mises_projection(coords, cells, u, mu, lmbda):
    """project(von_Mises, FunctionSpace(mesh, 'P', 1)) (LinearElasticitySolver.py:75-76): M p = int vm phi_a."""
    nv = coords.shape[0]
    vol, _ = p1_geometry(coords, cells)
    vm = von_mises_cells(coords, cells, u, mu, lmbda)
    nl = cells.shape[1]
    b = np.zeros(nv)
    np.add.at(b, cells.ravel(), np.repeat(vm * vol / nl, nl))
    M = assemble_matrix(cells, local_mass(coords, cells, 1.0), nv)
    return spla.spsolve(M.tocsc(), b)


def _facet_monomial_integral(alpha, fd):
    """(1/|F|) int_F prod l_a^alpha_a ds on a simplex of dimension fd."""
    num = math.factorial(fd)
    for k in alpha:
        num *= math.factorial(k)
    return num / math.factorial(fd + sum(alpha))


def radiation_terms(coords, fverts, T, m, T_ambient):
    """Newton terms of the boundary flux m (Ta^4 - T^4) (ScalarTransportSolver.py:334-359, 361-374), T_h linear on each
    facet, integrated EXACTLY by expanding the polynomials in barycentric monomials:
    J[f,a,b] = int 4 m T_h^3 l_a l_b ds,  r[f,a] = int m (T_h^4 - Ta^4) l_a ds."""
    import itertools
    meas = facet_measure(coords, fverts)
    nf, d = fverts.shape
    fd = d - 1
    t = np.asarray(T, dtype=np.float64)[fverts]                  # [nf, d]
    J = np.zeros((nf, d, d))
    r = np.zeros((nf, d))

    def power_terms(p):
        """(sum_a t_a l_a)^p as a list of (alpha, coefficient[nf])."""
        out = []
        for combo in itertools.product(range(d), repeat=p):
            alpha = [0] * d
            c = np.ones(nf)
            for a in combo:
                alpha[a] += 1
                c = c * t[:, a]
            out.append((alpha, c))
        return out
    T3, T4 = power_terms(3), power_terms(4)
    for a in range(d):
        for alpha, c in T4:
            al = list(alpha); al[a] += 1
            r[:, a] += m * c * _facet_monomial_integral(al, fd)
        r[:, a] -= m * T_ambient ** 4 * _facet_monomial_integral([1 if k == a else 0 for k in range(d)], fd)
        for b in range(d):
            for alpha, c in T3:
                al = list(alpha); al[a] += 1; al[b] += 1
                J[:, a, b] += 4.0 * m * c * _facet_monomial_integral(al, fd)
    return meas[:, None, None] * J, meas[:, None] * r


def solve_radiation_newton(coords, cells, k, dirichlet, rad_fverts, m, T_ambient, T0, neumann=(), source=None,
                           rtol=1e-12, maxit=50):
    """Steady heat conduction with the radiation boundary term on rad_fverts, solved by Newton's method as the
    reference's NonlinearVariationalSolver does (SolverBase.py:615-626): F(T) = K T - b + R(T), J = K + dR/dT,
    Dirichlet rows T = g.  Direct linear solves.  -> (T, newton iterations)."""
    nv = coords.shape[0]
    K = assemble_matrix(cells, local_laplace(coords, cells, k), nv)
    b = np.zeros(nv)
    if source is not None:
        b += assemble_source(coords, cells, source)
    for fv, g in neumann:
        b += assemble_facet_load(coords, fv, g, nv)
    dofs = np.concatenate([np.asarray(d[0]) for d in dirichlet])
    vals = np.concatenate([np.broadcast_to(np.asarray(d[1], dtype=np.float64), np.asarray(d[0]).shape) for d in dirichlet])
    T = np.full(nv, float(T0)) if np.isscalar(T0) else np.array(T0, dtype=np.float64)
    T[dofs] = vals
    r0 = None
    for it in range(maxit):
        Jf, rf = radiation_terms(coords, rad_fverts, T, m, T_ambient)
        A = (K + _scatter(rad_fverts, Jf, nv)).tocsr()
        res = K @ T - b
        np.add.at(res, rad_fverts.ravel(), rf.ravel())
        A, rhs = apply_dirichlet(A, -res, dofs, np.zeros(dofs.size), symmetric=True)
        nrm = np.linalg.norm(rhs)
        r0 = nrm if r0 is None else r0
        if nrm <= rtol * max(r0, 1e-300):
            return T, it
        T = T + solve_direct(A, rhs)
    return T, maxit


def nonlinear_k_terms(coords, cells, T, kfun, dkfun):
    """Residual and Jacobian of int k(T) grad T . grad q with k_h the P1 interpolant of the nodal values k(T_a)
    (ScalarTransportSolver.py:228-233, 284-285, 352-353): R_a = sum |T| kbar G_a.grad T,
    J_ab = sum |T| (kbar G_a.G_b + k'(T_b)/(d+1) G_a.grad T)."""
    vol, G = p1_geometry(coords, cells)
    nv = coords.shape[0]
    nl = cells.shape[1]
    Tc = np.asarray(T, dtype=np.float64)[cells]
    kbar = kfun(Tc).mean(axis=1)
    gT = np.einsum("ca,cai->ci", Tc, G)
    flux = np.einsum("cai,ci->ca", G, gT)
    Re = (vol * kbar)[:, None] * flux
    Je = vol[:, None, None] * (kbar[:, None, None] * np.einsum("cai,cbi->cab", G, G)
                                + flux[:, :, None] * (dkfun(Tc) / nl)[:, None, :])
    R = np.zeros(nv)
    np.add.at(R, cells.ravel(), Re.ravel())
    return _scatter(cells, Je, nv), R


def solve_nonlinear_k_newton(coords, cells, kfun, dkfun, dirichlet, T0, neumann=(), rtol=1e-12, maxit=50):
    """Steady conduction with temperature-dependent conductivity by Newton's method (the reference's
    NonlinearVariationalSolver path, SolverBase.py:615-626).  -> (T, iterations)."""
    nv = coords.shape[0]
    b = np.zeros(nv)
    for fv, g in neumann:
        b += assemble_facet_load(coords, fv, g, nv)
    dofs = np.concatenate([np.asarray(d[0]) for d in dirichlet])
    vals = np.concatenate([np.broadcast_to(np.asarray(d[1], dtype=np.float64), np.asarray(d[0]).shape) for d in dirichlet])
    T = np.full(nv, float(T0))
    T[dofs] = vals
    r0 = None
    for it in range(maxit):
        J, R = nonlinear_k_terms(coords, cells, T, kfun, dkfun)
        A, rhs = apply_dirichlet(J.tocsr(), b - R, dofs, np.zeros(dofs.size), symmetric=False)
        nrm = np.linalg.norm(rhs)
        r0 = nrm if r0 is None else r0
        if nrm <= rtol * max(r0, 1e-300):
            return T, it
        T = T + solve_direct(A, rhs)
    return T, maxit


# --------------------------------------------------------------------------- SUPG (ScalarTransportSolver.py:252-274)


def circumradius(coords, cells):
    """Circumradius per cell from the circumcentre (the point equidistant from all vertices): 2 (x_i - x_0).c =
    |x_i|^2 - |x_0|^2 -- a different route from the edge-length formulas of the CUDA kernel."""
    X = coords[cells]
    A = 2.0 * (X[:, 1:, :] - X[:, :1, :])
    rhs = np.sum(X[:, 1:, :] ** 2, axis=2) - np.sum(X[:, :1, :] ** 2, axis=2)
    ctr = np.linalg.solve(A, rhs[..., None])[..., 0]
    return np.linalg.norm(ctr - X[:, 0, :], axis=1)


def supg_weights(coords, cells, vel, Pe):
    """s[c, a] = tau_c (v . G_a): the extra part of the SUPG test function Tq = q + tau v.grad(q) on P1,
    tau = 0.5 h / (4/(Pe h) + 2 |v|), h = 2 * Circumradius (:256-266)."""
    vol, G = p1_geometry(coords, cells)
    vel = np.asarray(vel, dtype=np.float64)
    h = 2.0 * circumradius(coords, cells)
    tau = 0.5 * h / (4.0 / (Pe * h) + 2.0 * np.linalg.norm(vel))
    return tau[:, None] * np.einsum("cai,i->ca", G, vel)


def local_supg(coords, cells, vel, Pe, mass=0.0, adv=0.0):
    """Extra local matrix  s_a int (mass phi_b + adv v.grad phi_b)."""
    vol, G = p1_geometry(coords, cells)
    nl = cells.shape[1]
    s = supg_weights(coords, cells, vel, Pe)
    col = vol[:, None] * (mass / nl + adv * np.einsum("cbi,i->cb", G, np.asarray(vel, dtype=np.float64)))
    return s[:, :, None] * col[:, None, :]


def supg_source(coords, cells, S, vel, Pe):
    vol, _ = p1_geometry(coords, cells)
    b = np.zeros(coords.shape[0])
    np.add.at(b, cells.ravel(), (S * vol[:, None] * supg_weights(coords, cells, vel, Pe)).ravel())
    return b


def supg_facet_terms(coords, fverts, opp, vel, Pe, g=0.0, h=0.0):
    """Extra facet terms on the cells (fverts, opp): b_a = g s_a |F| and the matrix h s_a |F|/d over the facet's nodes."""
    nv = coords.shape[0]
    cells = np.hstack([fverts, opp[:, None]]).astype(np.int64)
    s = supg_weights(coords, cells, vel, Pe)
    meas = facet_measure(coords, fverts)
    d = fverts.shape[1]
    b = np.zeros(nv)
    np.add.at(b, cells.ravel(), (g * s * meas[:, None]).ravel())
    rows = np.repeat(cells, d, axis=1).ravel()
    cols = np.tile(fverts, (1, d + 1)).ravel()
    vals = np.repeat(h * s * meas[:, None] / d, d, axis=1).ravel()
    A = sp.coo_matrix((vals, (rows, cols)), shape=(nv, nv)).tocsr()
    return A, b


# --------------------------------------------------------------------------- whole-problem restatements


def lame(E, nu):
    """mu, lambda as LinearElasticitySolver.py:210-213."""
    return E / (2.0 * (1.0 + nu)), E * nu / ((1.0 + nu) * (1.0 - 2.0 * nu))


def heat_system(coords, cells, k, dirichlet, source=None, neumann=(), robin=(), velocity=None,
                capacity=1.0, symmetric=True):
    """Steady scalar transport system (ScalarTransportSolver.generate_form, steady branch):
    A = K(k) + c C(v) + sum h M_F ;  b = int S q + sum int g q ds + sum h Ta int q ds.
    dirichlet: list of (vertex_ids, value-or-array); neumann: list of (fverts, g);
    robin: list of (fverts, h, Ta)."""
    nv = coords.shape[0]
    Ke = local_laplace(coords, cells, k)
    if velocity is not None:
        Ke = Ke + local_advection(coords, cells, velocity, capacity)
    A = assemble_matrix(cells, Ke, nv)
    b = np.zeros(nv)
    if source is not None:
        b += assemble_source(coords, cells, source)
    for fverts, g in neumann:
        b += assemble_facet_load(coords, fverts, g, nv)
    for fverts, h, Ta in robin:
        A = A + _scatter(fverts, local_facet_mass(coords, fverts, h), nv)
        b += assemble_facet_load(coords, fverts, h * Ta, nv)
    rp, ci = csr_pattern(cells, nv)
    A = conform(A, rp, ci)
    dofs = np.concatenate([np.asarray(d[0]) for d in dirichlet]) if dirichlet else np.zeros(0, dtype=np.int64)
    vals = np.concatenate([np.broadcast_to(np.asarray(d[1], dtype=np.float64), np.asarray(d[0]).shape)
                           for d in dirichlet]) if dirichlet else np.zeros(0)
    return apply_dirichlet(A, b, dofs, vals, symmetric)


def conform(A, row_ptr, col_idx):
    """Re-express A on the canonical pattern (adds explicit zeros where A has no entry)."""
    n = A.shape[0]
    P = sp.csr_matrix((np.zeros(col_idx.size), col_idx.astype(np.int64), row_ptr.astype(np.int64)), shape=(n, n))
    A = A.tocsr()
    A.sort_indices()
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(A.indptr))
    key_p = np.repeat(np.arange(n, dtype=np.int64), np.diff(P.indptr)) * n + P.indices
    key_a = rows * n + A.indices
    pos = np.searchsorted(key_p, key_a)
    assert np.all(key_p[pos] == key_a), "matrix has entries outside the canonical pattern"
    P.data[pos] = A.data
    return P


def relative_l2(x, ref):
    return float(np.linalg.norm(np.asarray(x) - np.asarray(ref)) / np.linalg.norm(ref))
