"""CPU oracle, degree-2 Lagrange (P2) on affine simplices.  TEST INFRASTRUCTURE ONLY (see fem_oracle.py;
the same "parity unpinned at the dolfin boundary" caveat applies: pinned by patch tests with quadratic
fields, manufactured solutions and P1/P2 consistency, tests/test_oracle_p2.py).

Reference use of P2: examples/test_linear_elasticity.py:105-106 (`VectorFunctionSpace(mesh, "Lagrange", 2)`),
ScalarTransportSolver.py:136 (P(k+1) velocity space).

Conventions (ours; dolfin's dof numbering is not reproducible without dolfin):
* nodes = vertices 0..nv-1, then edges nv + e, e = lexicographic rank of the sorted vertex pair;
* local nodes of a cell (vertices sorted ascending): vertices 0..d, then edges in UFC order:
  triangle (1,2) (0,2) (0,1); tetrahedron (2,3) (1,3) (1,2) (0,3) (0,2) (0,1);
* basis in barycentric coordinates: vertex a: l_a (2 l_a - 1); edge (a,b): 4 l_a l_b.

Everything is written as a contraction of per-cell geometry (volume, gradients of the barycentric
coordinates) with reference tensors integrated EXACTLY with the monomial formula
    int_T l^alpha dx = |T| d! alpha! / (d + |alpha|)!
-- an independent route from the CUDA path, which builds its tables by collapsed Gauss quadrature.
"""
from __future__ import annotations

import math
from functools import lru_cache

import numpy as np

from . import fem_oracle as fo

EDGES = {1: ((0, 1),), 2: ((1, 2), (0, 2), (0, 1)), 3: ((2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1))}


# --------------------------------------------------------------------------- polynomials in barycentric coordinates
def _poly_mul(p, q):
    out = {}
    for a, ca in p.items():
        for b, cb in q.items():
            k = tuple(x + y for x, y in zip(a, b))
            out[k] = out.get(k, 0.0) + ca * cb
    return out


def _poly_int(p, d):
    """(1/|T|) int_T p(lambda) dx."""
    s = 0.0
    for a, c in p.items():
        num = math.factorial(d)
        for k in a:
            num *= math.factorial(k)
        s += c * num / math.factorial(d + sum(a))
    return s


def _basis(d):
    """P2 basis functions and their derivatives w.r.t. each barycentric coordinate, as polynomial dicts."""
    n = d + 1
    def mono(*idx):
        e = [0] * n
        for i in idx:
            e[i] += 1
        return tuple(e)
    phi, dphi = [], []
    for a in range(n):
        phi.append({mono(a, a): 2.0, mono(a): -1.0})
        dphi.append([({mono(a): 4.0, mono(): -1.0} if c == a else {}) for c in range(n)])
    for (a, b) in EDGES[d]:
        phi.append({mono(a, b): 4.0})
        dphi.append([({mono(b): 4.0} if c == a else ({mono(a): 4.0} if c == b else {})) for c in range(n)])
    return phi, dphi


@lru_cache(maxsize=None)
def reference_tensors(d):
    """R[i,j,c,e] = (1/|T|) int dphi_i/dl_c dphi_j/dl_e ; M[i,j] = (1/|T|) int phi_i phi_j ;
    S[i,j,e] = (1/|T|) int phi_i dphi_j/dl_e ; F[i] = (1/|T|) int phi_i."""
    phi, dphi = _basis(d)
    nn, n = len(phi), d + 1
    R = np.zeros((nn, nn, n, n))
    M = np.zeros((nn, nn))
    S = np.zeros((nn, nn, n))
    F = np.zeros(nn)
    for i in range(nn):
        F[i] = _poly_int(phi[i], d)
        for j in range(nn):
            M[i, j] = _poly_int(_poly_mul(phi[i], phi[j]), d)
            for e in range(n):
                if dphi[j][e]:
                    S[i, j, e] = _poly_int(_poly_mul(phi[i], dphi[j][e]), d)
                for c in range(n):
                    if dphi[i][c] and dphi[j][e]:
                        R[i, j, c, e] = _poly_int(_poly_mul(dphi[i][c], dphi[j][e]), d)
    return R, M, S, F


# --------------------------------------------------------------------------- dof maps
def edge_table(cells):
    """Unique edges (sorted vertex pairs, lexicographic numbering) and the cell->edge map in UFC local order."""
    d = cells.shape[1] - 1
    pairs = np.stack([cells[:, list(e)] for e in EDGES[d]], axis=1)          # [nc, ne_loc, 2] (already sorted: cells sorted)
    edges, inv = np.unique(pairs.reshape(-1, 2), axis=0, return_inverse=True)
    return edges, inv.reshape(cells.shape[0], len(EDGES[d]))


def p2_dofmap(coords, cells):
    """-> (cell_nodes[nc, nn], node_coords[nnodes, d], edges)."""
    edges, ce = edge_table(cells)
    nv = coords.shape[0]
    cell_nodes = np.hstack([cells.astype(np.int64), nv + ce]).astype(np.int32)
    node_coords = np.vstack([coords, 0.5 * (coords[edges[:, 0]] + coords[edges[:, 1]])])
    return cell_nodes, node_coords, edges


def facet_nodes(fverts, edges, nv):
    """P2 nodes of each facet: its vertices then its edges (facet-local UFC order)."""
    d = fverts.shape[1]                       # vertices per facet
    key = {tuple(e): i for i, e in enumerate(map(tuple, edges))}
    loc = EDGES[d - 1]
    fe = np.array([[key[(min(f[a], f[b]), max(f[a], f[b]))] for (a, b) in loc] for f in fverts], dtype=np.int64).reshape(len(fverts), -1)
    return np.hstack([fverts.astype(np.int64), nv + fe])


def csr_pattern(cell_nodes, nnodes, ncomp=1):
    return fo.csr_pattern(cell_nodes, nnodes, ncomp)


# --------------------------------------------------------------------------- local matrices
def local_laplace(coords, cells, k=1.0):
    vol, G = fo.p1_geometry(coords, cells)
    d = coords.shape[1]
    R, _, _, _ = reference_tensors(d)
    kt = fo._as_cell_tensor(k, cells.shape[0], d)
    gkg = np.einsum("cai,cij,cbj->cab", G, kt, G)                 # G_c . K G_e
    return vol[:, None, None] * np.einsum("ijab,cab->cij", R, gkg)


def local_mass(coords, cells, c=1.0):
    vol, _ = fo.p1_geometry(coords, cells)
    _, M, _, _ = reference_tensors(coords.shape[1])
    return (c * vol)[:, None, None] * M


def local_advection(coords, cells, vel, c=1.0):
    vol, G = fo.p1_geometry(coords, cells)
    _, _, S, _ = reference_tensors(coords.shape[1])
    vg = np.einsum("j,cej->ce", np.asarray(vel, dtype=np.float64), G)          # v . grad l_e
    return (c * vol)[:, None, None] * np.einsum("ije,ce->cij", S, vg)


def local_elasticity(coords, cells, mu, lmbda):
    vol, G = fo.p1_geometry(coords, cells)
    d = coords.shape[1]
    R, _, _, _ = reference_tensors(d)
    gg = np.einsum("cai,cbi->cab", G, G)
    # T[c, a, e, i, j] = mu (G_a.G_e d_ij + G_a[j] G_e[i]) + lambda G_a[i] G_e[j]   (vertex-gradient pair a, e)
    T = (mu * (np.einsum("cae,ij->caeij", gg, np.eye(d)) + np.einsum("caj,cei->caeij", G, G))
         + lmbda * np.einsum("cai,cej->caeij", G, G))
    K = np.einsum("pqae,caeij->cpiqj", R, T)                       # [(node p, comp i), (node q, comp j)]
    nn = R.shape[0]
    return vol[:, None, None] * K.reshape(cells.shape[0], nn * d, nn * d)


def assemble_matrix(cell_nodes, Ke, nnodes, ncomp=1):
    return fo._scatter(cell_nodes, Ke, nnodes, ncomp)


def assemble_source(coords, cells, cell_nodes, nnodes, S, ncomp=1):
    """b_i = int S phi_i for constant S (scalar or [ncomp])."""
    vol, _ = fo.p1_geometry(coords, cells)
    _, _, _, F = reference_tensors(coords.shape[1])
    Sv = np.broadcast_to(np.asarray(S, dtype=np.float64), (ncomp,))
    w = np.zeros(nnodes)
    np.add.at(w, cell_nodes.ravel(), (vol[:, None] * F[None, :]).ravel())
    return (w[:, None] * Sv[None, :]).reshape(-1)


def assemble_facet_load(coords, fverts, fnodes, g, nnodes, ncomp=1):
    """b_i += int_F g phi_i ds, constant g (scalar, [ncomp] or per facet [nf, ncomp])."""
    meas = fo.facet_measure(coords, fverts)
    _, _, _, F = reference_tensors(fverts.shape[1] - 1)
    g = np.asarray(g, dtype=np.float64)
    gf = np.broadcast_to(g.reshape(-1, ncomp) if g.ndim else g, (fverts.shape[0], ncomp))
    b = np.zeros((nnodes, ncomp))
    for k in range(ncomp):
        np.add.at(b[:, k], fnodes.ravel(), (meas[:, None] * F[None, :] * gf[:, k:k + 1]).ravel())
    return b.reshape(-1)


def local_facet_mass(coords, fverts, h=1.0):
    meas = fo.facet_measure(coords, fverts)
    _, M, _, _ = reference_tensors(fverts.shape[1] - 1)           # on an edge: [[4,-1,2],[-1,4,2],[2,2,16]]/30
    return (h * meas)[:, None, None] * M


def interpolate(fn, node_coords):
    return fn(node_coords)


# --------------------------------------------------------------------------- thermal stress, von Mises
def thermal_load(coords, cells, cell_nodes, nnodes, beta, T, T_ref):
    """b[(a,i)] = int beta (T_h - T_ref) d(phi_a)/dx_i dx with T_h the P2 interpolant of the nodal values (or a
    constant): |T| sum_j dT_j sum_e S[j,a,e] G_e[i]  (LinearElasticitySolver.py:78-85, 232-238)."""
    vol, G = fo.p1_geometry(coords, cells)
    d = coords.shape[1]
    _, _, S, _ = reference_tensors(d)
    Tn = np.broadcast_to(np.asarray(T, dtype=np.float64), (nnodes,))
    dT = Tn[cell_nodes] - T_ref                                    # [nc, nn]
    be = beta * vol[:, None, None] * np.einsum("cj,jae,cei->cai", dT, S, G)
    b = np.zeros((nnodes, d))
    for i in range(d):
        np.add.at(b[:, i], cell_nodes.ravel(), be[:, :, i].ravel())
    return b.reshape(-1)


def _collapsed_rule(d, n):
    """Collapsed Gauss-Legendre rule on the reference simplex: barycentric points [np, d+1], weights summing to 1."""
    x, w = np.polynomial.legendre.leggauss(n)
    x, w = 0.5 * (x + 1.0), 0.5 * w
    pts, wts = [], []
    if d == 2:
        for i in range(n):
            for j in range(n):
                X, Y = x[i], x[j] * (1 - x[i])
                pts.append((1 - X - Y, X, Y)); wts.append(2 * w[i] * w[j] * (1 - x[i]))
    else:
        for i in range(n):
            for j in range(n):
                for k in range(n):
                    X, Y, Z = x[i], x[j] * (1 - x[i]), x[k] * (1 - x[i]) * (1 - x[j])
                    pts.append((1 - X - Y - Z, X, Y, Z)); wts.append(6 * w[i] * w[j] * w[k] * (1 - x[i]) ** 2 * (1 - x[j]))
    return np.array(pts), np.array(wts)


def von_mises_load(coords, cells, cell_nodes, u, mu, lmbda, n=4):
    """b_a = int vm(u_h) lambda_a dx over the vertices, u_h degree 2; the integrand is not polynomial, so the value
    depends on the rule: a collapsed Gauss rule with n = 4 points per axis (degree 5, what UFL would estimate)."""
    vol, G = fo.p1_geometry(coords, cells)
    d = coords.shape[1]
    _, dphi = _basis(d)
    nn = len(dphi)
    U = np.asarray(u, dtype=np.float64).reshape(-1, d)[cell_nodes]      # [nc, nn, d]
    pts, wts = _collapsed_rule(d, n)
    b = np.zeros(coords.shape[0])
    for l, w in zip(pts, wts):
        D = np.zeros((nn, d + 1))
        for j in range(nn):
            for e in range(d + 1):
                D[j, e] = sum(c * np.prod([l[a] ** k for a, k in enumerate(al)]) for al, c in dphi[j][e].items()) if dphi[j][e] else 0.0
        gphi = np.einsum("je,cek->cjk", D, G)                            # grad phi_j per cell
        H = np.einsum("cji,cjk->cik", U, gphi)
        eps = 0.5 * (H + np.transpose(H, (0, 2, 1)))
        sig = 2 * mu * eps + lmbda * np.trace(H, axis1=1, axis2=2)[:, None, None] * np.eye(d)
        s = sig - np.trace(sig, axis1=1, axis2=2)[:, None, None] / 3.0 * np.eye(d)
        vm = np.sqrt(1.5 * np.einsum("cij,cij->c", s, s))
        for a in range(d + 1):
            np.add.at(b, cells[:, a], w * vol * vm * l[a])
    return b
