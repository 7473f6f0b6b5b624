"""CPU tests pinning the degree-2 oracle (oracle/fem_oracle_p2.py): partition of unity, exactness for quadratic
fields (patch tests), consistency with the P1 oracle, rigid-body modes, known reference-element numbers."""
import numpy as np
import pytest

from oracle import fem_oracle as fo
from oracle import fem_oracle_p2 as p2


def jittered(dim, seed=0):
    c0, t = (fo.unit_square_mesh(4, 3) if dim == 2 else fo.unit_cube_mesh(3, 2, 2))
    rng = np.random.default_rng(seed)
    c = c0 + 0.15 / 4 * (rng.random(c0.shape) * 2 - 1)
    bnd = np.any((c0 == 0) | (c0 == 1), axis=1)
    c[bnd] = c0[bnd]
    return c, t


def test_reference_tensors_known_values():
    R, M, S, F = p2.reference_tensors(3)
    assert np.allclose(F, [-1 / 20] * 4 + [1 / 5] * 6)                      # P2 vertex functions integrate to -|T|/20
    assert abs(M.sum() - 1.0) < 1e-14 and abs(M[0, 0] - 1 / 70) < 1e-15 and abs(M[0, 1] - 1 / 420) < 1e-15
    # sum_i dphi_i/dl_c = 4 sum(l) - 1 = 3 for every c, so the physical gradient of the sum (contracted with
    # sum_c G_c = 0) vanishes: the c-dependence of the summed tensors must be trivial
    Rs = R.sum(axis=0)
    assert np.allclose(Rs, Rs[:, :1, :]) and np.allclose(S.sum(axis=1), S.sum(axis=1)[:, :1])
    R2, M2, S2, F2 = p2.reference_tensors(2)
    assert np.allclose(F2, [0, 0, 0, 1 / 3, 1 / 3, 1 / 3])
    assert np.allclose(p2.reference_tensors(1)[1] * 30, [[4, -1, 2], [-1, 4, 2], [2, 2, 16]])


@pytest.mark.parametrize("dim", [2, 3])
def test_quadratic_fields_are_reproduced(dim):
    c, t = jittered(dim)
    cn, xc, edges = p2.p2_dofmap(c, t)
    nn = xc.shape[0]
    assert nn == c.shape[0] + edges.shape[0] and cn.shape[1] == (dim + 1) * (dim + 2) // 2
    rng = np.random.default_rng(1)
    Q = rng.random((dim, dim)); Q = Q + Q.T
    lin = rng.random(dim)
    u = np.einsum("ni,ij,nj->n", xc, Q, xc) + xc @ lin + 1.0
    K = p2.assemble_matrix(cn, p2.local_laplace(c, t, 1.0), nn)
    b = p2.assemble_source(c, t, cn, nn, -2 * np.trace(Q))
    fv, _, _ = fo.exterior_facets(t)
    bd = np.unique(p2.facet_nodes(fv, edges, c.shape[0]))
    A, bb = fo.apply_dirichlet(K, b, bd, u[bd], True)
    assert fo.relative_l2(fo.solve_direct(A, bb), u) < 1e-13
    # mass matrix integrates products exactly: 1^T M u = int u (quadratic), checked against the P2 load vector
    Mm = p2.assemble_matrix(cn, p2.local_mass(c, t, 1.0), nn)
    assert abs(Mm.sum() - 1.0) < 1e-13
    assert abs(np.ones(nn) @ (Mm @ u) - p2.assemble_source(c, t, cn, nn, 1.0) @ u) < 1e-12
    # advection: C 1 = 0 and 1^T C u = int v.grad u = boundary flux of (v u) for constant v
    v = rng.random(dim)
    C = p2.assemble_matrix(cn, p2.local_advection(c, t, v, 1.0), nn)
    assert np.abs(C @ np.ones(nn)).max() < 1e-13
    fn = p2.facet_nodes(fv, edges, c.shape[0])
    meas, nrm = fo.facet_measure(c, fv, fo.exterior_facets(t)[1])
    flux = sum((p2.assemble_facet_load(c, fv[i:i + 1], fn[i:i + 1], float(nrm[i] @ v), nn) @ u) for i in range(fv.shape[0]))
    assert abs(np.ones(nn) @ (C @ u) - flux) < 1e-11


@pytest.mark.parametrize("dim", [2, 3])
def test_elasticity_rigid_modes_and_linear_patch(dim):
    c, t = jittered(dim, seed=2)
    cn, xc, edges = p2.p2_dofmap(c, t)
    nn = xc.shape[0]
    mu, lam = 3.0, 5.0
    A = p2.assemble_matrix(cn, p2.local_elasticity(c, t, mu, lam), nn, dim)
    assert abs(A - A.T).max() < 1e-12 * abs(A).max()
    assert np.abs(A @ np.tile(np.eye(dim)[0], nn)).max() < 1e-12 * abs(A).max()
    rot = np.zeros((nn, dim)); rot[:, 0] = -xc[:, 1]; rot[:, 1] = xc[:, 0]
    assert np.abs(A @ rot.ravel()).max() < 1e-12 * abs(A).max()
    # P2 contains P1: on linear fields the P2 stiffness energy equals the P1 one
    G = np.random.default_rng(3).random((dim, dim)) * 1e-2
    u2 = (xc @ G.T).ravel()
    u1 = (c @ G.T).ravel()
    A1 = fo.assemble_matrix(t, fo.local_elasticity(c, t, mu, lam), c.shape[0], dim)
    assert abs(u2 @ (A @ u2) - u1 @ (A1 @ u1)) < 1e-12 * abs(u1 @ (A1 @ u1))


def test_facet_terms_and_dofmap_conventions():
    c, t = fo.unit_cube_mesh(2, 2, 2)
    cn, xc, edges = p2.p2_dofmap(c, t)
    nv = c.shape[0]
    assert np.all(np.diff(edges, axis=1) > 0) and np.all((edges[1:] > edges[:-1]).any(axis=1))      # sorted pairs, lexicographic order
    # local node order: vertices then UFC edges (2,3)(1,3)(1,2)(0,3)(0,2)(0,1)
    k = 7
    for loc, (a, b) in enumerate(p2.EDGES[3]):
        e = edges[cn[k, 4 + loc] - nv]
        assert tuple(e) == (t[k, a], t[k, b])
        assert np.allclose(xc[cn[k, 4 + loc]], 0.5 * (c[t[k, a]] + c[t[k, b]]))
    fv, opp, _ = fo.exterior_facets(t)
    fn = p2.facet_nodes(fv, edges, nv)
    assert fn.shape == (fv.shape[0], 6)
    load = p2.assemble_facet_load(c, fv, fn, 1.0, xc.shape[0])
    assert abs(load.sum() - 6.0) < 1e-13                                    # surface area of the unit cube
    assert np.abs(load[:nv]).max() < 1e-15                                  # P2 vertex functions integrate to zero on a facet
    Mf = fo._scatter(fn, p2.local_facet_mass(c, fv, 1.0), xc.shape[0])
    assert abs(Mf.sum() - 6.0) < 1e-13


def test_c_oracle_degree_2_heat_matches_numpy_oracle():
    """oracle/fem_oracle_c.c fo_assemble_heat_p2 (the CPU figure of the bench's `p2` block) against the numpy degree-2 oracle on a 5^3 cube:
    matrix and load vector to rounding, and the solved problem reproduces the exact quadratic profile at every node."""
    from oracle import c_oracle as co
    from oracle import fem_oracle as fo
    N = 5
    h = co.HeatCubeP2(N)
    c, t = fo.unit_cube_mesh(N, N, N)
    cn, xn, _ = p2.p2_dofmap(c, t)
    assert np.array_equal(cn, h.cell_nodes) and np.array_equal(xn, h.node_coords)
    A = fo.conform(p2.assemble_matrix(cn, p2.local_laplace(c, t, h.k), xn.shape[0]), h.rp, h.ci)
    h.lib.fo_zero(co._p(h.vals), h.vals.size)
    h.lib.fo_zero(co._p(h.b), h.b.size)
    h.lib.fo_assemble_heat_p2(h.cell_nodes.shape[0], co._p(h.cell_nodes), co._p(h.coords), co._p(h.R), co._p(h.F), h.k, h.S,
                              co._p(h.rp), co._p(h.ci), co._p(h.vals), co._p(h.b))
    assert np.abs(h.vals - A.data).max() <= 1e-14 * np.abs(A.data).max()
    assert np.abs(h.b - p2.assemble_source(c, t, cn, xn.shape[0], h.S)).max() <= 1e-13 * np.abs(h.b).max()
    r = h.step()
    assert r["rel_l2_vs_exact"] < 1e-10 and r["iterations"] > 10
