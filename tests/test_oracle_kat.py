"""CPU tests: pin the oracle (numpy and C restatements) against the reference's shipped fixtures and the
analytic known-answer tests of SURVEY 8c.  The reference's own tests assert no numbers (parity unpinned
at the dolfin boundary), so these pins are what "correct" means for the GPU parity tests."""
import os

import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import fem_oracle as fo


@pytest.fixture(scope="module")
def fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "fixture_mesh.npz"))
    e = np.load(os.path.join(golden_dir, "fixture_expected.npz"))
    return g, e


def test_fixture_mesh_shape_and_volume(fixture):
    g, _ = fixture
    c, t = g["coords"], g["cells"]
    assert c.shape == (1069, 3) and t.shape == (4355, 4)
    vol, _ = fo.p1_geometry(c, t)
    assert vol.min() > 0 and abs(vol.sum() - 1000.0) < 1e-9            # box 10 x 5 x 20
    assert np.all(np.diff(t, axis=1) > 0)                               # cells sorted (mesh.order())
    assert np.all(g["cell_tags"] == 3)


def test_facet_numbering_rule_against_shipped_markers(fixture):
    """Global facet id = lexicographic rank of the sorted vertex tuple: with this rule every tagged facet of
    mesh_facet_region.xml is exterior and lies on z=0 (tag 1) / z=20 (tag 2)."""
    g, e = fixture
    c, t, tags = g["coords"], g["cells"], g["facet_tags"]
    facets, cf, count = fo.facet_table(t)
    assert facets.shape[0] == tags.size == 9410
    assert (count == 1).sum() == 1400 == int(e["n_exterior_facets"])
    assert np.bincount(tags).tolist() == [9210, 100, 100]
    for tag, z in ((1, 0.0), (2, 20.0)):
        f = facets[tags == tag]
        assert np.all(count[tags == tag] == 1)
        assert np.all(c[f.ravel(), 2] == z)
    # a first-appearance numbering would put most tags on interior facets: the rule is not vacuous
    assert (count[tags > 0] == 1).all() and (count == 2).sum() > 0


def test_fixture_json_case_known_answer(fixture):
    """data/TestHeatTransfer.json: Dirichlet 350 on tag 1, 300 on tag 2, k=20 -> T = 350 - 2.5 z exactly."""
    g, e = fixture
    c, t, tags = g["coords"], g["cells"], g["facet_tags"]
    facets, _, _ = fo.facet_table(t)
    d1, d2 = np.unique(facets[tags == 1]), np.unique(facets[tags == 2])
    assert d1.size == 66 and d2.size == 66
    assert np.array_equal(d1, e["dofs_tag1"]) and np.array_equal(d2, e["dofs_tag2"])
    rp, ci = fo.csr_pattern(t, c.shape[0])
    assert ci.size == 13315 and np.array_equal(rp, e["row_ptr"]) and np.array_equal(ci, e["col_idx"])
    for symmetric in (True, False):
        A, b = fo.heat_system(c, t, 20.0, [(d1, 350.0), (d2, 300.0)], symmetric=symmetric)
        x = fo.solve_direct(A, b)
        assert fo.relative_l2(x, 350 - 2.5 * c[:, 2]) < 1e-13
    K = fo.assemble_matrix(t, fo.local_laplace(c, t, 20.0), c.shape[0])
    assert np.abs(K @ np.ones(c.shape[0])).max() < 1e-11              # stiffness row sums
    A, b = fo.heat_system(c, t, 20.0, [(d1, 350.0), (d2, 300.0)])
    x, it, rel = fo.pcg_jacobi(A, b, rtol=1e-12)
    assert 100 <= it <= 140 and fo.relative_l2(x, e["solution"]) < 1e-10


@pytest.mark.parametrize("N", [2, 4, 8])
def test_cube_closed_forms(N):
    c, t = fo.unit_cube_mesh(N, N, N)
    assert c.shape[0] == (N + 1) ** 3 and t.shape[0] == 6 * N ** 3
    rp, ci = fo.csr_pattern(t, c.shape[0])
    assert ci.size == (N + 1) ** 3 + 2 * (3 * N * (N + 1) ** 2 + 3 * N * N * (N + 1) + N ** 3)
    vol, _ = fo.p1_geometry(c, t)
    assert abs(vol.sum() - 1.0) < 1e-13
    if N == 8:
        k = 20.0
        K = fo.conform(fo.assemble_matrix(t, fo.local_laplace(c, t, k), c.shape[0]), rp, ci)
        p = N + 1
        r = 4 + 4 * p + 4 * p * p
        row = K.getrow(r)
        assert row.nnz == 15 and (np.abs(row.data) > 1e-12).sum() == 7
        assert abs(K[r, r] - 6 * k / N) < 1e-12 and abs(K[r, r + 1] + k / N) < 1e-12 and abs(K[r, r + p * p] + k / N) < 1e-12
        assert abs(K - K.T).max() < 1e-12


def test_patch_tests_linear_fields():
    """P1 reproduces linear fields exactly: heat (scalar) and elasticity (constant strain), jittered lattice."""
    N = 5
    c0, t = fo.unit_cube_mesh(N, N, N)
    rng = np.random.default_rng(0)
    c = c0 + 0.2 / N * (rng.random(c0.shape) * 2 - 1)
    bnd = np.any((c0 == 0) | (c0 == 1), axis=1)
    c[bnd] = c0[bnd]
    bv = np.nonzero(bnd)[0]
    lin = 3.0 + c @ np.array([1.0, -2.0, 0.5])
    A, b = fo.heat_system(c, t, 7.0, [(bv, lin[bv])])
    assert fo.relative_l2(fo.solve_direct(A, b), lin) < 1e-13
    mu, lam = fo.lame(10.0, 0.3)
    G = np.array([[0.01, 0.02, -0.01], [0.0, -0.02, 0.03], [0.015, 0.0, 0.01]])
    uex = (c @ G.T).reshape(-1)
    K = fo.assemble_matrix(t, fo.local_elasticity(c, t, mu, lam), c.shape[0], 3)
    dofs = (bv[:, None] * 3 + np.arange(3)).ravel()
    Ab, bb = fo.apply_dirichlet(K, np.zeros(uex.size), dofs, uex[dofs], symmetric=True)
    assert fo.relative_l2(fo.solve_direct(Ab, bb), uex) < 1e-12
    assert abs(K - K.T).max() < 1e-12 * abs(K).max()


def test_kat4_one_dimensional_profile_and_cg_tolerance():
    """Unit cube, Dirichlet on z faces, constant source: nodally exact T(z) (config C2's exact answer);
    rtol 1e-12 lands within 1e-10 of the direct solve, rtol 1e-8 does not."""
    N = 8
    c, t = fo.unit_cube_mesh(N, N, N)
    z0, z1 = np.nonzero(c[:, 2] == 0)[0], np.nonzero(c[:, 2] == 1)[0]
    A, b = fo.heat_system(c, t, 20.0, [(z0, 350.0), (z1, 300.0)], source=1000.0)
    xd = fo.solve_direct(A, b)
    z = c[:, 2]
    assert fo.relative_l2(xd, 350 - 50 * z + 1000 * z * (1 - z) / 40) < 1e-13
    x12, it12, _ = fo.pcg_jacobi(A, b, rtol=1e-12)
    x8, it8, _ = fo.pcg_jacobi(A, b, rtol=1e-6)
    assert fo.relative_l2(x12, xd) < 1e-10 < fo.relative_l2(x8, xd)
    xb, itb, _ = fo.bicgstab_jacobi(A, b, rtol=1e-12)
    assert fo.relative_l2(xb, xd) < 1e-10


def test_square_kat5_flux_boundary():
    """examples/test_heat_transfer.py geometry: top Dirichlet 360, bottom flux 36 with k=0.6 -> 360 + 60(1-y)."""
    c, t = fo.unit_square_mesh(40, 40)
    fv, opp, _ = fo.exterior_facets(t)
    mid = c[fv].mean(axis=1)
    A, b = fo.heat_system(c, t, 0.6, [(np.unique(fv[mid[:, 1] == 1]), 360.0)], neumann=[(fv[mid[:, 1] == 0], 36.0)])
    assert fo.relative_l2(fo.solve_direct(A, b), 360 + 60 * (1 - c[:, 1])) < 1e-12
    assert abs(fo.boundary_area(c, fv) - 4.0) < 1e-13
    meas, nrm = fo.facet_measure(c, fv, opp)
    assert np.allclose(nrm[mid[:, 1] == 0], [0, -1]) and np.allclose(nrm[mid[:, 0] == 1], [1, 0])


def test_element_matrices_against_quadrature():
    """Closed-form P1 matrices against brute-force quadrature on a random tetrahedron (mass, advection,
    tensor Laplace), so the closed forms restate what FFC's generated kernels integrate."""
    rng = np.random.default_rng(3)
    c = rng.random((4, 3))
    t = np.array([[0, 1, 2, 3]], dtype=np.int32)
    vol, G = fo.p1_geometry(c, t)
    # degree-2 exact 4-point rule on the tetrahedron
    a, b_ = 0.5854101966249685, 0.1381966011250105
    pts = np.full((4, 4), b_) + (a - b_) * np.eye(4)          # barycentric
    w = vol[0] / 4
    M = sum(w * np.outer(p, p) for p in pts)
    assert np.allclose(fo.local_mass(c, t, 1.0)[0], M, rtol=1e-13, atol=1e-16)
    v = np.array([0.3, -0.2, 0.5])
    C = sum(w * np.outer(p, G[0] @ v) for p in pts)
    assert np.allclose(fo.local_advection(c, t, v, 1.0)[0], C, rtol=1e-13, atol=1e-16)
    vn = rng.random((4, 3))
    Cn = sum(w * np.outer(p, G[0] @ (p @ vn)) for p in pts)
    assert np.allclose(fo.local_advection(c, t, vn, 1.0)[0], Cn, rtol=1e-12, atol=1e-16)
    Kt = rng.random((3, 3))
    assert np.allclose(fo.local_laplace(c, t, Kt)[0], vol[0] * G[0] @ Kt @ G[0].T, rtol=1e-13)


def test_dirichlet_variants_give_same_solution():
    c, t = fo.unit_square_mesh(9, 7)
    fv, _, _ = fo.exterior_facets(t)
    bv = np.unique(fv)
    g = np.sin(3 * c[bv, 0]) + c[bv, 1]
    A1, b1 = fo.heat_system(c, t, 2.0, [(bv, g)], source=5.0, symmetric=True)
    A2, b2 = fo.heat_system(c, t, 2.0, [(bv, g)], source=5.0, symmetric=False)
    assert abs(A1 - A1.T).max() < 1e-13 and abs(A2 - A2.T).max() > 1e-3
    assert fo.relative_l2(fo.solve_direct(A1, b1), fo.solve_direct(A2, b2)) < 1e-13
    assert A1.nnz == A2.nnz == fo.csr_pattern(t, c.shape[0])[1].size           # pattern kept (explicit zeros)


# ---------------------------------------------------------------------------------- C restatement
def test_c_oracle_matches_numpy_oracle():
    c, t = co.box_mesh((5, 4, 3))
    c0, t0 = fo.unit_cube_mesh(5, 4, 3)
    assert np.array_equal(c, c0) and np.array_equal(t, t0)
    rp, ci = co.csr_pattern(t, c.shape[0])
    rp0, ci0 = fo.csr_pattern(t0, c0.shape[0])
    assert np.array_equal(rp, rp0) and np.array_equal(ci, ci0)
    N = 12
    h = co.HeatCube(N)
    r = h.step(rtol=1e-12)
    c, t = fo.unit_cube_mesh(N, N, N)
    z0, z1 = np.nonzero(c[:, 2] == 0)[0], np.nonzero(c[:, 2] == 1)[0]
    A, b = fo.heat_system(c, t, 20.0, [(z0, 350.0), (z1, 300.0)], source=1000.0)
    assert np.abs(h.vals - A.data).max() <= 1e-13 * np.abs(A.data).max()
    assert np.abs(h.b - b).max() <= 1e-13 * np.abs(b).max()
    x0 = np.full(c.shape[0], 293.0); x0[z0] = 350; x0[z1] = 300      # HeatCube's start: initial field with the BCs imposed
    x, it, _ = fo.pcg_jacobi(A, b, x0=x0, rtol=1e-12)
    assert abs(r["iterations"] - it) <= 1
    assert fo.relative_l2(h.x, fo.solve_direct(A, b)) < 1e-10
    assert co.num_threads() >= 1


def test_fenics_tutorial_poisson_is_nodally_exact():
    """The first program of the FEniCS tutorial (Langtangen & Logg, "Solving PDEs in Python", ft01_poisson.py): -Laplace u = -6
    on UnitSquareMesh(8, 8) with u_D = 1 + x^2 + 2 y^2 on the whole boundary, P1.  dolfin's published output is
    error_max = O(1e-15): on this structured triangulation the P1 Galerkin solution interpolates the quadratic exactly.  The same
    must hold for the oracle on the dolfin-layout mesh (a wrong diagonal direction or vertex order would still be exact here, a
    wrong load or stiffness scaling would not)."""
    for n in (8, 16):
        c, t = fo.unit_square_mesh(n, n)
        nv = c.shape[0]
        uD = 1 + c[:, 0] ** 2 + 2 * c[:, 1] ** 2
        bnd = np.nonzero((c[:, 0] == 0) | (c[:, 0] == 1) | (c[:, 1] == 0) | (c[:, 1] == 1))[0]
        A, b = fo.heat_system(c, t, 1.0, [(bnd, uD[bnd])], source=-6.0)
        u = fo.solve_direct(A, b)
        assert np.abs(u - uD).max() < 5e-14
    # the program also prints errornorm(u_D, u, 'L2'); the tutorial's published output is  error_L2 = 0.00823509807335  (and
    # error_max = 1.33226762955e-15).  With u nodally exact that number is the L2 norm of the P1 interpolation error of u_D on
    # UnitSquareMesh(8, 8): a golden value from dolfin's own documentation that the mesh layout + element geometry reproduce.
    from oracle import fem_oracle_p2 as fp
    c, t = fo.unit_square_mesh(8, 8)
    uD = lambda x: 1 + x[..., 0] ** 2 + 2 * x[..., 1] ** 2       # noqa: E731
    nv = c.shape[0]
    bnd = np.nonzero((c[:, 0] == 0) | (c[:, 0] == 1) | (c[:, 1] == 0) | (c[:, 1] == 1))[0]
    A, b = fo.heat_system(c, t, 1.0, [(bnd, uD(c)[bnd])], source=-6.0)
    uh = fo.solve_direct(A, b)
    pts, w = fp._collapsed_rule(2, 6)
    vol, _ = fo.p1_geometry(c, t)
    xq = np.einsum('pa,cai->cpi', pts, c[t])
    e = uD(xq) - np.einsum('pa,ca->cp', pts, uh[t])
    error_L2 = float(np.sqrt(np.sum(vol[:, None] * w[None, :] * e ** 2)))
    assert abs(error_L2 - 0.00823509807335) < 5e-15
    # and in 3D on the dolfin-layout cube: u = 1 + x^2 + 2 y^2 + 3 z^2, f = -12
    c, t = fo.unit_cube_mesh(4, 4, 4)
    uD = 1 + c[:, 0] ** 2 + 2 * c[:, 1] ** 2 + 3 * c[:, 2] ** 2
    bnd = np.nonzero(np.any((c == 0) | (c == 1), axis=1))[0]
    A, b = fo.heat_system(c, t, 1.0, [(bnd, uD[bnd])], source=-12.0)
    assert np.abs(fo.solve_direct(A, b) - uD).max() < 5e-14


def test_fenics_tutorial_heat_equation_crank_nicolson_is_nodally_exact():
    """The tutorial's heat-equation test problem (ft03_heat.py): u = 1 + x^2 + alpha y^2 + beta t, f = beta - 2 - 2 alpha, which dolfin
    reproduces to 1e-15 at every step because the quadratic is interpolated exactly and the time dependence is linear.  The reference's
    time scheme (ScalarTransportSolver.py:287-293) is Crank-Nicolson with the load not theta-weighted; it is exact for this solution
    too: (c/dt) M (T - T_prev) + K (T + T_prev)/2 = int S q, with c = 1, k = 1, S = beta - 2 - 2 alpha."""
    alpha, beta, dt, nsteps = 3.0, 1.2, 0.3, 5
    c, t = fo.unit_square_mesh(8, 8)
    nv = c.shape[0]
    K = fo.assemble_matrix(t, fo.local_laplace(c, t, 1.0), nv)
    M = fo.assemble_matrix(t, fo.local_mass(c, t, 1.0), nv)
    load = fo.assemble_source(c, t, beta - 2 - 2 * alpha)
    bnd = np.nonzero((c[:, 0] == 0) | (c[:, 0] == 1) | (c[:, 1] == 0) | (c[:, 1] == 1))[0]
    exact = lambda tt: 1 + c[:, 0] ** 2 + alpha * c[:, 1] ** 2 + beta * tt      # noqa: E731
    T = exact(0.0)
    for n in range(1, nsteps + 1):
        A = M / dt + 0.5 * K
        b = (M / dt) @ T - 0.5 * (K @ T) + load
        Ab, bb = fo.apply_dirichlet(A, b, bnd, exact(n * dt)[bnd], symmetric=False)
        T = fo.solve_direct(Ab, bb)
        assert np.abs(T - exact(n * dt)).max() < 2e-13


def test_c_oracle_elasticity_and_transient_match_numpy_oracle():
    """The C/OpenMP restatements behind the bench's c3 / c4 CPU figures (oracle/fem_oracle_c.c fo_assemble_elasticity,
    fo_assemble_scalar, fo_apply_scalar, fo_bicgstab_jacobi) against the numpy oracle on a 6^3 cube: expanded CSR pattern
    bit-exact, matrix 1e-14, cantilever solution and three Crank-Nicolson steps against direct solves."""
    import scipy.sparse as sp
    from oracle import c_oracle as co
    N = 6
    c, t = fo.unit_cube_mesh(N, N, N)
    ec = co.ElasticityCube(N)
    ec.step()
    rp0, ci0 = fo.csr_pattern(t, c.shape[0], ncomp=3)
    assert np.array_equal(rp0, ec.rp) and np.array_equal(ci0, ec.ci)
    rp1, ci1 = co.expand_pattern(*co.csr_pattern(t, c.shape[0]), 3)
    assert np.array_equal(rp1, ec.rp) and np.array_equal(ci1, ec.ci)
    A = fo.conform(fo.assemble_matrix(t, fo.local_elasticity(c, t, ec.mu, ec.lam), c.shape[0], ncomp=3), rp0, ci0)
    b = fo.assemble_source(c, t, ec.f, ncomp=3)
    dofs = np.nonzero(ec.flag)[0]
    A2, b2 = fo.apply_dirichlet(A, b, dofs, np.zeros(dofs.size), True)
    Ac = sp.csr_matrix((ec.vals, ec.ci, ec.rp), shape=(ec.n, ec.n))
    assert abs(Ac - A2).max() <= 1e-14 * abs(A2).max()
    assert fo.relative_l2(ec.x, fo.solve_direct(A2, b2)) < 1e-10
    tc = co.TransientCube(N)
    T = np.full(tc.nv, 300.0)
    K = fo.assemble_matrix(t, fo.local_laplace(c, t, tc.k), c.shape[0])
    M = fo.assemble_matrix(t, fo.local_mass(c, t, tc.c), c.shape[0])
    Cm = fo.assemble_matrix(t, fo.local_advection(c, t, tc.vel, tc.c), c.shape[0])
    rp, ci = fo.csr_pattern(t, c.shape[0])
    d = np.nonzero(tc.flag)[0]
    for _ in range(3):
        tc.step()
        A2, b2 = fo.apply_dirichlet(fo.conform(M / tc.dt + 0.5 * K + Cm, rp, ci), (M / tc.dt - 0.5 * K) @ T, d, tc.g[d], False)
        T = fo.solve_direct(A2, b2)
        assert fo.relative_l2(tc.T, T) < 1e-10
