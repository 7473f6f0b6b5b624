"""GPU parity tests, kernel level: every libfsb entry point against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star): integer work (connectivity, CSR row_ptr/col_idx) bit-exact; assembled
values to 1e-13 of the row scale (atomic summation order differs from the oracle's); solution vectors
within 1e-10 relative L2 of the oracle's direct solve.
"""
import os

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu

from oracle import fem_oracle as fo  # noqa: E402  (checker only)
from fenicssolver_b200 import _lib  # noqa: E402

VAL_TOL = 1e-13
SOL_TOL = 1e-10


def csr_from_device(A):
    rp, ci, va = A.download_csr()
    n = rp.size - 1
    return sp.csr_matrix((va, ci.astype(np.int64), rp), shape=(n, n)), rp, ci


def assert_vals_close(dev, ref, tol=VAL_TOL):
    assert dev.shape == ref.shape
    scale = np.abs(ref).max()
    err = np.abs(dev - ref).max()
    assert err <= tol * scale, "max abs err %.3e vs scale %.3e" % (err, scale)


def jitter(coords, n, seed=0, amp=0.2):
    rng = np.random.default_rng(seed)
    return coords + amp / n * (rng.random(coords.shape) * 2 - 1)


# ----------------------------------------------------------------------------------- meshes
@pytest.mark.parametrize("n", [(3, 2), (5, 7), (4, 3, 2), (6, 5, 7)])
def test_box_mesh_bit_exact(ctx, n):
    if len(n) == 2:
        p0, p1 = (0.25, -1.0), (2.0, 0.5)
        c, t = fo.rectangle_mesh(p0[0], p0[1], p1[0], p1[1], *n)
    else:
        p0, p1 = (0.0, -1.0, 0.5), (10.0, 1.0, 1.75)
        c, t = fo.box_mesh(p0, p1, *n)
    m = _lib.DeviceMesh.box(ctx, n, p0, p1)
    xyz, cells = m.download()
    assert np.array_equal(cells, t)
    assert np.array_equal(xyz, c)          # bit-exact coordinates


# ----------------------------------------------------------------------------------- K1: boundary search + facet numbering
def _shuffled(c, t, seed):
    """The same mesh with its cells in a random order and its vertices renumbered at random (cells re-sorted per cell)."""
    rng = np.random.default_rng(seed)
    perm = rng.permutation(c.shape[0])
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.size)
    t2 = np.sort(inv[t], axis=1).astype(np.int32)
    return c[perm], t2[rng.permutation(t2.shape[0])]


@pytest.mark.parametrize("case", ["fixture", "cube", "cube_shuffled", "square", "square_shuffled", "two_cells", "non_manifold"])
def test_exterior_facets_and_facet_ids_bit_exact(ctx, golden_dir, case):
    """fsb_mesh_exterior_facets (csrc/fsb_facets.cu) against oracle.fem_oracle.facet_table / exterior_facets: the same facets in
    the same (lexicographic) order, opposite vertices, owning cells and dolfin facet ids — integer work, bit-exact.  The shipped
    fixture mesh is the one whose facet numbering was verified against data/mesh_facet_region.xml."""
    if case == "fixture":
        g = np.load(os.path.join(golden_dir, "fixture_mesh.npz"))
        c, t = g["coords"], g["cells"]
    elif case.startswith("cube"):
        c, t = fo.unit_cube_mesh(5, 4, 6)
        if case.endswith("shuffled"):
            c, t = _shuffled(c, t, 3)
    elif case.startswith("square"):
        c, t = fo.unit_square_mesh(9, 7)
        if case.endswith("shuffled"):
            c, t = _shuffled(c, t, 4)
    elif case == "two_cells":
        c = np.array([[0.0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 1]])
        t = np.array([[0, 1, 2, 3], [1, 2, 3, 4]], dtype=np.int32)
    else:
        # three tetrahedra around one triangle (not a manifold): that facet has three holders, is interior, and is counted once
        c = np.array([[0.0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [0, 0, -1], [1, 1, 0.3]])
        t = np.array([[0, 1, 2, 3], [0, 1, 2, 4], [0, 1, 2, 5]], dtype=np.int32)
    t = np.ascontiguousarray(t, dtype=np.int32)
    m = _lib.DeviceMesh.upload(ctx, c, t)
    fv, opp, cell, fid = m.exterior_facets()
    facets, cf, count = fo.facet_table(t)
    f0, o0, id0 = fo.exterior_facets(t)
    assert m.num_facets == facets.shape[0]
    assert np.array_equal(fv, f0) and np.array_equal(opp, o0) and np.array_equal(fid, id0)
    # the owning cell really holds the facet opposite `opp`
    assert np.all(np.sort(np.hstack([fv, opp[:, None]]), axis=1) == t[cell])
    # second request: served from the mesh, same arrays
    fv2, opp2, cell2, fid2 = m.exterior_facets()
    assert np.array_equal(fv, fv2) and np.array_equal(fid, fid2)


@pytest.mark.parametrize("dim", [2, 3])
def test_boundary_geometry_for_subdomain_marking(ctx, dim):
    """fsb_mesh_boundary_geometry: distinct boundary vertices, facets in terms of them, coordinates and midpoints — bit-exact against
    the numpy statement (np.unique / fancy indexing / mean) that FacetMarkers.mark_subdomain used on the host."""
    c, t = (fo.unit_square_mesh(7, 5) if dim == 2 else fo.unit_cube_mesh(4, 5, 3))
    c = jitter(c, 7, seed=11)
    m = _lib.DeviceMesh.upload(ctx, c, np.ascontiguousarray(t, dtype=np.int32))
    fv, _, _, _ = m.exterior_facets()
    bv, finv, bxyz, mid = m.boundary_geometry()
    uv, inv = np.unique(fv, return_inverse=True)
    assert np.array_equal(bv, uv) and np.array_equal(finv, inv.reshape(fv.shape))
    assert np.array_equal(bxyz, c[uv]) and np.array_equal(mid, c[uv][inv.reshape(fv.shape)].mean(axis=1))


def test_exterior_facets_through_mesh_api(ctx):
    """Mesh.exterior_facets()/exterior_facet_ids() of an array mesh and of a generated box both come from the device search and agree
    with the direct enumeration of the dolfin box layout."""
    from fenicssolver_b200.dolfin_compat import Mesh, UnitCubeMesh, box_exterior_facets
    box = UnitCubeMesh(6, 5, 4)
    fv, opp = box.exterior_facets()
    f1, o1 = box_exterior_facets((6, 5, 4))
    assert np.array_equal(fv, f1) and np.array_equal(opp, o1)
    arr = Mesh(box.coordinates().copy(), box.cells().copy())
    fv2, opp2 = arr.exterior_facets()
    assert np.array_equal(fv2, f1) and np.array_equal(opp2, o1)
    assert np.array_equal(arr.exterior_facet_ids(), fo.exterior_facets(box.cells())[2])


def test_box_mesh_slab_matches_global(ctx):
    n = (4, 3, 6)
    c, t = fo.unit_cube_mesh(*n)
    plane = (n[0] + 1) * (n[1] + 1)
    m = _lib.DeviceMesh.box(ctx, n, (0, 0, 0), (1, 1, 1), 2, 5)
    xyz, cells = m.download()
    assert np.array_equal(xyz, c[2 * plane:6 * plane])
    ncl = 6 * n[0] * n[1]
    assert np.array_equal(cells + 2 * plane, t[2 * ncl:5 * ncl])


# ----------------------------------------------------------------------------------- pattern
@pytest.mark.parametrize("n", [(4, 4), (9, 5), (2, 2, 2), (8, 8, 8), (5, 3, 4)])
def test_pattern_exact_box(ctx, n):
    c, t = (fo.rectangle_mesh(0, 0, 1, 1, *n) if len(n) == 2 else fo.unit_cube_mesh(*n))
    m = _lib.DeviceMesh.box(ctx, n, (0,) * len(n), (1,) * len(n))
    A = _lib.DeviceMatrix.create(m, 1)
    rp, ci, _ = A.download_csr(values=False)
    rp0, ci0 = fo.csr_pattern(t, c.shape[0])
    assert np.array_equal(rp, rp0) and np.array_equal(ci, ci0)
    if len(n) == 3 and n[0] == n[1] == n[2]:
        N = n[0]
        assert ci.size == (N + 1) ** 3 + 2 * (3 * N * (N + 1) ** 2 + 3 * N * N * (N + 1) + N ** 3)


def test_pattern_exact_fixture_golden(ctx, golden_dir):
    g = np.load(os.path.join(golden_dir, "fixture_mesh.npz"))
    e = np.load(os.path.join(golden_dir, "fixture_expected.npz"))
    m = _lib.DeviceMesh.upload(ctx, g["coords"], g["cells"])
    A = _lib.DeviceMatrix.create(m, 1)
    rp, ci, _ = A.download_csr(values=False)
    assert np.array_equal(rp, e["row_ptr"]) and np.array_equal(ci, e["col_idx"])
    assert ci.size == 13315


@pytest.mark.parametrize("ncomp,n", [(2, (4, 3)), (3, (3, 2, 4))])
def test_pattern_exact_vector(ctx, ncomp, n):
    c, t = (fo.rectangle_mesh(0, 0, 1, 1, *n) if len(n) == 2 else fo.unit_cube_mesh(*n))
    m = _lib.DeviceMesh.upload(ctx, c, t)
    A = _lib.DeviceMatrix.create(m, ncomp)
    rp, ci, _ = A.download_csr(values=False)
    rp0, ci0 = fo.csr_pattern(t, c.shape[0], ncomp)
    assert np.array_equal(rp, rp0) and np.array_equal(ci, ci0)


def test_pattern_high_valence_vertex(ctx):
    """A fan of triangles around one vertex: valence above the shared-memory row cache (uncached path)."""
    k = 200
    ang = np.linspace(0, 2 * np.pi, k, endpoint=False)
    coords = np.vstack([[0.0, 0.0], np.stack([np.cos(ang), np.sin(ang)], 1)])
    cells = np.sort(np.stack([np.zeros(k, int), 1 + np.arange(k), 1 + (np.arange(k) + 1) % k], 1), axis=1).astype(np.int32)
    m = _lib.DeviceMesh.upload(ctx, coords, cells)
    A = _lib.DeviceMatrix.create(m, 1)
    rp, ci, _ = A.download_csr(values=False)
    rp0, ci0 = fo.csr_pattern(cells, coords.shape[0])
    assert np.array_equal(rp, rp0) and np.array_equal(ci, ci0)
    assert (rp[1] - rp[0]) == k + 1


# ----------------------------------------------------------------------------------- assembly
@pytest.mark.parametrize("asm_mode", [0, 1, 2, 3])
@pytest.mark.parametrize("dim", [2, 3])
def test_assemble_scalar_terms(ctx, dim, asm_mode):
    n = (7, 5) if dim == 2 else (5, 4, 6)
    c, t = (fo.rectangle_mesh(0, 0, 2, 1, *n) if dim == 2 else fo.box_mesh((0, 0, 0), (2, 1, 3), *n))
    c = jitter(c, max(n), seed=1)
    nv = c.shape[0]
    m = _lib.DeviceMesh.upload(ctx, c, t)
    ctx.set_option("asm_mode", asm_mode)
    try:
        rng = np.random.default_rng(2)
        kt = rng.random((dim, dim)) + dim * np.eye(dim)
        vel = rng.random(dim) - 0.5
        cases = [
            dict(kscale=20.0),
            dict(kscale=0.0, mass=3.5),
            dict(kscale=0.0, adv=2.0, vel=vel),
            dict(kscale=0.7, ktensor=kt, mass=1.25, adv=4.0, vel=vel),
        ]
        rp0, ci0 = fo.csr_pattern(t, nv)
        for kw in cases:
            A = _lib.DeviceMatrix.create(m, 1)
            A.assemble_scalar(**kw)
            dev, rp, ci = csr_from_device(A)
            Ke = kw.get("kscale", 1.0) * fo.local_laplace(c, t, kw.get("ktensor", 1.0))
            if kw.get("mass"):
                Ke = Ke + fo.local_mass(c, t, kw["mass"])
            if kw.get("adv"):
                Ke = Ke + fo.local_advection(c, t, kw["vel"], kw["adv"])
            ref = fo.conform(fo.assemble_matrix(t, Ke, nv), rp0, ci0)
            assert np.array_equal(rp, rp0) and np.array_equal(ci, ci0)
            assert_vals_close(dev.data, ref.data)
    finally:
        ctx.set_option("asm_mode", 1)


@pytest.mark.parametrize("dim", [2, 3])
def test_row_gather_assembly_overwrite_action_and_reproducibility(ctx, dim):
    """asm_mode 2 (opt-in), k_scalar_rows: `overwrite` equals zero + add on a matrix holding garbage, the matrix-free action equals
    A x, two assemblies are bitwise identical (fixed summation order — the atomic scatter cannot promise that), and the result equals
    the scatter kernel's to rounding."""
    n = (9, 6) if dim == 2 else (6, 5, 4)
    c, t = (fo.rectangle_mesh(0, 0, 2, 1, *n) if dim == 2 else fo.box_mesh((0, 0, 0), (2, 1, 3), *n))
    c = jitter(c, max(n), seed=5)
    nv = c.shape[0]
    m = _lib.DeviceMesh.upload(ctx, c, t)
    rng = np.random.default_rng(7)
    vel = rng.random(dim) - 0.5
    kw = dict(kscale=0.7, mass=1.25, adv=4.0, vel=vel)
    A = _lib.DeviceMatrix.create(m, 1)
    try:
        ctx.set_option("asm_mode", 2)
        A.assemble_scalar(kscale=123.0, mass=7.0)                    # garbage to be overwritten
        A.assemble_scalar(overwrite=True, **kw)
        v1 = A.download_csr()[2].copy()
        A.zero()
        A.assemble_scalar(**kw)
        v2 = A.download_csr()[2].copy()
        assert np.array_equal(v1, v2)                                # bitwise: same kernel, same order, with and without the zero-fill
        A.assemble_scalar(overwrite=True, **kw)
        assert np.array_equal(A.download_csr()[2], v1)
        ctx.set_option("asm_mode", 1)
        B = _lib.DeviceMatrix.create(m, 1)
        B.assemble_scalar(overwrite=True, **kw)
        assert_vals_close(B.download_csr()[2], v1)
        ctx.set_option("asm_mode", 2)
        xh = rng.standard_normal(nv)
        x, y = _lib.DeviceVector.from_numpy(ctx, xh), _lib.DeviceVector.from_numpy(ctx, np.ones(nv))
        _lib.apply_scalar(m, x, y, **kw)
        Am, _, _ = csr_from_device(A)
        ref = 1.0 + Am @ xh
        assert np.abs(y.numpy() - ref).max() <= 1e-13 * np.abs(ref).max()
    finally:
        ctx.set_option("asm_mode", 1)


def test_combine_plan_assembly_on_fixture_and_ragged_meshes(ctx, golden_dir):
    """asm_mode 3 (per-warp combine plan) on meshes whose cell count is not a multiple of 32 and whose numbering is unstructured
    (the reference's shipped fixture mesh, a shuffled cube): same matrix as the oracle, accumulation across calls, reuse of the plan."""
    g = np.load(os.path.join(golden_dir, "fixture_mesh.npz"))
    cases = [(g["coords"], np.ascontiguousarray(g["cells"], dtype=np.int32)), _shuffled(*fo.unit_cube_mesh(5, 3, 4), seed=9),
             (np.array([[0.0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]]), np.array([[0, 1, 2, 3]], dtype=np.int32))]
    ctx.set_option("asm_mode", 3)
    try:
        for c, t in cases:
            nv = c.shape[0]
            m = _lib.DeviceMesh.upload(ctx, c, t)
            A = _lib.DeviceMatrix.create(m, 1)
            A.assemble_scalar(kscale=2.0, mass=0.5)
            A.assemble_scalar(kscale=1.0)                        # second call: the plan is reused and the values accumulate
            dev, rp, ci = csr_from_device(A)
            rp0, ci0 = fo.csr_pattern(t, nv)
            ref = fo.conform(fo.assemble_matrix(t, 3.0 * fo.local_laplace(c, t, 1.0) + fo.local_mass(c, t, 0.5), nv), rp0, ci0)
            assert np.array_equal(rp, rp0) and np.array_equal(ci, ci0)
            assert_vals_close(dev.data, ref.data)
            A.assemble_scalar(kscale=1.0, overwrite=True)
            ref1 = fo.conform(fo.assemble_matrix(t, fo.local_laplace(c, t, 1.0), nv), rp0, ci0)
            assert_vals_close(A.download_csr()[2], ref1.data)
    finally:
        ctx.set_option("asm_mode", 1)


def test_assemble_accumulates_and_zero(ctx):
    c, t = fo.unit_cube_mesh(3, 3, 3)
    m = _lib.DeviceMesh.upload(ctx, c, t)
    A = _lib.DeviceMatrix.create(m, 1)
    A.assemble_scalar(kscale=1.0)
    A.assemble_scalar(kscale=2.0)
    dev, _, _ = csr_from_device(A)
    ref = fo.assemble_matrix(t, 3.0 * fo.local_laplace(c, t), c.shape[0])
    assert_vals_close(dev.toarray(), ref.toarray())
    A.zero()
    assert np.all(csr_from_device(A)[0].data == 0.0)


def test_cube_stencil_invariants(ctx):
    """SURVEY 8c KAT 3: 15 structural / 7 numerical non-zeros per interior row, 7-point stencil k*h*(6,-1..)."""
    N, k = 8, 20.0
    m = _lib.DeviceMesh.box(ctx, (N, N, N), (0, 0, 0), (1, 1, 1))
    A = _lib.DeviceMatrix.create(m, 1)
    A.assemble_scalar(kscale=k)
    M, rp, ci = csr_from_device(A)
    p = N + 1
    r = 4 + 4 * p + 4 * p * p
    row = M.getrow(r)
    assert row.nnz == 15
    h = 1.0 / N
    dense = np.zeros(M.shape[0]); dense[row.indices] = row.data
    assert abs(dense[r] - 6 * k * h) < 1e-12
    for off in (1, p, p * p):
        assert abs(dense[r + off] + k * h) < 1e-12 and abs(dense[r - off] + k * h) < 1e-12
    assert (np.abs(row.data) > 1e-12).sum() == 7
    assert np.abs(M @ np.ones(M.shape[0])).max() < 1e-11
    assert abs(M - M.T).max() < 1e-12


@pytest.mark.parametrize("dim", [2, 3])
def test_apply_scalar_matches_matrix_action(ctx, dim):
    n = (6, 5) if dim == 2 else (4, 5, 3)
    c, t = (fo.rectangle_mesh(0, 0, 1, 1, *n) if dim == 2 else fo.unit_cube_mesh(*n))
    c = jitter(c, max(n), seed=3)
    nv = c.shape[0]
    m = _lib.DeviceMesh.upload(ctx, c, t)
    rng = np.random.default_rng(4)
    xh = rng.random(nv)
    x = _lib.DeviceVector.from_numpy(ctx, xh)
    y = _lib.DeviceVector(ctx, nv)
    y.fill(1.0)
    _lib.apply_scalar(m, x, y, kscale=-0.5 * 0.6, mass=4.2e3)
    ref = 1.0 + fo.assemble_matrix(t, -0.3 * fo.local_laplace(c, t) + fo.local_mass(c, t, 4.2e3), nv) @ xh
    assert_vals_close(y.numpy(), ref)


@pytest.mark.parametrize("asm_mode", [0, 1, 2])
@pytest.mark.parametrize("dim", [2, 3])
def test_assemble_elasticity(ctx, dim, asm_mode):
    n = (5, 4) if dim == 2 else (3, 4, 3)
    c, t = (fo.rectangle_mesh(0, 0, 2, 1, *n) if dim == 2 else fo.box_mesh((0, 0, 0), (4, 1, 1), *n))
    c = jitter(c, max(n), seed=5)
    nv = c.shape[0]
    mu, lam = fo.lame(2e11, 0.27)
    m = _lib.DeviceMesh.upload(ctx, c, t)
    ctx.set_option("asm_mode", asm_mode)
    try:
        A = _lib.DeviceMatrix.create(m, dim)
        A.assemble_elasticity(mu, lam)
    finally:
        ctx.set_option("asm_mode", 1)
    dev, rp, ci = csr_from_device(A)
    rp0, ci0 = fo.csr_pattern(t, nv, dim)
    ref = fo.conform(fo.assemble_matrix(t, fo.local_elasticity(c, t, mu, lam), nv, dim), rp0, ci0)
    assert np.array_equal(rp, rp0) and np.array_equal(ci, ci0)
    assert_vals_close(dev.data, ref.data)


@pytest.mark.parametrize("dim", [2, 3])
def test_rhs_terms(ctx, dim):
    n = (6, 4) if dim == 2 else (4, 3, 5)
    c, t = (fo.rectangle_mesh(0, 0, 1, 2, *n) if dim == 2 else fo.box_mesh((0, 0, 0), (1, 2, 1), *n))
    c = jitter(c, max(n), seed=6)
    nv = c.shape[0]
    m = _lib.DeviceMesh.upload(ctx, c, t)
    fverts, opp, _ = fo.exterior_facets(t)
    rng = np.random.default_rng(7)
    # constant scalar source, restricted to tagged cells
    tags = (rng.random(t.shape[0]) < 0.5).astype(np.int32) + 1
    b = _lib.DeviceVector(ctx, nv)
    _lib.assemble_source(m, b, 1000.0)
    _lib.assemble_source(m, b, 3.0, scale=-2.0, cell_tags=tags, tag=2)
    ref = fo.assemble_source(c, t, 1000.0) - 2.0 * fo.assemble_source(c, t, 3.0, cell_mask=(tags == 2))
    assert_vals_close(b.numpy(), ref)
    # nodal source and vector constant source
    Sn = rng.random(nv * dim)
    bv = _lib.DeviceVector(ctx, nv * dim)
    S = _lib.DeviceVector.from_numpy(ctx, Sn)
    _lib.assemble_source_nodal(m, bv, S, ncomp=dim, scale=1.5)
    gvec = rng.random(dim)
    _lib.assemble_source(m, bv, gvec, ncomp=dim)
    ref = 1.5 * fo.assemble_source(c, t, Sn.reshape(nv, dim), ncomp=dim) + fo.assemble_source(c, t, gvec, ncomp=dim)
    assert_vals_close(bv.numpy(), ref)
    # facet loads: scalar flux, vector traction, pressure along the outward normal
    sel = rng.random(fverts.shape[0]) < 0.6
    b2 = _lib.DeviceVector(ctx, nv)
    _lib.assemble_facet_load(m, b2, fverts[sel], 36.0)
    assert_vals_close(b2.numpy(), fo.assemble_facet_load(c, fverts[sel], 36.0, nv))
    b3 = _lib.DeviceVector(ctx, nv * dim)
    _lib.assemble_facet_load(m, b3, fverts[sel], gvec, ncomp=dim, scale=-1.0)
    _lib.assemble_facet_load(m, b3, fverts[sel], 1e6, ncomp=dim, opp=opp[sel], normal=True)
    meas, nrm = fo.facet_measure(c, fverts[sel], opp[sel])
    ref = -fo.assemble_facet_load(c, fverts[sel], gvec, nv, dim) + fo.assemble_facet_load(c, fverts[sel], 1e6 * nrm, nv, dim)
    assert_vals_close(b3.numpy(), ref)
    # boundary area functional and the Robin boundary matrix
    assert abs(_lib.facet_area(m, fverts[sel]) - fo.boundary_area(c, fverts[sel])) < 1e-12 * fo.boundary_area(c, fverts[sel])
    A = _lib.DeviceMatrix.create(m, 1)
    A.assemble_facet_mass(fverts[sel], 100.0)
    dev, rp, ci = csr_from_device(A)
    ref = fo.conform(fo._scatter(fverts[sel], fo.local_facet_mass(c, fverts[sel], 100.0), nv), rp, ci)
    assert_vals_close(dev.data, ref.data)


@pytest.mark.parametrize("symmetric", [False, True])
@pytest.mark.parametrize("ncomp", [1, 3])
def test_apply_dirichlet(ctx, ncomp, symmetric):
    c, t = fo.unit_cube_mesh(4, 3, 3)
    c = jitter(c, 4, seed=8)
    nv = c.shape[0]
    m = _lib.DeviceMesh.upload(ctx, c, t)
    A = _lib.DeviceMatrix.create(m, ncomp)
    if ncomp == 1:
        A.assemble_scalar(kscale=2.0, mass=0.3)
        ref = fo.assemble_matrix(t, 2.0 * fo.local_laplace(c, t) + fo.local_mass(c, t, 0.3), nv)
    else:
        A.assemble_elasticity(3.0, 5.0)
        ref = fo.assemble_matrix(t, fo.local_elasticity(c, t, 3.0, 5.0), nv, 3)
    rng = np.random.default_rng(9)
    n = nv * ncomp
    bh = rng.random(n)
    dofs = np.sort(rng.choice(n, size=n // 5, replace=False))
    vals = rng.random(dofs.size) * 10
    b = _lib.DeviceVector.from_numpy(ctx, bh)
    x = _lib.DeviceVector(ctx, n)
    A.apply_dirichlet(b, dofs, vals, symmetric=symmetric, x=x)
    dev, rp, ci = csr_from_device(A)
    Aref, bref = fo.apply_dirichlet(fo.conform(ref, rp, ci), bh, dofs, vals, symmetric)
    assert_vals_close(dev.data, Aref.data)
    assert_vals_close(b.numpy(), bref)
    xh = x.numpy()
    assert np.array_equal(xh[dofs], vals) and np.count_nonzero(xh) <= dofs.size


# ----------------------------------------------------------------------------------- SpMV / Krylov
def random_csr(n, avg, seed, empty_rows=False, long_row=0):
    rng = np.random.default_rng(seed)
    M = sp.random(n, n, density=avg / n, format="lil", random_state=seed, data_rvs=lambda k: rng.random(k) - 0.5)
    if not empty_rows:
        M.setdiag(rng.random(n) + 1.0)
    if long_row:
        M[n // 2, :long_row] = rng.random(long_row)
    M = M.tocsr()
    M.sort_indices()
    return M


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("case", ["fem", "random", "ragged_empty_rows", "long_row", "tiny"])
def test_spmv_matches_scipy(ctx, case, mode):
    if case == "fem":
        c, t = fo.unit_cube_mesh(12, 11, 10)
        M = fo.assemble_matrix(t, fo.local_laplace(c, t, 3.0) + fo.local_mass(c, t), c.shape[0])
    elif case == "random":
        M = random_csr(20011, 9, 1)
    elif case == "ragged_empty_rows":
        M = random_csr(5003, 2, 2, empty_rows=True)
    elif case == "long_row":
        M = random_csr(3001, 5, 3, long_row=2500)
    else:
        M = random_csr(3, 2, 4)
    ctx.set_option("spmv_mode", mode)
    try:
        A = _lib.DeviceMatrix.from_csr(ctx, M.indptr, M.indices, M.data)
        xh = np.random.default_rng(5).random(M.shape[0]) - 0.5
        x = _lib.DeviceVector.from_numpy(ctx, xh)
        y = _lib.DeviceVector(ctx, M.shape[0])
        y.fill(7.0)
        A.spmv(x, y)
        yh = y.numpy()
    finally:
        ctx.set_option("spmv_mode", 0)
    ref = M @ xh
    assert np.abs(yh - ref).max() <= 1e-14 * max(1.0, np.abs(M).dot(np.abs(xh)).max())


def test_spmv_block3_matches_scalar_csr(ctx):
    c, t = fo.box_mesh((0, 0, 0), (4, 1, 1), 8, 5, 4)
    nv = c.shape[0]
    m = _lib.DeviceMesh.upload(ctx, c, t)
    A = _lib.DeviceMatrix.create(m, 3)
    A.assemble_elasticity(1.0, 2.0)
    M, _, _ = csr_from_device(A)
    xh = np.random.default_rng(6).random(3 * nv) - 0.5
    for mode in (0, 1):
        ctx.set_option("spmv_mode", mode)
        x = _lib.DeviceVector.from_numpy(ctx, xh)
        y = _lib.DeviceVector(ctx, 3 * nv)
        A.spmv(x, y)
        assert np.abs(y.numpy() - M @ xh).max() <= 1e-13 * np.abs(M @ xh).max()
    ctx.set_option("spmv_mode", 0)


def test_dot_is_reproducible(ctx):
    rng = np.random.default_rng(10)
    a, b = rng.random(1_000_003), rng.random(1_000_003)
    x, y = _lib.DeviceVector.from_numpy(ctx, a), _lib.DeviceVector.from_numpy(ctx, b)
    d1, d2 = x.dot(y), x.dot(y)
    assert d1 == d2
    assert abs(d1 - a @ b) <= 1e-12 * abs(a @ b)


def heat_problem(N, jit=False):
    c, t = fo.unit_cube_mesh(N, N, N)
    z0 = np.nonzero(c[:, 2] == 0)[0]
    z1 = np.nonzero(c[:, 2] == 1)[0]
    if jit:
        keep = c.copy()
        c = jitter(c, N, seed=11)
        bnd = (keep == 0) | (keep == 1)
        c[bnd] = keep[bnd]
    return c, t, z0, z1


@pytest.mark.parametrize("symmetric", [True, False])
def test_cg_heat_matches_oracle_direct_and_iteration_count(ctx, symmetric):
    """SURVEY 8c KAT 4 + iteration-for-iteration agreement with the oracle's PCG."""
    N = 16
    c, t, z0, z1 = heat_problem(N)
    nv = c.shape[0]
    Ao, bo = fo.heat_system(c, t, 20.0, [(z0, 350.0), (z1, 300.0)], source=1000.0, symmetric=symmetric)
    xd = fo.solve_direct(Ao, bo)
    m = _lib.DeviceMesh.box(ctx, (N, N, N), (0, 0, 0), (1, 1, 1))
    A = _lib.DeviceMatrix.create(m, 1)
    A.assemble_scalar(kscale=20.0)
    b = _lib.DeviceVector(ctx, nv)
    _lib.assemble_source(m, b, 1000.0)
    x = _lib.DeviceVector(ctx, nv)
    dofs = np.concatenate([z0, z1])
    vals = np.concatenate([np.full(z0.size, 350.0), np.full(z1.size, 300.0)])
    A.apply_dirichlet(b, dofs, vals, symmetric=symmetric, x=x)
    info = A.solve(b, x, "cg" if symmetric else "bicgstab", rtol=1e-12)
    xh = x.numpy()
    assert info["converged"] == 1
    assert fo.relative_l2(xh, xd) < SOL_TOL
    z = c[:, 2]
    assert fo.relative_l2(xh, 350 - 50 * z + 1000 * z * (1 - z) / 40) < SOL_TOL   # nodally exact profile
    x0 = np.zeros(nv); x0[dofs] = vals
    if symmetric:
        _, it, _ = fo.pcg_jacobi(Ao, bo, x0=x0, rtol=1e-12)
    else:
        _, it, _ = fo.bicgstab_jacobi(Ao, bo, x0=x0, rtol=1e-12)
    assert abs(info["iterations"] - it) <= max(2, it // 20)
    assert info["rnorm"] <= 1e-12 * info["bnorm"]


def test_cg_is_bitwise_reproducible(ctx):
    N = 12
    c, t, z0, z1 = heat_problem(N, jit=True)
    nv = c.shape[0]
    m = _lib.DeviceMesh.upload(ctx, c, t)
    dofs = np.concatenate([z0, z1])
    vals = np.concatenate([np.full(z0.size, 1.0), np.full(z1.size, 0.0)])
    sols = []
    A = _lib.DeviceMatrix.create(m, 1)
    A.assemble_scalar(kscale=1.0)
    b = _lib.DeviceVector(ctx, nv)
    _lib.assemble_source(m, b, 1.0)
    x0 = _lib.DeviceVector(ctx, nv)
    A.apply_dirichlet(b, dofs, vals, symmetric=True, x=x0)
    for _ in range(2):
        x = _lib.DeviceVector(ctx, nv)
        x.copy_from(x0)
        info = A.solve(b, x, "cg", rtol=1e-10)
        sols.append((x.numpy(), info["iterations"]))
    assert sols[0][1] == sols[1][1]
    assert np.array_equal(sols[0][0], sols[1][0])


def test_cg_maxit_and_zero_rhs(ctx):
    N = 8
    c, t, z0, z1 = heat_problem(N)
    nv = c.shape[0]
    m = _lib.DeviceMesh.upload(ctx, c, t)
    A = _lib.DeviceMatrix.create(m, 1)
    A.assemble_scalar(kscale=1.0, mass=1.0)
    b = _lib.DeviceVector(ctx, nv)
    x = _lib.DeviceVector(ctx, nv)
    info = A.solve(b, x, "cg", rtol=1e-12)          # b = 0: converged at iteration 0
    assert info["iterations"] == 0 and info["converged"] == 1
    _lib.assemble_source(m, b, 1.0)
    info = A.solve(b, x, "cg", rtol=1e-14, maxit=5)
    assert info["iterations"] == 5 and info["converged"] == 0


def test_fixture_kat_on_device(ctx, golden_dir):
    """data/TestHeatTransfer.json on data/mesh.xml: exact discrete answer T = 350 - 2.5 z (golden fixture)."""
    g = np.load(os.path.join(golden_dir, "fixture_mesh.npz"))
    e = np.load(os.path.join(golden_dir, "fixture_expected.npz"))
    c, t = g["coords"], g["cells"]
    nv = c.shape[0]
    m = _lib.DeviceMesh.upload(ctx, c, t)
    A = _lib.DeviceMatrix.create(m, 1)
    A.assemble_scalar(kscale=20.0)
    b = _lib.DeviceVector(ctx, nv)
    x = _lib.DeviceVector(ctx, nv)
    x.fill(293.0)
    dofs = np.concatenate([e["dofs_tag1"], e["dofs_tag2"]])
    vals = np.concatenate([np.full(e["dofs_tag1"].size, 350.0), np.full(e["dofs_tag2"].size, 300.0)])
    A.apply_dirichlet(b, dofs, vals, symmetric=True, x=x)
    info = A.solve(b, x, "cg", rtol=1e-13)
    assert info["converged"] == 1
    assert fo.relative_l2(x.numpy(), e["analytic"]) < SOL_TOL
    assert fo.relative_l2(x.numpy(), e["solution"]) < SOL_TOL


def test_elasticity_patch_and_cg(ctx):
    """Linear displacement field is reproduced exactly (constant strain patch test); CG on the 3x3-block matrix."""
    n = (6, 3, 3)
    c, t = fo.box_mesh((0, 0, 0), (4, 1, 1), *n)
    c0 = c.copy()
    c = jitter(c, 6, seed=12)
    bnd = np.zeros(c.shape[0], bool)
    for d, (lo, hi) in enumerate([(0, 4), (0, 1), (0, 1)]):
        bnd |= (c0[:, d] == lo) | (c0[:, d] == hi)
    c[bnd] = c0[bnd]
    nv = c.shape[0]
    mu, lam = fo.lame(10.0, 0.3)
    G = np.array([[0.01, 0.02, -0.01], [0.0, -0.02, 0.03], [0.015, 0.0, 0.01]])
    uex = (c @ G.T + np.array([0.1, -0.2, 0.3])).reshape(-1)
    bv = np.nonzero(bnd)[0]
    dofs = (bv[:, None] * 3 + np.arange(3)).reshape(-1)
    m = _lib.DeviceMesh.upload(ctx, c, t)
    A = _lib.DeviceMatrix.create(m, 3)
    A.assemble_elasticity(mu, lam)
    b = _lib.DeviceVector(ctx, 3 * nv)
    x = _lib.DeviceVector(ctx, 3 * nv)
    A.apply_dirichlet(b, dofs, uex[dofs], symmetric=True, x=x)
    info = A.solve(b, x, "cg", rtol=1e-13)
    assert info["converged"] == 1
    assert fo.relative_l2(x.numpy(), uex) < SOL_TOL


def test_errors_are_reported(ctx):
    c, t = fo.unit_cube_mesh(2, 2, 2)
    m = _lib.DeviceMesh.upload(ctx, c, t)
    A = _lib.DeviceMatrix.create(m, 1)
    bad = _lib.DeviceVector(ctx, 5)
    with pytest.raises(_lib.SolverError):
        A.spmv(bad, bad)
    with pytest.raises(_lib.SolverError):
        A.assemble_elasticity(1.0, 1.0)
    with pytest.raises(_lib.SolverError):
        _lib.DeviceMesh.box(ctx, (0, 2, 2), (0, 0, 0), (1, 1, 1))
    with pytest.raises(_lib.SolverError):
        ctx.set_option("no_such_option", 1)


@pytest.mark.parametrize("method", ["cg", "bicgstab"])
def test_drop_zeros_operand_gives_the_same_solution(ctx, method):
    """'drop_zeros': the Krylov SpMVs run on a compacted copy without the exactly-zero entries (8 of 15 per interior
    row on the right-angled cube, SURVEY 8c KAT 3, plus the symmetric Dirichlet eliminations).  Same solution, same
    iteration count to within rounding; the assembled CSR (the parity object) is untouched."""
    N = 12
    c, t, z0, z1 = heat_problem(N)
    nv = c.shape[0]
    m = _lib.DeviceMesh.box(ctx, (N, N, N), (0, 0, 0), (1, 1, 1))
    dofs = np.concatenate([z0, z1])
    vals = np.concatenate([np.full(z0.size, 350.0), np.full(z1.size, 300.0)])
    out = {}
    for dz in (0, 1):
        A = _lib.DeviceMatrix.create(m, 1)
        if method == "cg":
            A.assemble_scalar(kscale=20.0)
        else:
            A.assemble_scalar(kscale=20.0, adv=1.0, vel=np.array([3.0, -2.0, 1.0]))
        b = _lib.DeviceVector(ctx, nv)
        _lib.assemble_source(m, b, 1000.0)
        x = _lib.DeviceVector(ctx, nv)
        A.apply_dirichlet(b, dofs, vals, symmetric=(method == "cg"), x=x)
        ctx.set_option("drop_zeros", dz)
        try:
            info = A.solve(b, x, method, rtol=1e-12)
        finally:
            ctx.set_option("drop_zeros", 0)
        assert info["converged"] == 1
        out[dz] = (x.numpy(), info, A.download_csr())
    full, sq = out[0], out[1]
    nnz = full[2][0][-1]
    assert full[1]["operand_nnzb"] == nnz
    if method == "cg":
        nz_expected = int(np.count_nonzero(full[2][2]))          # every diagonal entry is non-zero here
        assert sq[1]["operand_nnzb"] == nz_expected
        assert sq[1]["operand_nnzb"] < 0.8 * nnz
    else:
        assert sq[1]["operand_nnzb"] < nnz
    assert fo.relative_l2(sq[0], full[0]) < 1e-11
    assert abs(sq[1]["iterations"] - full[1]["iterations"]) <= 2
    for a, b_ in zip(full[2], sq[2]):                      # the assembled matrix is the same object either way
        assert np.array_equal(a, b_) or np.abs(a - b_).max() <= 1e-13 * np.abs(a).max()      # atomic summation order


def test_drop_zeros_block_matrix(ctx):
    c, t = fo.box_mesh((0, 0, 0), (2, 1, 1), 8, 4, 4)
    nv = c.shape[0]
    mu, lam = fo.lame(2e11, 0.27)
    m = _lib.DeviceMesh.upload(ctx, c, t)
    fixed = np.flatnonzero(c[:, 0] == 0)
    dofs = (fixed[:, None] * 3 + np.arange(3)).ravel()
    sols = []
    for dz in (0, 1):
        A = _lib.DeviceMatrix.create(m, 3)
        A.assemble_elasticity(mu, lam)
        b = _lib.DeviceVector(ctx, 3 * nv)
        _lib.assemble_source(m, b, (0.0, 0.0, -7800 * 9.81), ncomp=3)
        x = _lib.DeviceVector(ctx, 3 * nv)
        A.apply_dirichlet(b, dofs, 0.0, symmetric=True, x=x)
        ctx.set_option("drop_zeros", dz)
        try:
            info = A.solve(b, x, "cg", rtol=1e-12, maxit=100000)
        finally:
            ctx.set_option("drop_zeros", 0)
        assert info["converged"] == 1
        sols.append((x.numpy(), info))
    assert sols[1][1]["operand_nnzb"] <= sols[0][1]["operand_nnzb"]
    assert fo.relative_l2(sols[1][0], sols[0][0]) < 1e-9


@pytest.mark.parametrize("drop_zeros", [0, 1])
def test_single_reduction_cg_matches_classic(ctx, drop_zeros):
    """cg_variant 2 (Chronopoulos-Gear recurrences, two kernels and one synchronisation point per iteration; the
    distributed default) against the classic three-kernel chain: same solution, iteration count within rounding,
    maxit honoured, zero right-hand side / converged start handled."""
    N = 14
    c, t, z0, z1 = heat_problem(N, jit=True)
    nv = c.shape[0]
    m = _lib.DeviceMesh.upload(ctx, c, t)
    dofs = np.concatenate([z0, z1])
    vals = np.concatenate([np.full(z0.size, 350.0), np.full(z1.size, 300.0)])
    A = _lib.DeviceMatrix.create(m, 1)
    A.assemble_scalar(kscale=20.0)
    b = _lib.DeviceVector(ctx, nv)
    _lib.assemble_source(m, b, 1000.0)
    x = _lib.DeviceVector(ctx, nv)
    A.apply_dirichlet(b, dofs, vals, symmetric=True, x=x)
    x0 = x.numpy()
    res = {}
    ctx.set_option("drop_zeros", drop_zeros)
    try:
        for variant in (1, 2):
            ctx.set_option("cg_variant", variant)
            xv = _lib.DeviceVector.from_numpy(ctx, x0)
            info = A.solve(b, xv, "cg", rtol=1e-12)
            assert info["converged"] == 1 and info["rnorm"] <= 1e-12 * info["bnorm"]
            res[variant] = (xv.numpy(), info["iterations"])
        assert fo.relative_l2(res[2][0], res[1][0]) < 1e-11
        assert abs(res[2][1] - res[1][1]) <= 2
        # maxit: exactly that many updates, outcome 0
        ctx.set_option("cg_variant", 2)
        for maxit in (5, 32, 64):
            xv = _lib.DeviceVector.from_numpy(ctx, x0)
            info = A.solve(b, xv, "cg", rtol=1e-30, maxit=maxit)
            assert info["converged"] == 0 and info["iterations"] == maxit
            ctx.set_option("cg_variant", 1)
            xc = _lib.DeviceVector.from_numpy(ctx, x0)
            A.solve(b, xc, "cg", rtol=1e-30, maxit=maxit)
            ctx.set_option("cg_variant", 2)
            assert fo.relative_l2(xv.numpy(), xc.numpy()) < 1e-9
        # start vector already converged: zero iterations, x untouched
        xs = _lib.DeviceVector.from_numpy(ctx, res[1][0])
        info = A.solve(b, xs, "cg", rtol=1e-9)
        assert info["converged"] == 1 and info["iterations"] == 0
        assert np.array_equal(xs.numpy(), res[1][0])
    finally:
        ctx.set_option("cg_variant", 0)
        ctx.set_option("drop_zeros", 0)


def _heat_system_on_device(ctx, N, dim=3, jit=False):
    if dim == 3:
        c, t, z0, z1 = heat_problem(N, jit=jit)
    else:
        c, t = fo.unit_square_mesh(N, N)
        z0 = np.nonzero(c[:, 1] == 0)[0]
        z1 = np.nonzero(c[:, 1] == 1)[0]
    nv = c.shape[0]
    m = _lib.DeviceMesh.upload(ctx, c, t)
    dofs = np.concatenate([z0, z1])
    vals = np.concatenate([np.full(z0.size, 350.0), np.full(z1.size, 300.0)])
    A = _lib.DeviceMatrix.create(m, 1)
    A.assemble_scalar(kscale=20.0)
    b = _lib.DeviceVector(ctx, nv)
    _lib.assemble_source(m, b, 1000.0)
    x = _lib.DeviceVector(ctx, nv)
    A.apply_dirichlet(b, dofs, vals, symmetric=True, x=x)
    return m, A, b, x.numpy()


@pytest.mark.parametrize("case", ["cube6", "cube20_jitter", "cube48", "square40"])
def test_persistent_cg_kernel_matches_kernel_chains(ctx, case):
    """cg_variant 3 — the whole iteration loop in one cooperative launch (fsb_cgp.cu: worker CTAs + a service CTA, grid
    barrier on an arrival counter, reductions through the mailbox) — against the classic chain (1) and the two-kernel
    single-reduction chain (2) it shares its recurrences with: same solution, same iteration count as (2), maxit honoured,
    converged start untouched, bitwise reproducible.  cube6 has fewer tiles than SMs (few workers); cube48 fills the grid;
    square40 has short rows (one lane per row)."""
    dim = 2 if case.startswith("square") else 3
    N = int("".join(ch for ch in case.split("_")[0] if ch.isdigit()))
    m, A, b, x0 = _heat_system_on_device(ctx, N, dim=dim, jit=case.endswith("jitter"))
    res = {}
    try:
        for variant in (1, 2, 3, 3):
            ctx.set_option("cg_variant", variant)
            xv = _lib.DeviceVector.from_numpy(ctx, x0)
            info = A.solve(b, xv, "cg", rtol=1e-12)
            assert info["converged"] == 1 and info["rnorm"] <= 1e-12 * info["bnorm"]
            res.setdefault(variant, []).append((xv.numpy(), info["iterations"]))
        assert fo.relative_l2(res[3][0][0], res[1][0][0]) < 1e-11
        assert fo.relative_l2(res[3][0][0], res[2][0][0]) < 1e-11
        assert abs(res[3][0][1] - res[2][0][1]) <= 1 and abs(res[3][0][1] - res[1][0][1]) <= 2
        assert res[3][0][1] == res[3][1][1] and np.array_equal(res[3][0][0], res[3][1][0])       # reproducible
        ctx.set_option("cg_variant", 3)
        for maxit in (1, 5, 33):
            xv = _lib.DeviceVector.from_numpy(ctx, x0)
            info = A.solve(b, xv, "cg", rtol=1e-30, maxit=maxit)
            assert info["converged"] == 0 and info["iterations"] == maxit
            ctx.set_option("cg_variant", 2)
            xc = _lib.DeviceVector.from_numpy(ctx, x0)
            A.solve(b, xc, "cg", rtol=1e-30, maxit=maxit)
            ctx.set_option("cg_variant", 3)
            assert fo.relative_l2(xv.numpy(), xc.numpy()) < 1e-7      # rounding differences grow with the iteration count
        xs = _lib.DeviceVector.from_numpy(ctx, res[1][0][0])
        info = A.solve(b, xs, "cg", rtol=1e-9)
        assert info["converged"] == 1 and info["iterations"] == 0
        assert np.array_equal(xs.numpy(), res[1][0][0])
        z = _lib.DeviceVector(ctx, x0.size)
        info = A.solve(z, _lib.DeviceVector(ctx, x0.size), "cg", rtol=1e-12)       # b = 0
        assert info["converged"] == 1 and info["iterations"] == 0
    finally:
        ctx.set_option("cg_variant", 0)


def test_persistent_cg_kernel_block3(ctx):
    """The 3x3-block instantiation of the persistent kernel on the elasticity cantilever, against the classic chain."""
    N = 10
    c, t = fo.box_mesh((0, 0, 0), (4, 1, 1), 2 * N, N, N)
    m = _lib.DeviceMesh.upload(ctx, c, t)
    A = _lib.DeviceMatrix.create(m, 3)
    E, nu = 2e11, 0.27
    A.assemble_elasticity(E / (2 * (1 + nu)), E * nu / ((1 + nu) * (1 - 2 * nu)))
    nv = c.shape[0]
    b = _lib.DeviceVector(ctx, 3 * nv)
    _lib.assemble_source(m, b, [0.0, 0.0, -7800 * 9.81], ncomp=3)
    fixed = np.nonzero(c[:, 0] == 0)[0]
    dofs = (3 * fixed[:, None] + np.arange(3)).ravel()
    x = _lib.DeviceVector(ctx, 3 * nv)
    A.apply_dirichlet(b, dofs, np.zeros(dofs.size), symmetric=True, x=x)
    res = {}
    try:
        for variant in (1, 3):
            ctx.set_option("cg_variant", variant)
            xv = _lib.DeviceVector(ctx, 3 * nv)
            info = A.solve(b, xv, "cg", rtol=1e-12, maxit=20000)
            assert info["converged"] == 1
            res[variant] = (xv.numpy(), info["iterations"])
        assert fo.relative_l2(res[3][0], res[1][0]) < 1e-9
        assert abs(res[3][1] - res[1][1]) <= max(3, res[1][1] // 50)
    finally:
        ctx.set_option("cg_variant", 0)


def test_pinned_result_pool_and_drop_zeros_auto(ctx):
    """DeviceVector.numpy() of a large vector lands in a page-locked block of the library's pool (reused once released); drop_zeros
    'auto' (2) compacts the Krylov operand of a right-angled box (exact zeros) and leaves an unstructured mesh's operand alone."""
    n = (1 << 20) + 5
    v = _lib.DeviceVector(ctx, n)
    v.fill(2.5)
    a = v.numpy()
    assert a.shape == (n,) and np.all(a == 2.5) and a.flags.writeable
    a[0] = 7.0                                                       # caller-owned memory
    addr = a.ctypes.data
    del a
    b = v.numpy()
    assert b.ctypes.data == addr and b[0] == 2.5                     # the released block served the next request of that size
    small = _lib.DeviceVector(ctx, 100)
    small.fill(1.0)
    assert np.all(small.numpy() == 1.0)
    c, t = fo.unit_cube_mesh(6, 6, 6)
    ctx.set_option("drop_zeros", 2)
    try:
        for coords, squeezed in ((c, True), (jitter(c, 6, seed=3), False)):
            m = _lib.DeviceMesh.upload(ctx, coords, t)
            A = _lib.DeviceMatrix.create(m, 1)
            A.assemble_scalar(kscale=1.0)
            nv = coords.shape[0]
            bvec, x = _lib.DeviceVector(ctx, nv), _lib.DeviceVector(ctx, nv)
            bvec.fill(1.0)
            d = np.nonzero(c[:, 2] == 0)[0]                          # the same vertices on both meshes (the jitter moves the boundary too)
            A.apply_dirichlet(bvec, d, np.zeros(d.size), True, x)
            info = A.solve(bvec, x, "cg", rtol=1e-12, maxit=1000)
            assert info["converged"] == 1
            assert (info["operand_nnzb"] < A.sizes()["nnz"]) == squeezed
    finally:
        ctx.set_option("drop_zeros", 0)
