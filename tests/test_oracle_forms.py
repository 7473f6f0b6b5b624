"""Known-answer tests that pin the oracle's restatement of thermal stress, the von Mises projection, the radiation
boundary term with its Newton solve, and point sources (the reference asserts no numbers for them; these are
closed forms), plus the host-side pieces around them (PointSource weights, .vtu/.pvd writer).  CPU only."""
import os
import xml.etree.ElementTree as ET

import numpy as np
import pytest
from scipy.optimize import brentq

from oracle import fem_oracle as fo
from oracle import fem_oracle_p2 as fp


def jitter(c, n, seed=0, amp=0.2):
    rng = np.random.default_rng(seed)
    return c + amp / n * (rng.random(c.shape) * 2 - 1)


@pytest.mark.parametrize("dim", [2, 3])
def test_thermal_load_divergence_theorem(dim):
    """sum_a x_a[i] b[(a,i)] = beta int (T - T_ref) dx (take v = x_i e_i, div v = 1), and a uniform temperature
    loads only the boundary: the entries of every interior node vanish."""
    n = 4
    c, t = fo.unit_square_mesh(n, n) if dim == 2 else fo.unit_cube_mesh(n, n, n)
    interior = np.all((c > 0) & (c < 1), axis=1)
    cj = jitter(c, n)
    cj[~interior] = c[~interior]
    beta = 3.0
    b = fo.thermal_load(cj, t, beta, 343.0, 293.0).reshape(-1, dim)
    assert np.abs(b[interior]).max() < 1e-12 * beta * 50
    assert np.allclose((cj * b).sum(axis=0), beta * 50.0, rtol=1e-12)
    T = 293.0 + 100 * cj[:, 1]
    b = fo.thermal_load(cj, t, beta, T, 293.0).reshape(-1, dim)
    vol, _ = fo.p1_geometry(cj, t)
    assert np.allclose((cj * b).sum(axis=0), beta * np.sum(vol * (T[t] - 293).mean(axis=1)), rtol=1e-12)
    # degree 2: quadratic temperature, integrated exactly
    cn, xn, _ = fp.p2_dofmap(cj, t)
    T2 = 293.0 + 100 * xn[:, 1] ** 2
    b2 = fp.thermal_load(cj, t, cn, xn.shape[0], beta, T2, 293.0).reshape(-1, dim)
    _, M, _, _ = fp.reference_tensors(dim)
    integral = beta * float(np.sum(vol[:, None] * np.einsum("ij,cj->ci", M, (T2 - 293.0)[cn])))
    assert np.allclose((xn * b2).sum(axis=0), integral, rtol=1e-11)
    # the P2 load of a uniform temperature agrees with the P1 one in total force on any node patch (here: all)
    b2c = fp.thermal_load(cj, t, cn, xn.shape[0], beta, 343.0, 293.0).reshape(-1, dim)
    assert np.allclose((xn * b2c).sum(axis=0), beta * 50.0, rtol=1e-12)


def test_free_thermal_expansion_is_stress_free():
    """A body heated uniformly and held only against rigid motion expands by eps = beta/(3 lambda + 2 mu) dT without
    stress: u = eps x is the exact discrete solution (it is linear), and its von Mises stress is zero."""
    n = 3
    c, t = fo.unit_cube_mesh(n, n, n)
    c = jitter(c, n, 3)
    nv = c.shape[0]
    E, nu, tec = 2e11, 0.27, 2e-6
    mu, lam = fo.lame(E, nu)
    beta = E / (1 - 2 * nu) * tec
    A = fo.assemble_matrix(t, fo.local_elasticity(c, t, mu, lam), nv, 3)
    b = fo.thermal_load(c, t, beta, 343.0, 293.0)
    eps = beta * 50.0 / (3 * lam + 2 * mu)
    assert abs(eps - tec * 50.0) < 1e-18                        # E/(1-2nu) = 3 lambda + 2 mu
    uex = (eps * c).ravel()
    assert np.abs(A @ uex - b).max() < 1e-9 * np.abs(b).max()
    vm = fo.von_mises_cells(c, t, uex, mu, lam) - 0.0
    sig_t = beta * 50.0
    # sigma(u) = (3 lambda + 2 mu) eps I is hydrostatic: no deviator
    assert vm.max() < 1e-9 * sig_t


def test_von_mises_uniaxial_tension_and_projection():
    """Uniaxial stress: u = eps (x, -nu y, -nu z) gives sigma = diag(E eps, 0, 0) and von Mises = E eps exactly;
    the P1 projection of a constant is that constant."""
    c, t = fo.unit_cube_mesh(3, 2, 2)
    c = jitter(c, 3, 4)
    E, nu, eps = 2e11, 0.27, 1e-4
    mu, lam = fo.lame(E, nu)
    u = eps * np.stack([c[:, 0], -nu * c[:, 1], -nu * c[:, 2]], axis=1)
    vm = fo.von_mises_cells(c, t, u, mu, lam)
    assert np.allclose(vm, E * eps, rtol=1e-11)
    assert np.allclose(fo.von_mises_projection(c, t, u, mu, lam), E * eps, rtol=1e-10)
    cn, xn, _ = fp.p2_dofmap(c, t)
    u2 = eps * np.stack([xn[:, 0], -nu * xn[:, 1], -nu * xn[:, 2]], axis=1)
    vol, _ = fo.p1_geometry(c, t)
    assert np.allclose(fp.von_mises_load(c, t, cn, u2, mu, lam).sum(), E * eps * vol.sum(), rtol=1e-11)
    # pure shear in 2D with the reference's 1/3 (not 1/2) deviator: sigma_xy = mu g, vm = sqrt(3) mu g
    c2, t2 = fo.unit_square_mesh(3, 3)
    g = 1e-3
    us = np.stack([g * c2[:, 1], np.zeros(c2.shape[0])], axis=1)
    assert np.allclose(fo.von_mises_cells(c2, t2, us, mu, lam), np.sqrt(3.0) * mu * g, rtol=1e-12)


@pytest.mark.parametrize("dim", [2, 3])
def test_radiation_terms_exact_integration_and_jacobian(dim):
    n = 3
    c, t = fo.unit_square_mesh(n, n) if dim == 2 else fo.unit_cube_mesh(n, n, n)
    c = jitter(c, n, 7)
    fv, _, _ = fo.exterior_facets(t)
    T = 300 + 50 * c[:, -1] + 10 * c[:, 0] ** 2
    m, Ta = 0.9 * 5.670367e-8, 280.0
    J, r = fo.radiation_terms(c, fv, T, m, Ta)
    # against a high-order Gauss rule on the facets
    if dim == 3:
        pts, w = fp._collapsed_rule(2, 6)
    else:
        x, w = np.polynomial.legendre.leggauss(6)
        pts, w = np.stack([0.5 * (1 - x), 0.5 * (1 + x)], axis=1), 0.5 * w
    meas = fo.facet_measure(c, fv)
    Tq = T[fv] @ pts.T
    rq = m * np.einsum("fp,p,pa->fa", Tq ** 4 - Ta ** 4, w, pts) * meas[:, None]
    Jq = 4 * m * np.einsum("fp,p,pa,pb->fab", Tq ** 3, w, pts, pts) * meas[:, None, None]
    assert np.abs(r - rq).max() < 1e-13 * np.abs(r).max() and np.abs(J - Jq).max() < 1e-13 * np.abs(J).max()
    # total radiated power of an isothermal body: m (T^4 - Ta^4) * area
    _, r1 = fo.radiation_terms(c, fv, np.full(c.shape[0], 350.0), m, Ta)
    assert np.allclose(r1.sum(), m * (350.0 ** 4 - Ta ** 4) * meas.sum(), rtol=1e-13)


def test_radiation_one_dimensional_balance_known_answer():
    """Bar heated to 360 K at y = 1 that radiates to 280 K from y = 0 only: the profile is linear (nodally exact for
    P1) and the end temperature solves k (360 - Tb) = m (Tb^4 - Ta^4).  Newton converges quadratically."""
    n = 6
    c, t = fo.unit_square_mesh(n, n)
    fv, _, _ = fo.exterior_facets(t)
    bottom = fv[np.all(c[fv][:, :, 1] == 0, axis=1)]
    top = np.nonzero(c[:, 1] == 1)[0]
    k, m, Ta = 0.6, 0.9 * 5.670367e-8, 280.0
    T, its = fo.solve_radiation_newton(c, t, k, [(top, 360.0)], bottom, m, Ta, 300.0)
    Tb = brentq(lambda x: k * (360.0 - x) - m * (x ** 4 - Ta ** 4), 200.0, 360.0, xtol=1e-13)
    assert np.abs(T - (Tb + (360.0 - Tb) * c[:, 1])).max() < 1e-9
    assert its <= 6


def test_point_source_weights_and_vtu_writer(tmp_path):
    from fenicssolver_b200 import SolverBase as sb
    from fenicssolver_b200.dolfin_compat import FunctionSpace, Point, PointSource, UnitCubeMesh, UnitSquareMesh
    mesh = UnitSquareMesh(4, 4)
    for degree in (1, 2):
        V = FunctionSpace(mesh, "CG", degree)
        nodes, w = PointSource(V, Point(0.3, 0.6), 2.5).entries()
        assert abs(w.sum() - 2.5) < 1e-14                       # partition of unity
        x = V.node_coordinates()[nodes]
        if degree == 1:
            assert np.allclose((w[:, None] * x).sum(axis=0), 2.5 * np.array([0.3, 0.6]))     # linear reproduction
        else:
            assert np.allclose((w * x[:, 0] * x[:, 1]).sum(), 2.5 * 0.18)                   # quadratic reproduction
    with pytest.raises(sb.SolverError):
        mesh.locate_point(Point(1.5, 0.5))
    m3 = UnitCubeMesh(2, 2, 2)
    vals = np.arange(27.0)
    p = os.path.join(str(tmp_path), "T.vtu")
    sb.write_vtu(p, m3, vals, "temperature")
    piece = ET.parse(p).getroot().find("./UnstructuredGrid/Piece")
    assert piece.attrib == {"NumberOfPoints": "27", "NumberOfCells": "48"}
    conn = np.array(piece.find("./Cells/DataArray[@Name='connectivity']").text.split(), dtype=int).reshape(-1, 4)
    assert np.array_equal(conn, m3.cells())
    assert set(piece.find("./Cells/DataArray[@Name='types']").text.split()) == {"10"}
    got = np.array(piece.find("./PointData/DataArray").text.split(), dtype=float)
    assert np.array_equal(got, vals)
    sb.write_pvd(os.path.join(str(tmp_path), "T.pvd"), [(0.0, "T000000.vtu"), (0.5, "T000001.vtu")])
    ds = ET.parse(os.path.join(str(tmp_path), "T.pvd")).getroot().findall("./Collection/DataSet")
    assert [d.attrib["file"] for d in ds] == ["T000000.vtu", "T000001.vtu"] and ds[1].attrib["timestep"] == "0.5"


def test_forms_carry_thermal_stress_radiation_and_point_sources():
    """Host-side form generation (no device call): LinearElasticitySolver.py:230-238, ScalarTransportSolver.py:334-374."""
    import copy
    from fenicssolver_b200 import LinearElasticitySolver, ScalarTransportSolver, SolverBase
    from fenicssolver_b200.dolfin_compat import (AutoSubDomain, BoxMesh, Constant, Expression, FunctionSpace, Point, UnitSquareMesh,
                                                 VectorFunctionSpace, near)
    quiet = {'logging_level': 40, 'logging_file': None, 'plotting_freq': 0, 'saving_freq': 0}
    mesh = BoxMesh(Point(0, 0, 0), Point(10, 1, 1), 4, 2, 2)
    s = copy.deepcopy(SolverBase.default_case_settings)
    s.update({'material': {'elastic_modulus': 2e11, 'poisson_ratio': 0.27, 'density': 7800, 'thermal_expansion_coefficient': 2e-6},
              'function_space': VectorFunctionSpace(mesh, "Lagrange", 2), 'report_settings': quiet,
              'boundary_conditions': {'fixed': {'boundary': AutoSubDomain(lambda x: near(x[0], 0.0)), 'boundary_id': 1, 'type': 'Dirichlet',
                                                'value': Constant((0, 0, 0))}},
              'temperature_distribution': Expression("343", degree=2)})
    s['solver_settings']['reference_values'] = {'temperature': 293}
    solver = LinearElasticitySolver.LinearElasticitySolver(s)
    solver.init_solver()
    solver.current_step = 0
    F, bcs = solver.generate_form(0, None, None, solver.w_current, solver.w_prev)
    beta, T, Tref = F.thermal
    assert abs(beta - 2e11 / (1 - 0.54) * 2e-6) < 1e-6 and Tref == 293.0
    assert T.shape == (solver.function_space.num_nodes(),) and np.all(T == 343.0)
    s['temperature_distribution'] = None
    solver = LinearElasticitySolver.LinearElasticitySolver(s)
    solver.init_solver()
    solver.current_step = 0
    assert solver.generate_form(0, None, None, solver.w_current, solver.w_prev)[0].thermal is None

    Q = FunctionSpace(UnitSquareMesh(4, 4), "CG", 1)
    hs = {'solver_name': 'ScalarTransportSolver', 'scalar_name': 'temperature', 'mesh': None, 'function_space': Q,
          'boundary_conditions': {'hot': {'boundary': AutoSubDomain(lambda x: near(x[1], 1.0)), 'boundary_id': 1, 'type': 'Dirichlet', 'value': 360}},
          'body_source': None, 'initial_values': {'temperature': 300}, 'report_settings': quiet,
          'material': {'density': 1000, 'specific_heat_capacity': 4200, 'thermal_conductivity': 0.1},
          'solver_settings': {'transient_settings': {'transient': False, 'starting_time': 0, 'time_step': 0.1, 'ending_time': 1},
                              'reference_values': {'temperature': 300}, 'solver_parameters': {}},
          'radiation_settings': {'emissivity': 0.5}, 'point_source': [((0.3, 0.3), 2.0)]}
    hsolver = ScalarTransportSolver.ScalarTransportSolver(hs)
    hsolver.init_solver()
    hsolver.current_step = 0
    F, bcs = hsolver.generate_form(0, None, None, hsolver.w_current, hsolver.w_prev)
    assert hsolver.nonlinear and F.radiation == (0.5 * 5.670367e-8, 300.0)       # ambient falls back to the reference value
    hsolver.material['emissivity'] = 0.9                                        # material wins over radiation_settings
    assert hsolver.radiation_coefficients()[0] == 0.9 * 5.670367e-8
    assert len(F.point_sources) == 1 and F.point_sources[0].magnitude == 2.0


def test_nonlinear_conductivity_kirchhoff_known_answer():
    """k(T) = k0 (1 + beta (T - 300)) between T = 360 at y = 1 and T = 300 at y = 0: the Kirchhoff variable
    theta = int k dT is linear in y, so T(y) solves k0 (s + beta s^2 / 2) = y * k0 (60 + beta 60^2 / 2), s = T - 300.
    The P1 solution (k interpolated at the nodes) converges to it at second order; the Jacobian is the derivative
    of the residual."""
    k0, beta = 0.6, 0.02
    kf, dkf = (lambda T: k0 * (1 + beta * (T - 300.0))), (lambda T: k0 * beta + 0 * T)
    errs = []
    for n in (4, 8, 16):
        c, t = fo.unit_square_mesh(n, n)
        top, bot = np.nonzero(c[:, 1] == 1)[0], np.nonzero(c[:, 1] == 0)[0]
        T, its = fo.solve_nonlinear_k_newton(c, t, kf, dkf, [(top, 360.0), (bot, 300.0)], 300.0)
        q = c[:, 1] * (60 + beta * 1800)
        s = (-1 + np.sqrt(1 + 2 * beta * q)) / beta
        errs.append(np.abs(T - 300 - s).max())
        assert its <= 8
    assert errs[0] < 0.2 and errs[1] < errs[0] / 3 and errs[2] < errs[1] / 3.5
    c, t = fo.unit_square_mesh(4, 4)
    rng = np.random.default_rng(0)
    T = 300 + 60 * rng.random(c.shape[0])
    J, R = fo.nonlinear_k_terms(c, t, T, kf, dkf)
    d = rng.standard_normal(c.shape[0])
    eps = 1e-4
    fd = (fo.nonlinear_k_terms(c, t, T + eps * d, kf, dkf)[1] - fo.nonlinear_k_terms(c, t, T - eps * d, kf, dkf)[1]) / (2 * eps)
    assert np.abs(J @ d - fd).max() < 1e-7 * np.abs(fd).max()
    from fenicssolver_b200.ScalarTransportSolver import nodal_function_and_derivative
    k, dk = nodal_function_and_derivative(lambda T: (T - 300) / 300 * 0.6 + 0.1 * T ** 2, T)
    assert np.allclose(dk, 0.6 / 300 + 0.2 * T, rtol=1e-14) and np.allclose(k, (T - 300) / 300 * 0.6 + 0.1 * T ** 2)
    k, dk = nodal_function_and_derivative(lambda T: np.where(np.real(T) > 330, 1.0, 0.5) * np.abs(T), T)   # no complex step
    assert np.allclose(dk, np.where(T > 330, 1.0, 0.5), rtol=1e-6)


@pytest.mark.parametrize("dim", [2, 3])
def test_supg_weights_consistency_and_monotone_profile(dim):
    """SUPG (ScalarTransportSolver.py:252-274): the weights s_a = tau v.G_a sum to zero per cell; with Dirichlet data
    of a linear field T = a + g.x and the source S = c v.g, that field solves the stabilised system exactly
    (advection residual zero, diffusion of a linear field zero); in the advection-dominated 1-D layer problem the
    Galerkin solution oscillates wildly and the SUPG one nearly not at all."""
    n = 4
    c, t = fo.unit_square_mesh(n, n) if dim == 2 else fo.unit_cube_mesh(n, n, n)
    bnd = np.any((c == 0) | (c == 1), axis=1)
    cj = jitter(c, n, 11)
    cj[bnd] = c[bnd]
    vel = np.array([0.7, -0.3, 0.2][:dim])
    Pe, cap, k = 5.0, 3.0, 0.05
    s = fo.supg_weights(cj, t, vel, Pe)
    assert np.abs(s.sum(axis=1)).max() < 1e-14
    h = 2 * fo.circumradius(cj, t)
    X = cj[t]
    assert np.allclose(np.linalg.norm(X - (X[:, :1] + 0), axis=2).max(axis=1) <= h + 1e-12, True)     # every vertex within the circumsphere diameter
    nv = cj.shape[0]
    g = np.array([2.0, -1.0, 0.5][:dim])
    Tlin = 300 + cj @ g
    A = fo.assemble_matrix(t, fo.local_laplace(cj, t, k) + fo.local_advection(cj, t, vel, cap) + fo.local_supg(cj, t, vel, Pe, adv=cap), nv)
    S = cap * float(vel @ g)
    b = fo.assemble_source(cj, t, S) + fo.supg_source(cj, t, S, vel, Pe)
    dofs = np.nonzero(bnd)[0]
    Ab, bb = fo.apply_dirichlet(A, b, dofs, Tlin[dofs], symmetric=False)
    assert np.abs(fo.solve_direct(Ab, bb) - Tlin).max() < 1e-10
    if dim == 2:
        # boundary layer: T = 0 at y = 0, T = 1 at y = 1, v = (0, 1), cell Peclet = c |v| h / (2k) = 10
        n = 10
        c, t = fo.unit_square_mesh(n, n)
        nv = c.shape[0]
        v, kk = np.array([0.0, 1.0]), 0.005
        top, bot = np.nonzero(c[:, 1] == 1)[0], np.nonzero(c[:, 1] == 0)[0]
        dofs = np.concatenate([top, bot]); vals = np.concatenate([np.ones(top.size), np.zeros(bot.size)])
        sols = []
        for supg in (False, True):
            Ke = fo.local_laplace(c, t, kk) + fo.local_advection(c, t, v, 1.0)
            if supg:
                Ke = Ke + fo.local_supg(c, t, v, 1e3, adv=1.0)
            Ab, bb = fo.apply_dirichlet(fo.assemble_matrix(t, Ke, nv), np.zeros(nv), dofs, vals, symmetric=False)
            sols.append(fo.solve_direct(Ab, bb))
        # Galerkin: wild node-to-node oscillations; this tau is ~0.7 of the 1-D optimum, so a small undershoot remains
        tv = [np.abs(np.diff(s_.reshape(n + 1, n + 1)[:, 5])).sum() for s_ in sols]      # total variation along the flow
        assert sols[0].min() < -1.5 and tv[0] > 7
        assert sols[1].min() > -0.35 and sols[1].max() < 1 + 1e-10 and tv[1] < 1.3


@pytest.mark.parametrize("dim", [2, 3])
def test_box_meshes_are_nested_and_rediscretisation_is_galerkin(dim):
    """What the geometric multigrid preconditioner (csrc/fsb_mg.cu) rests on: dolfin's box triangulation with n/2 cells per
    axis is nested in the one with n cells, the fine vertex 2C + d being the midpoint of the coarse edge (C, C + d).  So
    the prolongation built from that rule reproduces P1 functions of the coarse mesh exactly, and the coarse stiffness,
    mass and elasticity matrices equal P^T A_fine P."""
    import scipy.sparse as sp
    n = (4, 6, 2)[:dim]
    nc = tuple(k // 2 for k in n)
    p1 = (1.0, 1.5, 0.5)[:dim]
    if dim == 2:
        cf, tf = fo.rectangle_mesh(0, 0, p1[0], p1[1], *n)
        cc, tc = fo.rectangle_mesh(0, 0, p1[0], p1[1], *nc)
    else:
        cf, tf = fo.box_mesh((0, 0, 0), p1, *n)
        cc, tc = fo.box_mesh((0, 0, 0), p1, *nc)
    dims_f = [k + 1 for k in n] + [1] * (3 - dim)
    dims_c = [k + 1 for k in nc] + [1] * (3 - dim)
    rows, cols, vals = [], [], []
    for f in range(cf.shape[0]):
        i, j, k = f % dims_f[0], (f // dims_f[0]) % dims_f[1], f // (dims_f[0] * dims_f[1])
        d = (i & 1, j & 1, k & 1)
        c0 = (i >> 1) + dims_c[0] * ((j >> 1) + dims_c[1] * (k >> 1))
        c1 = ((i >> 1) + d[0]) + dims_c[0] * (((j >> 1) + d[1]) + dims_c[1] * ((k >> 1) + d[2]))
        if d == (0, 0, 0):
            rows += [f]; cols += [c0]; vals += [1.0]
        else:
            rows += [f, f]; cols += [c0, c1]; vals += [0.5, 0.5]
    P = sp.csr_matrix((vals, (rows, cols)), shape=(cf.shape[0], cc.shape[0]))
    assert np.abs(P @ cc - cf).max() < 1e-15                     # coordinates are P1: interpolation reproduces them
    # every fine cell lies inside ONE coarse cell's vertex hull in the sense that matters: Galerkin = rediscretisation
    for local in (lambda c, t: fo.local_laplace(c, t, 2.0), lambda c, t: fo.local_mass(c, t, 3.0)):
        Af = fo.assemble_matrix(tf, local(cf, tf), cf.shape[0])
        Ac = fo.assemble_matrix(tc, local(cc, tc), cc.shape[0])
        G = (P.T @ Af @ P).toarray()
        assert np.abs(G - Ac.toarray()).max() < 1e-12 * np.abs(Ac.toarray()).max()
    mu, lam = fo.lame(2e11, 0.27)
    Pv = sp.kron(P, sp.identity(dim)).tocsr()
    Af = fo.assemble_matrix(tf, fo.local_elasticity(cf, tf, mu, lam), cf.shape[0], dim)
    Ac = fo.assemble_matrix(tc, fo.local_elasticity(cc, tc, mu, lam), cc.shape[0], dim)
    G = (Pv.T @ Af @ Pv).toarray()
    assert np.abs(G - Ac.toarray()).max() < 1e-11 * np.abs(Ac.toarray()).max()


@pytest.mark.parametrize("dim", [2, 3])
def test_stress_tensor_boundary_tractions(dim):
    """'stress' boundary with a tensor value, g = dot(g, mesh_normal) (LinearElasticitySolver.py:190-196): the host routine
    that turns it into right-hand-side entries, against the oracle's facet loads with the per-facet traction S.n; a
    hydrostatic tensor p I on the whole surface of a box is a closed surface integral of the normal: total force zero,
    and equals the pressure load."""
    from fenicssolver_b200.LinearElasticitySolver import facet_traction_entries
    from fenicssolver_b200.dolfin_compat import FunctionSpace, Mesh
    n = 3
    c, t = fo.unit_square_mesh(n, n + 1) if dim == 2 else fo.unit_cube_mesh(n, n + 1, n)
    c = jitter(c, n + 1, 5)
    mesh = Mesh(c, t)
    fv, opp, _ = fo.exterior_facets(t)
    S = np.arange(1.0, dim * dim + 1).reshape(dim, dim) * 1e5
    meas, nrm = fo.facet_measure(c, fv, opp)
    for degree in (1, 2):
        V = FunctionSpace(mesh, "CG", degree, ncomp=dim)
        dofs, vals = facet_traction_entries(mesh, V, fv, opp, S)
        b = np.zeros(V.num_nodes() * dim)
        np.add.at(b, dofs, vals)
        if degree == 1:
            ref = fo.assemble_facet_load(c, fv, nrm @ S.T, c.shape[0], dim)
        else:
            cn, xn, edges = fp.p2_dofmap(c, t)
            ref = fp.assemble_facet_load(c, fv, fp.facet_nodes(fv, edges, c.shape[0]), nrm @ S.T, xn.shape[0], dim)
        assert np.abs(b - ref).max() < 1e-12 * np.abs(ref).max()
        dofs, vals = facet_traction_entries(mesh, V, fv, opp, 2.5e5 * np.eye(dim))
        tot = np.zeros(dim)
        np.add.at(tot, dofs % dim, vals)
        assert np.abs(tot).max() < 1e-9 * 2.5e5


def test_c_multigrid_restatement_matches_numpy_and_the_exact_profile():
    """oracle/fem_oracle_c.c fo_mg_pcg / fo_mg_apply (the CPU figure of the bench's `gmg` block) against oracle/mg_oracle.py on the same
    level matrices, and config C2's exact 1-D profile; bench.py's helper functions around it."""
    import scipy.sparse as sp
    from oracle import c_oracle as co
    from oracle import mg_oracle as mo
    N = 16
    h = co.HeatCubeMG(N)
    r = h.step()
    assert r["levels"] == 4 and r["iterations"] <= 14 and r["relres"] < 1e-12
    z = (np.arange((N + 1) ** 3) // ((N + 1) ** 2)) / N
    assert fo.relative_l2(r["x"], 350 - 50 * z + 1000 * z * (1 - z) / 40) < 1e-11
    levels, transfers, nl = [], [], [N] * 3
    for l, cu in enumerate(h.cubes):
        M = sp.csr_matrix((cu.vals, cu.ci.astype(np.int64), cu.rp), shape=(cu.nv, cu.nv))
        levels.append({"A": M, "dinv": 1.0 / M.diagonal(), "bc": cu.flag.astype(bool), "omega": 4.0 / (3.0 * h.lmax[l])})
        if l + 1 < len(h.cubes):
            transfers.append(mo.prolongation(nl))
            nl = [k // 2 for k in nl]
    assert all(1.5 < lm <= 2.2 for lm in h.lmax)
    rng = np.random.default_rng(0)
    res = rng.standard_normal(h.cubes[0].nv)
    res[levels[0]["bc"]] = 0.0
    mg = co.MultigridLevels([{"rp": cu.rp, "ci": cu.ci, "va": cu.vals, "bc": cu.flag, "dims": (cu.N + 1,) * 3} for cu in h.cubes], h.lmax)
    zo = mo.vcycle(levels, transfers, res)
    assert np.abs(mg.apply(res) - zo).max() < 1e-13 * np.abs(zo).max()
    import bench
    g = bench.cpu_heat_gmg(h)
    assert g["value"] > 0 and g["iterations"] == r["iterations"] and g["kind"] == "port"
    j = bench.cpu_heat_full(N, nseg=4, warmup=1, cube=h.cubes[0])
    one = h.cubes[0].step(rtol=bench.RTOL)
    assert j["value"] > 0 and j["converged"] == 1 and j["iterations"] == one["iterations"] and len(j["segments_s"]) == 4
    assert j["rel_l2_vs_exact"] < 1e-10 and abs(sum(j["segments_s"]) - j["total_s"]) < 1e-9


def test_oracle_regression_pins(golden_dir):
    """tests/golden/forms_expected.npz (made by tests/golden/make_forms_golden.py): the oracle's outputs for fixed small inputs, so an
    accidental change of the checker itself shows up."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_forms_golden", os.path.join(golden_dir, "make_forms_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    now = mod.compute()
    g = np.load(os.path.join(golden_dir, "forms_expected.npz"))
    assert set(g.files) == set(now)
    for key in g.files:
        a, b = np.asarray(now[key], dtype=np.float64), np.asarray(g[key], dtype=np.float64)
        assert a.shape == b.shape, key
        scale = max(float(np.abs(b).max()), 1e-300)
        assert np.abs(a - b).max() <= 1e-12 * scale, key
