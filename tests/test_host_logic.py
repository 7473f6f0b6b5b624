"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/fsb.h declares,
settings parsing / boundary marking / value translation mirror the reference, the product path fails
loudly without a GPU, and the z-slab partition logic is consistent across a world_size-2 gloo group."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from fenicssolver_b200 import LinearElasticitySolver, ScalarTransportSolver, SolverBase, _lib, backend
from fenicssolver_b200.dolfin_compat import (AutoSubDomain, BoxMesh, Constant, Expression, FacetMarkers, FunctionSpace, Mesh, MeshFunction,
                                             Point, SubDomain, UnitCubeMesh, UnitSquareMesh, VectorFunctionSpace, near)
from fenicssolver_b200.main import load_settings, main
from oracle import fem_oracle as fo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    from fenicssolver_b200 import build
    lib_path = build.build()
    header = open(os.path.join(ROOT, "include", "fsb.h")).read()
    declared = set(re.findall(r"\b(fsb_[a-z0-9_]+)\s*\(", header))
    declared -= {"fsb_status"}
    assert len(declared) >= 40
    lib = _lib.load_library()
    for name in sorted(declared):
        assert hasattr(lib, name), "libfsb.so does not export %s" % name
        assert name in _lib.SIGNATURES, "no ctypes prototype for %s" % name
    assert set(_lib.SIGNATURES) == declared
    # the shared object is self-contained sm_100a code with TMA bulk copies in the SpMV kernel
    if os.environ.get("FSB_CHECK_SASS", "1") == "1":
        out = subprocess.run(["cuobjdump", "-sass", lib_path], capture_output=True, text=True)
        if out.returncode == 0:
            assert "sm_100a" in out.stdout and "UBLKCP" in out.stdout and "REDG.E.ADD.F64" in out.stdout


def test_no_gpu_means_loud_failure_not_fallback():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is visible")
    except ImportError:
        pass
    with pytest.raises(_lib.SolverError):
        _lib.Context(0)
    mesh = UnitSquareMesh(4, 4)
    s = heat_settings(mesh)
    solver = ScalarTransportSolver.ScalarTransportSolver(s)
    with pytest.raises(_lib.SolverError):
        solver.solve()
    # nothing in the product package imports the oracle
    for fn in os.listdir(os.path.join(ROOT, "fenicssolver_b200")):
        if fn.endswith(".py"):
            src = open(os.path.join(ROOT, "fenicssolver_b200", fn)).read()
            assert "oracle" not in src.replace("# oracle", ""), fn


def heat_settings(mesh, bcs=None):
    top = AutoSubDomain(lambda x: near(x[1], 1))
    bottom = AutoSubDomain(lambda x: near(x[1], 0))
    bcs = bcs or {"hot": {'boundary': top, 'boundary_id': 1, 'type': 'Dirichlet', 'value': Constant(360)},
                  "cold": {'boundary': bottom, 'boundary_id': 2, 'values': {'temperature': {'variable': 'temperature', 'type': 'HTC', 'value': Constant(100), 'ambient': Constant(300)}}}}
    return {'solver_name': 'ScalarEquationSolver', 'mesh': None, 'function_space': FunctionSpace(mesh, "CG", 1),
            'periodic_boundary': None, 'fe_degree': 1, 'boundary_conditions': bcs, 'body_source': None,
            'initial_values': {'temperature': 300},
            'material': {'density': 1000, 'specific_heat_capacity': 4200, 'thermal_conductivity': 0.1},
            'solver_settings': {'transient_settings': {'transient': False, 'starting_time': 0, 'time_step': 0.1, 'ending_time': 1},
                                'reference_values': {'temperature': 300},
                                'solver_parameters': {"relative_tolerance": 1e-9, "maximum_iterations": 500, "monitor_convergence": True}},
            'scalar_name': 'temperature'}


def test_box_meshes_match_dolfin_layout():
    for n in [(5, 3), (4, 3, 2)]:
        if len(n) == 2:
            m, (c, t) = UnitSquareMesh(*n), fo.unit_square_mesh(*n)
        else:
            m, (c, t) = BoxMesh(Point(0, 0, 0), Point(10, 1, 1), *n), fo.box_mesh((0, 0, 0), (10, 1, 1), *n)
        assert np.array_equal(m.coordinates(), c) and np.array_equal(m.cells(), t)
        fv, opp = m.exterior_facets()
        f2, o2, _ = fo.exterior_facets(t)
        assert set(map(tuple, np.hstack([fv, opp[:, None]]))) == set(map(tuple, np.hstack([f2, o2[:, None]])))
        assert m.num_vertices() == c.shape[0] and m.num_cells() == t.shape[0] and m.geometry().dim() == len(n)


def test_subdomain_marking_semantics():
    m = UnitCubeMesh(4, 4, 4)
    fm = FacetMarkers(m)
    fm.set_all(0)
    fm.mark_subdomain(AutoSubDomain(lambda x: near(x[2], 1.0)), 2)
    assert (fm.values == 2).sum() == 32
    fm.mark_subdomain(AutoSubDomain(lambda x, on_boundary: on_boundary and near(x[0], 0.0)), 3)
    assert (fm.values == 3).sum() == 32 and (fm.values == 2).sum() == 32      # disjoint facets: nothing overwritten

    class Corner(SubDomain):           # python `and` on arrays -> falls back to the per-point loop
        def inside(self, x, on_boundary):
            return bool(near(x[0], 0)) and bool(near(x[1], 0))
    fm.mark_subdomain(Corner(), 5)
    assert (fm.values == 5).sum() == 0                                        # an edge holds no whole facet
    fm.mark_subdomain(lambda x: x[2] > 0.5 - 1e-12, 7)                        # later ids overwrite earlier ones
    assert (fm.values == 2).sum() == 0 and (fm.values == 7).sum() >= 32
    assert np.array_equal(fm.vertices(3), np.unique(fm.facets(3)[0]))


def test_xml_readers_and_marker_files(tmp_path, golden_dir):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_solvers import write_dolfin_xml
    g = np.load(os.path.join(golden_dir, "fixture_mesh.npz"))
    path = write_dolfin_xml(str(tmp_path), g)
    mesh = Mesh(path)
    assert np.array_equal(mesh.coordinates(), g["coords"]) and np.array_equal(mesh.cells(), g["cells"])
    cm = MeshFunction("size_t", mesh, path[:-4] + "_physical_region.xml")
    assert np.all(cm.array() == 3) and cm.array().size == 4355
    with pytest.raises(SolverBase.SolverError):             # the boundary search of a file mesh is device work: loud without a GPU
        MeshFunction("size_t", mesh, path[:-4] + "_facet_region.xml")


@pytest.mark.gpu
def test_marker_files_and_solver_from_xml(tmp_path, golden_dir):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_solvers import write_dolfin_xml
    g = np.load(os.path.join(golden_dir, "fixture_mesh.npz"))
    path = write_dolfin_xml(str(tmp_path), g)
    mesh = Mesh(path)
    fm = MeshFunction("size_t", mesh, path[:-4] + "_facet_region.xml")
    assert (fm.values == 1).sum() == 100 and (fm.values == 2).sum() == 100 and fm.values.size == 1400
    assert np.all(mesh.coordinates()[fm.vertices(1), 2] == 0.0) and np.all(mesh.coordinates()[fm.vertices(2), 2] == 20.0)
    cm = MeshFunction("size_t", mesh, path[:-4] + "_physical_region.xml")
    assert np.all(cm.array() == 3) and cm.array().size == 4355
    settings = json.load(open(os.path.join(golden_dir, "TestHeatTransfer.json")))
    settings["mesh"] = path
    solver = ScalarTransportSolver.ScalarTransportSolver(load_settings(settings))
    assert solver.dimension == 3 and solver.function_space.dim() == 1069
    assert solver.conductivity() == 20 and solver.capacity() == 1000 * 500
    kp = solver.krylov_parameters()
    assert kp["rtol"] == 1e-12 and kp["maxit"] >= 100000                       # parity mode tightens the JSON's 1e-7 / 500
    solver.solver_settings['solver_parameters']['parity_mode'] = False
    kp = solver.krylov_parameters()
    assert kp["rtol"] == 1e-7 and kp["maxit"] == 500


def test_settings_errors_mirror_reference():
    with pytest.raises(SolverBase.SolverError):
        ScalarTransportSolver.ScalarTransportSolver("case.json")                # must be a dict (SolverBase.py:96-101)
    with pytest.raises(SolverBase.SolverError):
        ScalarTransportSolver.ScalarTransportSolver({'boundary_conditions': {}, 'solver_settings': {}})   # no mesh / space
    with pytest.raises(TypeError):
        load_settings(["not", "a", "path"])                                     # main.py:74-75
    with pytest.raises(NameError):
        main({'solver_name': 'NoSuchSolver'})                                   # main.py:92-93
    with pytest.raises(SolverBase.SolverError):
        FunctionSpace(UnitSquareMesh(2, 2), "CG", 3)                            # degrees 1 and 2 exist: loud, not silent P1
    m = UnitSquareMesh(2, 2)
    s = heat_settings(m)
    s['mesh'] = "/no/such/mesh.xml"
    with pytest.raises(SolverBase.SolverError):
        ScalarTransportSolver.ScalarTransportSolver(s)


def test_form_generation_matches_reference_rules():
    mesh = UnitSquareMesh(6, 6)
    solver = ScalarTransportSolver.ScalarTransportSolver(heat_settings(mesh))
    solver.init_solver()
    solver.current_step = 0
    F, bcs = solver.generate_form(0, None, None, solver.w_current, solver.w_prev)
    assert F.conductivity == 0.1 and F.capacity == 4200000 and F.robin == [(2, 100.0, 300.0)] and F.neumann == []
    d, v = SolverBase.collect_dirichlet(bcs, mesh)
    assert d.size == 7 and np.all(v == 360) and np.all(mesh.coordinates()[d, 1] == 1)
    # Neumann gradient is scaled by the capacity, flux is not (ScalarTransportSolver.py:176-200)
    left = AutoSubDomain(lambda x: near(x[0], 0))
    bc2 = {"a": {'boundary': left, 'boundary_id': 3, 'type': 'Neumann', 'value': 2.0},
           "b": {'boundary': AutoSubDomain(lambda x: near(x[0], 1)), 'boundary_id': 4, 'type': 'heatFlux', 'value': Constant(5.0)},
           "c": {'boundary': AutoSubDomain(lambda x: near(x[1], 1)), 'boundary_id': 1, 'type': 'Dirichlet', 'value': "300 + x[0]"}}
    s = heat_settings(mesh, bc2)
    s['convective_velocity'] = Constant((0.5, -0.5))
    solver = ScalarTransportSolver.ScalarTransportSolver(s)
    solver.init_solver()
    solver.current_step = 0
    F, bcs = solver.generate_form(0, None, None, solver.w_current, solver.w_prev)
    assert F.neumann == [(3, 2.0 * 4200000), (4, 5.0)] and np.array_equal(F.velocity, [0.5, -0.5])
    d, v = SolverBase.collect_dirichlet(bcs, mesh)
    assert np.allclose(v, 300 + mesh.coordinates()[d, 0])                       # string Expression interpolated at the vertices
    # initial values and translate_value
    assert np.all(solver.w_current.array() == 300)
    assert solver.translate_value((1, 2)).tolist() == [1.0, 2.0] and solver.translate_value(Constant(3)) == 3.0
    assert solver.get_variable_name() == 'temperature'
    with pytest.raises(TypeError):
        solver.translate_value(None)


def test_elasticity_form_and_per_component_dirichlet():
    mesh = BoxMesh(Point(0, 0, 0), Point(10, 1, 1), 6, 2, 2)

    class Left(SubDomain):
        def inside(self, x, on_boundary):
            return near(x[0], 0.0)

    class Right(SubDomain):
        def inside(self, x, on_boundary):
            return near(x[0], 10.0)
    import copy
    s = copy.deepcopy(SolverBase.default_case_settings)
    s['material'] = {'elastic_modulus': 2e11, 'poisson_ratio': 0.27, 'density': 7800}
    s['function_space'] = VectorFunctionSpace(mesh, "Lagrange", 1)
    s['boundary_conditions'] = {"fixed": {'boundary': Left(), 'boundary_id': 1, 'type': 'Dirichlet', 'value': (Constant(0), None, None)},
                                "displ": {'boundary': Right(), 'boundary_id': 2, 'type': 'Dirichlet', 'value': Constant((0, 0, 1e-3))}}
    s['body_source'] = (0.0, 0.0, -7800 * 9.8)
    solver = LinearElasticitySolver.LinearElasticitySolver(s)
    assert solver.settings['vector_name'] == 'displacement' and solver.function_space.ncomp == 3
    mu, lam = solver.lame_parameters()
    assert np.allclose((mu, lam), fo.lame(2e11, 0.27))
    solver.init_solver()
    solver.current_step = 0
    F, bcs = solver.generate_form(0, None, None, solver.w_current, solver.w_prev)
    assert F.load_sign == -1.0 and len(bcs) == 2 and bcs[0].component == 0 and bcs[1].component is None
    d, v = SolverBase.collect_dirichlet(bcs, mesh)
    nside = 9
    assert d.size == nside + 3 * nside
    c = mesh.coordinates()
    right = d[c[d // 3, 0] == 10]
    assert np.allclose(v[c[d // 3, 0] == 10][right % 3 == 2], 1e-3)
    assert np.all(d[c[d // 3, 0] == 0] % 3 == 0)


def test_expression_evaluation():
    c = np.array([[0.0, 0.5], [1.0, 2.0]])
    assert np.allclose(Expression("sin(x[0]) + pow(x[1], 2)", degree=1)(c), np.sin(c[:, 0]) + c[:, 1] ** 2)
    assert Expression(("10*rho", "0", "0.0"), rho=7800, degree=2)(np.zeros((3, 3))).shape == (3, 3)
    with pytest.raises(SolverBase.SolverError):
        Expression("__import__('os')")(c)
    # C++ conditionals and logic, as dolfin's JIT-compiled Expression strings allow
    p = np.array([[0.1, 0.2], [0.6, 0.9], [0.5, 0.5]])
    assert np.array_equal(Expression("x[0] < 0.5 ? 1.0 : (x[1] > 0.8 ? 3.0 : 2.0)", degree=0)(p), [1.0, 3.0, 2.0])
    assert np.array_equal(Expression("x[0] < 0.5 ? 1.0 : x[1] > 0.8 ? 3.0 : 2.0", degree=0)(p), [1.0, 3.0, 2.0])
    assert np.array_equal(Expression("(x[0] > 0.3 && x[1] < 0.95) ? T1 : T0", T0=300, T1=360, degree=1)(p), [300.0, 360.0, 360.0])
    assert np.array_equal(Expression("x[0] > 0.55 || !(x[1] > 0.3) ? 1 : 0", degree=1)(p), [1.0, 1.0, 0.0])
    assert np.allclose(Expression("pow(x[0] < 0.5 ? x[0] : 0.5, 2) + fmax(x[1], 0.5)", degree=1)(p), [0.01 + 0.5, 0.25 + 0.9, 0.25 + 0.5])


@pytest.mark.parametrize("mesh", [UnitSquareMesh(3, 2), UnitCubeMesh(2, 2, 3)], ids=["2d", "3d"])
def test_degree2_space_numbering_matches_oracle(mesh):
    """Host integer work of the P2 space (edge numbering, cell node lists, facet node lists, Dirichlet dofs)
    against the oracle's independent construction."""
    from oracle import fem_oracle_p2 as p2
    V = FunctionSpace(mesh, "Lagrange", 2)
    cn, xc, edges = p2.p2_dofmap(mesh.coordinates(), mesh.cells())
    assert V.num_nodes() == xc.shape[0] and np.array_equal(V.edges(), edges)
    assert np.array_equal(V.cell_nodes(), cn) and np.array_equal(V.node_coordinates(), xc)
    fv, _, _ = fo.exterior_facets(mesh.cells())
    assert np.array_equal(V.facet_nodes(fv), p2.facet_nodes(fv, edges, mesh.num_vertices()))
    W = VectorFunctionSpace(mesh, "CG", 2)
    assert W.dim() == xc.shape[0] * mesh.geometry().dim()
    # Dirichlet on x = 0 takes the edge midpoints of the marked facets as well as their vertices
    markers = FacetMarkers(mesh)
    markers.set_all(0)
    AutoSubDomain(lambda x: near(x[0], 0.0)).mark(markers, 1)
    from fenicssolver_b200.dolfin_compat import DirichletBC
    d, v = SolverBase.collect_dirichlet([DirichletBC(V, Constant(2.0), markers, 1)], V)
    assert np.array_equal(np.sort(d), np.flatnonzero(xc[:, 0] == 0.0)) and np.all(v == 2.0)
    f = Expression("x[0]*x[0] + x[1]", degree=2)
    d, v = SolverBase.collect_dirichlet([DirichletBC(V, f, markers, 1)], V)
    assert np.allclose(v, xc[d, 0] ** 2 + xc[d, 1])


def test_slab_partition_and_index_maps():
    parts = backend.slab_partition(257, 8)
    assert parts[0][0] == 0 and parts[-1][1] == 257 and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
    assert max(p1 - p0 for p0, p1 in parts) - min(p1 - p0 for p0, p1 in parts) <= 1
    assert backend.slab_partition(5, 2) == [(0, 3), (3, 5)]


GLOO_SCRIPT = r"""
import os, sys, numpy as np
sys.path.insert(0, %(root)r)
import torch.distributed as dist
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
from fenicssolver_b200 import backend
from fenicssolver_b200.dolfin_compat import UnitCubeMesh
comm = backend.Comm.from_torch()
assert comm.nranks == 2 and comm.rank == dist.get_rank()
uid = comm.bootstrap(b"x" * 128 if comm.rank == 0 else None)      # the NCCL-id broadcast path
assert uid == b"x" * 128
N = 6
mesh = UnitCubeMesh(N, N, N)
plane = (N + 1) ** 2
zp0, zp1 = backend.slab_partition(N + 1, comm.nranks)[comm.rank]
layer0, layer1 = max(zp0 - 1, 0), min(zp1, N)
v_off = layer0 * plane
nv_local = (layer1 - layer0 + 1) * plane
# every rank's local cells, renumbered, are exactly the global cells of its layers
per_layer = mesh.num_cells() // N
local_cells = mesh.cells()[per_layer * layer0:per_layer * layer1] - v_off
assert local_cells.min() >= 0 and local_cells.max() < nv_local
# owned rows of all ranks tile the global vertex range exactly once
owned = np.arange(zp0 * plane, zp1 * plane)
parts = [None, None]
dist.all_gather_object(parts, owned)
allv = np.concatenate(parts)
assert np.array_equal(allv, np.arange(mesh.num_vertices()))
# every neighbour column of an owned row is inside the local planes (one ghost plane suffices)
cells = mesh.cells()
touch = np.isin(cells, owned).any(axis=1)
assert cells[touch].min() >= v_off and cells[touch].max() < v_off + nv_local
# ---- slab-local boundary list (the rule Mesh.exterior_facets applies to the device search of a rank's slab): the exterior facets of
# the slab's own cells, minus those lying in a cut plane, shifted to global ids = the global boundary facets whose vertices are all
# local (what DeviceSpace.local_facets keeps of the global list)
from oracle import fem_oracle as fo0
from fenicssolver_b200.dolfin_compat import box_exterior_facets
fl, ol, _ = fo0.exterior_facets(np.ascontiguousarray(local_cells, dtype=np.int32))
zl = fl // plane
nplanes = layer1 - layer0 + 1
cut = np.zeros(fl.shape[0], dtype=bool)
if layer0 > 0:
    cut |= np.all(zl == 0, axis=1)
if layer1 < N:
    cut |= np.all(zl == nplanes - 1, axis=1)
mine = np.hstack([fl[~cut] + v_off, (ol[~cut] + v_off)[:, None]])
fg, og = box_exterior_facets((N, N, N))
keep = np.all((fg >= v_off) & (fg < v_off + nv_local), axis=1) & (og >= v_off) & (og < v_off + nv_local)
want = np.hstack([fg[keep], og[keep][:, None]])
assert np.array_equal(mine, want), (mine.shape, want.shape)
# a solver built with solver_settings['distributed'] on this 2-rank group marks its mesh as slab-distributed and, without a GPU,
# falls back to the direct enumeration of the box surface (host-only inspection): same markers on both ranks
from fenicssolver_b200 import ScalarTransportSolver
import bench
sv = ScalarTransportSolver.ScalarTransportSolver(bench.case_settings(N, distributed=True))
assert sv.mesh.distributed == (comm.rank, 2) and sv.mesh.slab_partition is True and sv.parallel
assert (sv.boundary_facets.values == 1).sum() == 2 * N * N and (sv.boundary_facets.values == 2).sum() == 2 * N * N
# ---- general node partition (unstructured meshes / degree 2): the halo lists drive a real exchange over gloo and the
# owner-computes SpMV of the locally assembled matrices reproduces the global product on the owned rows
import torch
from fenicssolver_b200.partition import NodePartition, rcb_partition
from oracle import fem_oracle as fo
c, t = fo.unit_cube_mesh(5, 4, 3)
rng = np.random.default_rng(0)
c = c + 0.02 * rng.standard_normal(c.shape)
part = rcb_partition(c, comm.nranks)
P = NodePartition(t, part, comm.rank, comm.nranks)
nv = c.shape[0]
A = fo.assemble_matrix(t, fo.local_laplace(c, t, 2.0) + fo.local_mass(c, t, 1.0), nv).tocsr()
xg = rng.standard_normal(nv)
Aloc = fo.assemble_matrix(np.sort(P.cell_nodes_local, axis=1), fo.local_laplace(c[P.l2g], np.sort(P.cell_nodes_local, axis=1), 2.0)
                          + fo.local_mass(c[P.l2g], np.sort(P.cell_nodes_local, axis=1), 1.0), P.n_local).tocsr()
xl = np.zeros(P.n_local)
xl[:P.n_owned] = xg[P.owned]                       # ghosts unknown until the halo exchange
reqs, bufs = [], []
for i, r in enumerate(P.neighbours):
    sb = torch.from_numpy(xl[P.send_idx[P.send_ptr[i]:P.send_ptr[i + 1]]].copy())
    rb = torch.empty(int(P.recv_cnt[i]), dtype=torch.float64)
    reqs += [dist.isend(sb, int(r)), dist.irecv(rb, int(r))]
    bufs.append((i, rb))
for q in reqs:
    q.wait()
for i, rb in bufs:
    xl[P.recv_off[i]:P.recv_off[i] + P.recv_cnt[i]] = rb.numpy()
assert np.array_equal(xl, xg[P.l2g])
yl = (Aloc @ xl)[:P.n_owned]
assert np.abs(yl - (A @ xg)[P.owned]).max() < 1e-12
dist.barrier()
dist.destroy_process_group()
print("rank", comm.rank, "ok")
"""


def test_world_size_two_gloo_partition(tmp_path):
    script = os.path.join(str(tmp_path), "gloo_part.py")
    open(script, "w").write(GLOO_SCRIPT % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", script], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


@pytest.mark.parametrize("nranks", [2, 3, 5])
@pytest.mark.parametrize("degree", [1, 2])
def test_node_partition_lists_are_consistent(nranks, degree):
    """partition.NodePartition for every rank of a jittered mesh: the owned sets tile the nodes, each neighbour pair agrees
    on the halo lists (what a sends b is b's ghost range from a, same order), local cells contain every cell of every
    owned node, and RCB is balanced and deterministic."""
    from fenicssolver_b200.dolfin_compat import FunctionSpace, Mesh
    from fenicssolver_b200.partition import NodePartition, rcb_partition
    from oracle import fem_oracle as fo
    c, t = fo.unit_cube_mesh(4, 3, 5)
    c = c + 0.03 * np.random.default_rng(1).standard_normal(c.shape)
    V = FunctionSpace(Mesh(c, t), "CG", degree)
    cn, xn = V.cell_nodes().astype(np.int64), V.node_coordinates()
    part = rcb_partition(xn, nranks)
    assert np.array_equal(part, rcb_partition(xn.copy(), nranks))
    counts = np.bincount(part, minlength=nranks)
    assert counts.max() - counts.min() <= 2
    P = [NodePartition(cn, part, r, nranks) for r in range(nranks)]
    assert np.array_equal(np.sort(np.concatenate([p.owned for p in P])), np.arange(xn.shape[0]))
    for a in range(nranks):
        pa = P[a]
        assert np.array_equal(pa.l2g[pa.g2l[pa.l2g]], pa.l2g) and pa.n_local == np.unique(pa.l2g).size
        # every cell touching an owned node is local, with all its nodes
        touch = np.isin(cn, pa.owned).any(axis=1)
        assert np.array_equal(np.nonzero(touch)[0], pa.cells_global)
        assert np.array_equal(pa.l2g[pa.cell_nodes_local], cn[touch])
        # ghosts are exactly the non-owned nodes of those cells
        assert np.array_equal(np.sort(pa.ghosts), np.setdiff1d(np.unique(cn[touch]), pa.owned))
        for i, b in enumerate(pa.neighbours):
            pb = P[b]
            j = int(np.nonzero(pb.neighbours == a)[0][0])
            sent = pa.l2g[pa.send_idx[pa.send_ptr[i]:pa.send_ptr[i + 1]]]
            got = pb.l2g[pb.recv_off[j]:pb.recv_off[j] + pb.recv_cnt[j]]
            assert np.array_equal(sent, got)
            assert np.all(part[sent] == a)
        assert pa.recv_cnt.sum() == pa.ghosts.size


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver times next to the GPU arm) on a small cube: exactly one JSON line on
    stdout with the contract's keys, timed on every core this process may use even when the launcher exported
    OMP_NUM_THREADS=1 (torchrun does, for every rank)."""
    import json
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--size", "16", "--steps", "4", "--warmup", "1"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "Mdof/s" and d["value"] > 0 and d["vs_baseline"] is None
    import bench
    assert d["config"] == {"workload": bench.workload(16)}                  # the very string the GPU arm prints
    assert d["converged"] == 1 and d["rel_l2_vs_exact"] < 1e-10 and d["steps"] == 4
    # the K steps are K segments of one complete step: value * (ms_per_step * steps) = DoF
    assert abs(d["value"] * 1e6 * d["ms_per_step"] * 1e-3 * d["steps"] - 17 ** 3) < 1e-6 * 17 ** 3
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mdof/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    assert d["cpu_baseline"]["cores"] == ncores


def test_multigrid_level_count_rule():
    """solve_amg picks the multigrid preconditioner only when the box can be coarsened (every cell count even and >= 4) on one GPU
    with a degree-1 space; otherwise Jacobi (SolverBase._multigrid_levels, no device call)."""
    import copy
    from fenicssolver_b200 import LinearElasticitySolver
    from fenicssolver_b200.dolfin_compat import BoxMesh, Constant, Mesh, Point, VectorFunctionSpace

    def solver(mesh, degree=1):
        s = copy.deepcopy(SolverBase.default_case_settings)
        s.update({'function_space': VectorFunctionSpace(mesh, "CG", degree),
                  'material': {'elastic_modulus': 2e11, 'poisson_ratio': 0.27, 'density': 7800},
                  'boundary_conditions': {'clamp': {'boundary': AutoSubDomain(lambda x: near(x[0], 0.0)), 'boundary_id': 1, 'type': 'Dirichlet',
                                                    'value': Constant((0, 0, 0))}},
                  'report_settings': {'logging_level': 40, 'logging_file': None, 'plotting_freq': 0, 'saving_freq': 0}})
        return LinearElasticitySolver.LinearElasticitySolver(s)
    box = lambda *n: BoxMesh(Point(0, 0, 0), Point(4, 1, 1), *n)       # noqa: E731
    assert solver(box(128, 128, 128))._multigrid_levels() == 7        # 128 -> 64 -> ... -> 2
    assert solver(box(32, 8, 8))._multigrid_levels() == 3             # 8 -> 4 -> 2 limits it
    assert solver(box(16, 4, 4))._multigrid_levels() == 2
    assert solver(box(12, 3, 3))._multigrid_levels() == 1             # odd count: no coarser level -> Jacobi
    assert solver(box(16, 4, 4), degree=2)._multigrid_levels() == 0   # degree 2: not covered
    m = box(4, 4, 4)
    sv = solver(m)
    sv.mesh = Mesh(m.coordinates().copy(), m.cells().copy())          # not a generated box (constructing a solver on it needs the device)
    assert sv._multigrid_levels() == 0


def test_xdmf_ascii_mesh_reader(tmp_path):
    """read_mesh('*.xdmf') (SolverBase.py:246-252): inline-XML XDMF as dolfin's ASCII encoding writes it; HDF5-backed files raise."""
    from fenicssolver_b200.dolfin_compat import read_xdmf_mesh
    c, t = fo.unit_cube_mesh(2, 2, 1)
    p = os.path.join(str(tmp_path), "mesh.xdmf")
    with open(p, "w") as f:
        f.write('<?xml version="1.0"?>\n<!DOCTYPE Xdmf SYSTEM "Xdmf.dtd" []>\n<Xdmf Version="3.0"><Domain><Grid Name="mesh" GridType="Uniform">\n')
        f.write('<Topology NumberOfElements="%d" TopologyType="Tetrahedron" NodesPerElement="4"><DataItem Dimensions="%d 4" NumberType="UInt" Format="XML">\n' % (t.shape[0], t.shape[0]))
        f.write("\n".join(" ".join(str(v) for v in row[::-1]) for row in t))          # unsorted on purpose
        f.write('\n</DataItem></Topology>\n<Geometry GeometryType="XYZ"><DataItem Dimensions="%d 3" Format="XML">\n' % c.shape[0])
        f.write("\n".join(" ".join(repr(float(x)) for x in row) for row in c))
        f.write('\n</DataItem></Geometry></Grid></Domain></Xdmf>\n')
    c2, t2 = read_xdmf_mesh(p)
    assert np.array_equal(c2, c) and np.array_equal(t2, t)
    from fenicssolver_b200 import ScalarTransportSolver
    s = {'solver_name': 'ScalarTransportSolver', 'scalar_name': 'temperature', 'mesh': p,
         'material': {'density': 1000, 'specific_heat_capacity': 500, 'thermal_conductivity': 20},
         'boundary_conditions': {'inlet': {'boundary': lambda x: near(x[2], 0.0), 'boundary_id': 1, 'type': 'Dirichlet', 'value': 350}},
         'body_source': None, 'initial_values': {'temperature': 293},
         'solver_settings': {'transient_settings': {'transient': False, 'starting_time': 0, 'time_step': 0.01, 'ending_time': 0.03},
                             'reference_values': {'temperature': 293}, 'solver_parameters': {}},
         'report_settings': {'logging_level': 40, 'logging_file': None, 'plotting_freq': 0, 'saving_freq': 0}}
    from fenicssolver_b200.dolfin_compat import _device_present
    if _device_present():              # the boundary search of a file mesh is device work (K1)
        solver = ScalarTransportSolver.ScalarTransportSolver(s)
        assert solver.mesh.num_vertices() == c.shape[0] and (solver.boundary_facets.values == 1).sum() == 8
    else:
        with pytest.raises(SolverBase.SolverError):
            ScalarTransportSolver.ScalarTransportSolver(s)
    h5 = os.path.join(str(tmp_path), "mesh_h5.xdmf")
    open(h5, "w").write(open(p).read().replace('Format="XML"', 'Format="HDF"'))
    with pytest.raises(SolverBase.SolverError):
        read_xdmf_mesh(h5)
    with pytest.raises(SolverBase.SolverError):
        s2 = dict(s, mesh=os.path.join(str(tmp_path), "mesh.h5"))
        open(s2['mesh'], "w").write("x")
        ScalarTransportSolver.ScalarTransportSolver(s2)


def test_function_point_evaluation():
    """u(x, y) / u(Point(...)): P1 reproduces linear fields, P2 quadratic ones, vector spaces return one value per component."""
    from fenicssolver_b200.dolfin_compat import Function, Point
    mesh = UnitSquareMesh(5, 4)
    for degree in (1, 2):
        V = FunctionSpace(mesh, "CG", degree)
        xn = V.node_coordinates()
        f = (lambda x: 1 + 2 * x[:, 0] - 3 * x[:, 1]) if degree == 1 else (lambda x: 1 + x[:, 0] ** 2 - 2 * x[:, 0] * x[:, 1])
        u = Function(V, f(xn))
        for p in ((0.33, 0.71), (0.0, 1.0), (0.6, 0.25)):
            assert abs(u(*p) - f(np.array([p]))[0]) < 1e-13
            assert abs(u(Point(*p)) - u(p)) == 0.0
    W = VectorFunctionSpace(UnitCubeMesh(2, 2, 2), "CG", 1)
    xn = W.node_coordinates()
    u = Function(W, np.stack([xn[:, 0], 2 * xn[:, 1], xn[:, 0] + xn[:, 2]], axis=1).ravel())
    assert np.allclose(u(0.3, 0.4, 0.9), [0.3, 0.8, 1.2])
    with pytest.raises(SolverBase.SolverError):
        u(1.5, 0.2, 0.2)


def test_dolfin_script_stand_ins(tmp_path):
    """The calls the reference's example scripts make around the solver: set_log_level(ERROR), plot(...), File(...) << u, and the flux
    integral their post_process() prints (examples/test_heat_transfer.py:181-190)."""
    from fenicssolver_b200 import dolfin_compat as dc
    from fenicssolver_b200.dolfin_compat import ERROR, File, Function, boundary_flux, interactive, plot, set_log_level
    set_log_level(ERROR)
    assert plot(None, title="x") is None and interactive() is None and dc.parameters["form_compiler"]["optimize"]
    mesh = UnitSquareMesh(6, 5)
    V = FunctionSpace(mesh, "CG", 1)
    c = mesh.coordinates()
    T = Function(V, 360 + 60 * (1 - c[:, 1]) + 5 * c[:, 0])          # linear: gradient (5, -60)
    markers = FacetMarkers(mesh)
    markers.set_all(0)
    AutoSubDomain(lambda x: near(x[1], 0.0)).mark(markers, 1)
    AutoSubDomain(lambda x: near(x[0], 1.0)).mark(markers, 2)
    assert abs(boundary_flux(T, markers, 1, 0.6) - 0.6 * 60.0) < 1e-12       # n = (0, -1) on y = 0, length 1
    assert abs(boundary_flux(T, markers, 2) - 5.0) < 1e-12                   # n = (1, 0) on x = 1
    assert boundary_flux(T, markers, 7) == 0.0
    f = File(os.path.join(str(tmp_path), "T.pvd"))
    f << (T, 0.0)
    f << (T, 0.5)
    import xml.etree.ElementTree as ET
    ds = ET.parse(os.path.join(str(tmp_path), "T.pvd")).getroot().findall("./Collection/DataSet")
    assert [d.attrib["file"] for d in ds] == ["T000000.vtu", "T000001.vtu"]
    assert os.path.exists(os.path.join(str(tmp_path), "T000001.vtu"))
    W = VectorFunctionSpace(UnitCubeMesh(2, 2, 2), "CG", 1)
    with pytest.raises(SolverBase.SolverError):
        boundary_flux(Function(W), markers, 1)


def test_cell_function_marking_for_subdomain_sources():
    """SubDomain.mark on a cell MeshFunction (what a script does to give `body_source` per subdomain, ScalarTransportSolver.py:213-221,
    without a _physical_region.xml): all vertices and the midpoint inside."""
    from fenicssolver_b200.dolfin_compat import MeshFunction
    mesh = UnitSquareMesh(4, 4)
    sub = MeshFunction("size_t", mesh, mesh.topology().dim())
    sub.set_all(0)
    AutoSubDomain(lambda x: x[0] <= 0.5 + 1e-12).mark(sub, 3)
    c, t = mesh.coordinates(), mesh.cells()
    expect = (c[t][:, :, 0] <= 0.5 + 1e-12).all(axis=1)
    assert np.array_equal(sub.array() == 3, expect) and expect.sum() == 16
    class Right(SubDomain):
        def inside(self, x, on_boundary):
            return x[0] >= 0.75 - 1e-12 and not on_boundary        # cells are never "on_boundary"
    Right().mark(sub, 5)
    assert (sub.array() == 5).sum() == 8
    vf = MeshFunction("size_t", mesh, 0)
    AutoSubDomain(lambda x: near(x[1], 1.0)).mark(vf, 1)
    assert (vf.array() == 1).sum() == 5


def test_generic_vector_surface():
    """u.vector(): the GenericVector calls FEniCS scripts use on a result (get_local/set_local/apply, norms, min/max/sum, indexing)."""
    from fenicssolver_b200.dolfin_compat import Function
    V = FunctionSpace(UnitSquareMesh(2, 2), "CG", 1)
    u = Function(V, np.arange(9.0) - 4.0)
    v = u.vector()
    assert len(v) == v.size() == 9 and v.max() == 4.0 and v.min() == -4.0 and v.sum() == 0.0
    assert abs(v.norm("l2") - np.sqrt(60.0)) < 1e-14 and v.norm("l1") == 20.0 and v.norm("linf") == 4.0
    v.set_local(np.ones(9))
    v.apply("insert")
    v[0] = 5.0
    assert u.values[0] == 5.0 and np.array_equal(v.get_local()[1:], np.ones(8))
    g = v.get_local()
    g[:] = 0.0
    assert u.values[1] == 1.0                     # get_local returns a copy
    with pytest.raises(SolverBase.SolverError):
        v.norm("frobenius")


def test_expression_is_validated_and_follows_cpp_integer_division():
    """Expression strings (they also arrive in JSON case files) are parsed and restricted to arithmetic on whitelisted names: no
    attribute access, no dunder names, no strings or lambdas; `!x` is logical not; a quotient of two integer literals truncates as
    in the reference's compiled C++ Expression."""
    c = np.array([[0.2, 0.7], [0.8, 0.1]])
    assert np.allclose(Expression("x[0] < 0.5 && !(x[1] < 0.5) ? 1/2 + 1.0/2 : pow(x[0], 2)")(c), [0.5, 0.64])
    assert np.allclose(Expression("a*sin(pi*x[0]) + 7/2", a=2.0)(c), 2.0 * np.sin(np.pi * c[:, 0]) + 3)
    assert np.allclose(Expression("-7/2 + x[1]")(c), -3 + c[:, 1])
    for bad in ["x.__class__", "__import__('os')", "x[0].real", "(lambda: 1)()", "'a'*3", "x[0] if 1 else 2", "[v for v in x]"]:
        with pytest.raises(SolverBase.SolverError):
            Expression(bad)(c)


def test_box_surface_enumeration_and_lazy_coordinates():
    """box_exterior_facets enumerates the surface of the dolfin box layout directly (no search): the same facets, opposite vertices AND
    order as the oracle's facet table; Mesh.vertex_coordinates computes selected vertices of a generated box bit-identically to
    coordinates() without materialising the mesh; cells_sorted adopts caller arrays as they are."""
    from fenicssolver_b200.dolfin_compat import box_exterior_facets
    for n in [(5, 3), (1, 1), (4, 3, 2), (1, 1, 1), (2, 1, 3), (6, 5, 7)]:
        c, t = (fo.unit_square_mesh(*n) if len(n) == 2 else fo.unit_cube_mesh(*n))
        fv, opp = box_exterior_facets(n)
        f0, o0, _ = fo.exterior_facets(t)
        assert np.array_equal(fv, f0) and np.array_equal(opp, o0), n
    m = BoxMesh(Point(0, -1, 0.5), Point(10, 1, 1.75), 5, 4, 3)
    ids = np.array([0, 7, 119, 3, 64])
    got = m.vertex_coordinates(ids)
    assert m._coords is None                                         # nothing was materialised
    assert np.array_equal(got, BoxMesh(Point(0, -1, 0.5), Point(10, 1, 1.75), 5, 4, 3).coordinates()[ids])
    cells = np.ascontiguousarray(m.cells(), dtype=np.int32)
    m2 = Mesh(m.coordinates(), cells, cells_sorted=True)
    assert m2.cells() is cells                                       # adopted, not copied
    assert Mesh(m.coordinates(), cells[:, ::-1]).cells() is not cells and np.array_equal(Mesh(m.coordinates(), cells[:, ::-1]).cells(), cells)
