"""Run under torchrun on >= 2 GPUs: distributed solves on a GENERAL node partition (recursive coordinate bisection,
owned-first local numbering, send-list halo) compared on rank 0 with single-GPU solves of the same problems:
 A  heat with flux + HTC + source on an unstructured (jittered, user-array) tetrahedral mesh, Jacobi-CG
 B  the reference's elasticity example on its own degree-2 space (P2 nodes partitioned), Jacobi-CG on 3x3 blocks
 C  transient advection-diffusion on the unstructured mesh (BiCGStab, T_prev halo every step)
 D  radiation boundary on the unstructured mesh: Newton iterations with distributed residuals and updates
 E  transient heat with a nodal velocity FIELD and a point source, on the unstructured mesh (RCB) and on a generated box (z-slabs)
 F  degree-1 elasticity with a nodal temperature distribution, then the von Mises projection of the result, RCB and z-slabs."""
import copy
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

QUIET = {'logging_level': 40, 'logging_file': None, 'plotting_freq': 0, 'saving_freq': 0, 'plotting_interactive': False}


def heat_case(mesh, distributed, transient=False, radiation=False):
    from fenicssolver_b200.dolfin_compat import near
    k, rho, cp = 0.6, 1000.0, 4200.0
    n = 10
    dt = rho * cp / (n * n) / k
    s = {'solver_name': 'ScalarTransportSolver', 'scalar_name': 'temperature', 'mesh': mesh, 'fe_degree': 1, 'fe_family': 'CG',
         'material': {'density': rho, 'specific_heat_capacity': cp, 'thermal_conductivity': k},
         'boundary_conditions': {
             'hot': {'boundary': lambda x: near(x[2], 0.0), 'boundary_id': 1, 'type': 'Dirichlet', 'value': 360},
             'htc': {'boundary': lambda x: near(x[2], 1.0), 'boundary_id': 2, 'type': 'HTC', 'value': 25.0, 'ambient': 300.0},
             'flux': {'boundary': lambda x: near(x[0], 0.0), 'boundary_id': 3, 'type': 'heatFlux', 'value': 40.0}},
         'body_source': 500.0, 'initial_values': {'temperature': 300},
         'solver_settings': {'transient_settings': {'transient': transient, 'starting_time': 0.0, 'time_step': dt, 'ending_time': dt * 3.5},
                             'reference_values': {'temperature': 300}, 'solver_parameters': {},
                             'distributed': distributed, 'gather_result': True},
         'report_settings': QUIET}
    if transient:
        s['convective_velocity'] = (0.0, 1e-6, 2e-6)
    if radiation:
        s['radiation_settings'] = {'ambient_temperature': 280.0, 'emissivity': 0.9}      # Newton, all exterior facets
    return s


def field_case(mesh, coords, distributed):
    """E: velocity field u(x) = (0, 1e-6 (1 + x), 2e-6 (1 + y)) given per vertex, and a point source inside the domain."""
    from fenicssolver_b200.dolfin_compat import Point
    s = heat_case(mesh, distributed, transient=True)
    vel = np.zeros((coords.shape[0], 3))
    vel[:, 1] = 1e-6 * (1.0 + coords[:, 0])
    vel[:, 2] = 2e-6 * (1.0 + coords[:, 1])
    s['convective_velocity'] = vel
    s['point_source'] = [(Point(0.52, 0.47, 0.55), 2.0e4)]
    return s


def thermoelastic_case(mesh, coords, distributed):
    """F: clamp on x = 0, gravity, and a nodal temperature field T(x) = 293 + 60 x z."""
    from fenicssolver_b200.dolfin_compat import Expression, near
    return {'solver_name': 'LinearElasticitySolver', 'mesh': mesh, 'fe_degree': 1, 'fe_family': 'CG', 'vector_name': 'displacement',
            'material': {'elastic_modulus': 2e11, 'poisson_ratio': 0.27, 'density': 7800, 'thermal_expansion_coefficient': 2e-6},
            'boundary_conditions': {'clamp': {'boundary': lambda x: near(x[0], 0.0), 'boundary_id': 1, 'type': 'Dirichlet', 'value': (0, 0, 0)}},
            'body_source': (0.0, 0.0, -7800 * 9.81), 'initial_values': {},
            'temperature_distribution': Expression("293 + 60*x[0]*x[2]", degree=1),       # interpolated: one value per node
            'solver_settings': {'transient_settings': {'transient': False, 'starting_time': 0, 'time_step': 0.01, 'ending_time': 0.03},
                                'reference_values': {'temperature': 293}, 'solver_parameters': {'preconditioner': 'jacobi'},
                                'distributed': distributed, 'gather_result': True},
            'report_settings': QUIET}


def elasticity_case(distributed):
    from fenicssolver_b200 import SolverBase
    from fenicssolver_b200.dolfin_compat import AutoSubDomain, BoxMesh, Constant, Point, VectorFunctionSpace, near
    mesh = BoxMesh(Point(0, 0, 0), Point(10, 1, 1), 8, 2, 2)
    s = copy.deepcopy(SolverBase.default_case_settings)
    s.update({'material': {'elastic_modulus': 2e11, 'poisson_ratio': 0.27, 'density': 7800, 'thermal_expansion_coefficient': 2e-6},
              'function_space': VectorFunctionSpace(mesh, "Lagrange", 2), 'report_settings': QUIET,
              'boundary_conditions': {'fixed': {'boundary': AutoSubDomain(lambda x: near(x[0], 0.0)), 'boundary_id': 1, 'type': 'Dirichlet',
                                                'value': Constant((0, 0, 0))},
                                      'pull': {'boundary': AutoSubDomain(lambda x: near(x[0], 10.0)), 'boundary_id': 2, 'type': 'pressure', 'value': 1e6}},
              'body_source': (0.0, 0.0, -7800 * 9.81), 'temperature_distribution': 343.0})
    s['solver_settings'] = dict(s['solver_settings'], reference_values={'temperature': 293}, distributed=distributed, gather_result=True)
    return s


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from fenicssolver_b200 import LinearElasticitySolver, ScalarTransportSolver, _lib, backend
    from fenicssolver_b200.dolfin_compat import Mesh, UnitCubeMesh
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    base = UnitCubeMesh(n, n, n)
    c = base.coordinates().copy()
    interior = np.all((c > 0) & (c < 1), axis=1)
    c[interior] += 0.2 / n * (np.random.default_rng(0).random((int(interior.sum()), 3)) * 2 - 1)
    mesh = Mesh(c, base.cells())                          # user arrays: no box description -> general partition
    ok = True
    ctx1 = _lib.Context(local) if rank == 0 else None

    def single(solver_cls, settings, ncomp):
        sv = solver_cls(settings)
        sv._space = backend.DeviceSpace(sv.mesh, ncomp, ctx=ctx1, space=sv.function_space)
        return sv.solve().vector().get_local(), sv

    cases = [("A heat/unstructured", ScalarTransportSolver.ScalarTransportSolver, lambda d: heat_case(mesh, d), 1),
             ("B elasticity/P2", LinearElasticitySolver.LinearElasticitySolver, elasticity_case, 3),
             ("C transient/unstructured", ScalarTransportSolver.ScalarTransportSolver, lambda d: heat_case(mesh, d, True), 1),
             ("D radiation-Newton/unstructured", ScalarTransportSolver.ScalarTransportSolver, lambda d: heat_case(mesh, d, radiation=True), 1)]
    box = UnitCubeMesh(n, n, n)                           # generated box -> z-slabs
    bc_ = box.coordinates()
    cases += [("E velocity field + point source/unstructured", ScalarTransportSolver.ScalarTransportSolver, lambda d: field_case(mesh, c, d), 1),
              ("E velocity field + point source/slabs", ScalarTransportSolver.ScalarTransportSolver, lambda d: field_case(UnitCubeMesh(n, n, n), bc_, d), 1),
              ("F nodal temperature + von Mises/unstructured", LinearElasticitySolver.LinearElasticitySolver, lambda d: thermoelastic_case(mesh, c, d), 3),
              ("F nodal temperature + von Mises/slabs", LinearElasticitySolver.LinearElasticitySolver, lambda d: thermoelastic_case(UnitCubeMesh(n, n, n), bc_, d), 3)]
    for name, cls, make, ncomp in cases:
        sv = cls(make(True))
        xd = sv.solve().vector().get_local()
        sp_ = sv.device_space()
        slabs = name.endswith("/slabs")
        assert (sp_.part is None) == slabs, "unexpected partition type for %s" % name
        info = sv.solve_info
        stats = (sp_.part.n_owned, sp_.part.n_local - sp_.part.n_owned, len(sp_.part.neighbours)) if not slabs else (sp_.own_v1 - sp_.own_v0, sp_.nv_local - (sp_.own_v1 - sp_.own_v0), sp_.ghost_lo + sp_.ghost_hi)
        vm_d = sv.von_Mises(sv.result).vector().get_local() if name.startswith("F") else None
        allst = [None] * world
        dist.all_gather_object(allst, stats)
        if rank == 0:
            x1, sv1 = single(cls, make(False), ncomp)
            d = np.linalg.norm(xd - x1) / np.linalg.norm(x1)
            good = d < 1e-9 and info["converged"] == 1
            if vm_d is not None:
                vm1 = sv1.von_Mises(sv1.result).vector().get_local()
                dv = np.linalg.norm(vm_d - vm1) / np.linalg.norm(vm1)
                name = name + " (von Mises rel_l2 %.2e)" % dv
                good = good and dv < 1e-9
            print("%s: world=%d (owned, ghosts, neighbours) per rank=%s iters dist=%d single=%d rel_l2(dist vs single)=%.2e %s"
                  % (name, world, allst, info["iterations"], sv1.solve_info["iterations"], d, "ok" if good else "FAILED"), flush=True)
            ok = ok and good
        dist.barrier()
    if rank == 0:
        print("DIST_GENERAL_OK" if ok else "DIST_GENERAL_FAILED", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
