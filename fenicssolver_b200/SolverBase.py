"""SolverBase — settings parsing, mesh/marker loading, value translation, time loop, linear solve.

Same class, attribute, method and settings-key names as the reference
(/root/reference/FenicsSolver/SolverBase.py:64-88 defaults, :95-182 construction, :285-465 helpers,
:484-546 time loop, :592-672 solve wrappers), with the dolfin/PETSc calls replaced by the libfsb
device pipeline (backend.DeviceSpace).  Reference quirks that change results are reproduced or listed
as deviations in DESIGN.md.
"""
from __future__ import annotations

import copy
import logging
import numbers
import os.path
import time

import numpy as np

from ._lib import SolverError
from .backend import Comm, DeviceSpace
from .dolfin_compat import (Constant, DirichletBC, Expression, FacetMarkers, Function, FunctionSpace, Mesh, MeshFunction,
                            VectorFunctionSpace)

default_report_settings = {"logging_level": logging.DEBUG, "logging_file": None,
                           "plotting_freq": 10, "plotting_interactive": True, "plotting_file": None,
                           "saving_freq": 10, "result_filename": None}

# the reference forwards these to dolfin, where the keys do not exist, so the scalar path is solved by
# sparse LU whatever they say (SolverBase.py:638-641).  Here the solve is Krylov:
# 'parity_mode' (default True) ignores loose user tolerances and converges to 1e-12 so the answer
# matches the reference's direct solve to 1e-10; set it False to honour the values below.
default_solver_parameters = {"relative_tolerance": 1e-5,
                             "maximum_iterations": 500,
                             "monitor_convergence": True,
                             }
default_case_settings = {"solver_name": None,
                         "case_name": "test", "case_folder": "./", "case_file": None,
                         "mesh": None, "fe_degree": 1, "fe_family": "CG",
                         "function_space": None, "periodic_boundary": None,
                         "boundary_conditions": None,
                         "body_source": None,
                         "surface_source": None,
                         "initial_values": {},
                         "material": {},
                         "solver_settings": {
                             "transient_settings": {"transient": False, "starting_time": 0, "time_step": 0.01, "ending_time": 0.03},
                             "reference_values": {},
                             "solver_parameters": default_solver_parameters,
                         },
                         "report_settings": default_report_settings
                         }

PARITY_RTOL = 1e-12
PARITY_MAXIT = 200000


class SolverBase():
    """shared base class: solve(), plot(), save(); generate_form() and update_boundary_conditions()
    are implemented by the physics solvers."""

    def __init__(self, case_input):
        if isinstance(case_input, (dict)):
            self.settings = case_input
            self.load_settings(case_input)
        else:
            raise SolverError('case setup data must be a python dict')
        # the reference probes MPI here (SolverBase.py:102-118); the equivalent is an initialised
        # torch.distributed group, used only when solver_settings['distributed'] asks for it
        self.comm = Comm()
        if self.solver_settings.get('distributed'):
            self.comm = Comm.from_torch()
        self.parallel = self.comm.nranks > 1
        if self.parallel and self.comm.rank != 0:
            self.logger.disabled = True
        self.timings = {}
        self.solve_info = None
        self._space = None

    def print(self):
        import pprint
        pprint.PrettyPrinter(indent=4).pprint(self.settings)

    # ------------------------------------------------------------------ settings / mesh
    def load_settings(self, s):
        if 'periodic_boundary' not in s:
            s['periodic_boundary'] = None
        self.boundary_conditions = s['boundary_conditions']
        if ('mesh' in s) and s['mesh']:
            if isinstance(s['mesh'], str):
                self.read_mesh(s['mesh'])
            elif isinstance(s['mesh'], Mesh):
                self.mesh = s['mesh']
                self.generate_boundary_facets()
            elif isinstance(s['mesh'], dict):        # plain-data mesh: {'type': 'UnitCubeMesh', 'n': [..], ...}
                self.mesh = mesh_from_dict(s['mesh'])
                self.generate_boundary_facets()
            else:
                raise SolverError('Error: mesh must be file path or Mesh object: {}')
            if 'fe_family' not in s:
                s['fe_family'] = 'CG'
            if 'fe_degree' not in s:
                s['fe_degree'] = 1
            self.generate_function_space(s['periodic_boundary'])
        elif ('mesh' not in s or s['mesh'] is None) and ('function_space' in s and s['function_space']):
            self.function_space = s['function_space']
            s['fe_degree'] = self.function_space._ufl_element.degree()
            if 'fe_family' not in s:
                s['fe_family'] = 'CG'
            self.mesh = self.function_space.mesh()
            self.generate_boundary_facets()
            self.is_mixed_function_space = False
        else:
            raise SolverError('mesh or function space must specified to construct solver object')
        self.dimension = self.mesh.geometry().dim()
        self.topo_dimension = self.mesh.topology().dim()

        if not hasattr(self, 'subdomains'):
            self.subdomains = None          # empty cell MeshFunction, materialised on demand
        if 'body_source' in s and s['body_source']:
            self.body_source = s['body_source']
        else:
            self.body_source = None

        if 'initial_values' in s:
            self.initial_values = s['initial_values']
        else:
            self.initial_values = {}
        self.reference_values = s['solver_settings']['reference_values']
        self.material = s['material']
        self.solver_settings = s['solver_settings']
        self.transient_settings = s['solver_settings']['transient_settings']
        self.transient = self.transient_settings['transient']

        if 'report_settings' not in self.settings:
            self.settings['report_settings'] = default_report_settings
        self.report_settings = self.settings['report_settings']
        self.set_logger(self.settings['report_settings'])

    def set_logger(self, s):
        logger = logging.getLogger(self.__class__.__name__)
        if not logger.handlers:
            if ('logging_file' not in s) or (s['logging_file'] is None):
                fh = logging.StreamHandler()
            else:
                fh = logging.FileHandler(s['logging_file'])
            fh.setLevel(s.get('logging_level', logging.DEBUG))
            fh.setFormatter(logging.Formatter('%(asctime)s - %(name)s - %(levelname)s - %(message)s'))
            logger.addHandler(fh)
        logger.setLevel(s.get('logging_level', logging.DEBUG))
        self.logger = logger

    def _read_xml_mesh(self, filename):
        mesh = Mesh(filename)
        bmeshfile = filename[:-4] + "_facet_region.xml"
        self.mesh = mesh
        self._mark_distribution()
        if os.path.exists(bmeshfile):
            self.boundary_facets = MeshFunction("size_t", mesh, bmeshfile)
        else:
            self.generate_boundary_facets()
        subdomain_meshfile = filename[:-4] + "_physical_region.xml"
        if os.path.exists(subdomain_meshfile):
            self.subdomains = MeshFunction("size_t", mesh, subdomain_meshfile)
        else:
            self.subdomains = None

    def read_mesh(self, filename):
        if not os.path.exists(filename):
            raise SolverError('mesh file: {} , does not exist'. format(filename))
        if filename[-4:] == ".xml":
            self._read_xml_mesh(filename)
        elif filename[-5:] == ".xdmf":
            # mesh only, boundaries from the predicates (SolverBase.py:246-252); ASCII-encoded XDMF (no HDF5 library here)
            from .dolfin_compat import read_xdmf_mesh
            self.mesh = Mesh(*read_xdmf_mesh(filename))
            self.generate_boundary_facets()
            self.subdomains = None
        elif filename[-3:] == ".h5" or filename[-5:] == ".hdf5":
            raise SolverError('HDF5 meshes cannot be read here (no HDF5 library in this image); convert to dolfin-XML or ASCII XDMF')
        else:
            raise SolverError('mesh or function space must specified to construct solver object')

    def generate_function_space(self, periodic_boundary):
        self.is_mixed_function_space = False
        if "scalar_name" in self.settings:
            self.function_space = FunctionSpace(self.mesh, self.settings['fe_family'], self.settings['fe_degree'],
                                                constrained_domain=periodic_boundary)
        elif "vector_name" in self.settings:
            self.function_space = VectorFunctionSpace(self.mesh, self.settings['fe_family'], self.settings['fe_degree'],
                                                      constrained_domain=periodic_boundary)
        else:
            raise SolverError('only scalar or vector solver has a base method of generate_function_space()')

    def _mark_distribution(self):
        """Tell the mesh whether this solver runs slab/RCB-distributed (settings['solver_settings']['distributed'] with an
        initialised torch.distributed group of more than one rank): no rank then holds the whole mesh on its device."""
        on = None
        if (self.settings.get('solver_settings') or {}).get('distributed'):
            try:
                import torch.distributed as dist
                if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                    on = (dist.get_rank(), dist.get_world_size())
            except ImportError:
                on = None
        # the z-slab layout (and with it the slab-local boundary search) is the one of degree-1 spaces on generated boxes; degree 2
        # and array meshes are RCB-partitioned and keep the global facet list
        fs = getattr(self, 'function_space', None) if 'function_space' in self.settings and self.settings.get('function_space') else None
        degree = fs.degree if fs is not None else int(self.settings.get('fe_degree', 1) or 1)
        slab = degree == 1
        if getattr(self.mesh, 'distributed', None) != on or getattr(self.mesh, 'slab_partition', True) != slab:
            for k in ('_exterior', '_boundary_geometry'):            # lists made for another layout do not carry over
                if k in self.mesh.__dict__:
                    setattr(self.mesh, k, None)
        self.mesh.distributed = on
        self.mesh.slab_partition = slab

    def generate_boundary_facets(self):
        self._mark_distribution()
        boundary_facets = FacetMarkers(self.mesh)
        boundary_facets.set_all(0)
        for name, bc in self.boundary_conditions.items():
            if bc.get('boundary') is None:
                raise SolverError("boundary '{}' has no 'boundary' predicate and the mesh has no marker file".format(name))
            boundary_facets.mark_subdomain(bc['boundary'], bc['boundary_id'])      # dict order: later ids overwrite
        self.boundary_facets = boundary_facets

    # ------------------------------------------------------------------ values
    def get_initial_field(self):
        if not self.initial_values:
            if self.is_mixed_function_space:
                return Function(self.function_space)
            elif 'vector_name' in self.settings:
                v0 = (0, ) * self.dimension
            elif 'scalar_name' in self.settings:
                v0 = 0
            else:
                raise SolverError('only vector and scalar equation can run this method')
        else:
            if self.is_mixed_function_space:
                raise SolverError('only vector and scalar function can run this method')
            elif 'vector_name' in self.settings:
                v0 = self.initial_values[self.settings['vector_name']]
            elif 'scalar_name' in self.settings:
                v0 = self.initial_values[self.settings['scalar_name']]
            else:
                raise SolverError('only vector and scalar function can run this method')

        nv = self.function_space.num_nodes()
        if 'vector_name' in self.settings and isinstance(v0, (tuple, list)) and isinstance(v0[0], (str, numbers.Number)):
            if all(isinstance(v, numbers.Number) for v in v0) and len(set(float(v) for v in v0)) == 1:
                # the same constant in every component (the usual zero start): stays symbolic, filled on the device — no host
                # arrays, no H2D copy (51 MB each for w_current / w_prev / w_pp of a 128^3 elasticity run)
                return Function(self.function_space, fill=float(v0[0]))
            if all(isinstance(v, numbers.Number) for v in v0):
                vals = np.tile(np.asarray(v0, dtype=np.float64), nv)
            else:
                vals = Expression(tuple(str(v) for v in v0), degree=self.settings['fe_degree'])(self.function_space.node_coordinates()).reshape(-1)
            u0 = Function(self.function_space, vals)
        elif 'scalar_name' in self.settings and isinstance(v0, numbers.Number):
            u0 = Function(self.function_space, fill=float(v0))        # stays symbolic: filled on the device
        elif 'scalar_name' in self.settings and isinstance(v0, str) and not os.path.exists(v0):
            u0 = Function(self.function_space, Expression(v0, degree=self.settings['fe_degree'])(self.function_space.node_coordinates()))
        elif isinstance(v0, Function):
            u0 = v0.copy()
        elif isinstance(v0, np.ndarray) and v0.size == self.function_space.dim():
            u0 = Function(self.function_space, v0)
        elif isinstance(v0, str) and os.path.exists(v0):
            u0 = Function(self.function_space, np.load(v0))            # restart from a saved nodal vector
        else:
            raise SolverError('only number, file, another function, str expr are supported as initial values')
        return u0

    def get_material_value(self, value):
        if isinstance(value, (list, tuple, np.ndarray)) and len(value) == self.dimension:
            if len(value[0]) == self.dimension and isinstance(value[0][0], numbers.Number):
                return np.asarray(value, dtype=np.float64)            # anisotropic tensor
        elif isinstance(value, dict):
            raise SolverError('multi-region material dicts are not implemented')
        if isinstance(value, Constant):
            v = value.values()
            return float(v[0]) if v.size == 1 else v.reshape(self.dimension, self.dimension)
        return value

    def translate_value(self, value, function_space=None):
        """Normalise a user value: number -> float, dim-tuple of numbers -> ndarray, Constant -> its value,
        str/Expression -> nodal array, callable(t) in transient -> value(t), per-step list -> entry."""
        if isinstance(value, (tuple, list, np.ndarray)):
            if len(value) == self.dimension and all(isinstance(v, numbers.Number) for v in value):
                values_0 = np.asarray(value, dtype=np.float64)
            elif len(value) == self.dimension and all(isinstance(v, str) for v in value):
                values_0 = Expression(tuple(value), degree=self.settings['fe_degree'])(self.function_space.node_coordinates())
            elif self.transient_settings['transient'] and len(value) > self.dimension:
                values_0 = self.translate_value(value[self.current_step])
            else:
                raise SolverError('{} is supplied, but only tuple of number and string expr of dim = len(v) are supported'.format(type(value)))
        elif isinstance(value, numbers.Number):
            values_0 = float(value)
        elif isinstance(value, Constant):
            v = value.values()
            values_0 = float(v[0]) if v.size == 1 else v
        elif isinstance(value, Function):
            values_0 = value.values
        elif isinstance(value, Expression):
            values_0 = value(self.function_space.node_coordinates())
        elif callable(value) and self.transient_settings['transient']:
            values_0 = self.translate_value(value(self.get_current_time()))
        elif isinstance(value, str):
            if os.path.exists(value):
                values_0 = np.load(value)
            else:
                values_0 = Expression(value, degree=self.settings['fe_degree'])(self.function_space.node_coordinates())
        elif value is None:
            raise TypeError('None type is supplied as value to be translated')
        else:
            raise SolverError('{} is supplied, not tuple, number, Constant, file name, Expression'.format(type(value)))
        return values_0

    def get_variable_name(self):
        if 'scalar_name' in self.settings:
            return self.settings['scalar_name']
        elif 'vector_name' in self.settings:
            return self.settings['vector_name']
        else:
            return 'unknown'

    def get_boundary_variable(self, bc, variable=None):
        if not variable:
            variable = self.get_variable_name()
        bvariable = bc
        if 'values' in bc:
            if isinstance(bc['values'], dict) and variable in bc['values']:
                bvariable = bc['values'][variable]
            if isinstance(bc['values'], list):
                for vbc in bc['values']:
                    if 'variable' in vbc and vbc['variable'] == variable:
                        bvariable = vbc
        return bvariable

    def get_body_source(self):
        if isinstance(self.body_source, (dict)):
            vdict = copy.copy(self.body_source)
            for k in vdict:
                vdict[k] = dict(vdict[k])
                vdict[k]['value'] = self.translate_value(self.body_source[k]['value'])
            return vdict
        else:
            if self.body_source:
                return self.translate_value(self.body_source)
            else:
                return None

    def get_time_step(self, time_iter_):
        try:
            dt = float(self.transient_settings['time_step'])
        except (TypeError, ValueError, KeyError):
            raise SolverError('only a scalar time_step is supported (the reference time_series branch yields dt = 0)')
        return dt

    def get_current_time(self, time_iter_=None):
        # replicated as written (SolverBase.py:453-465): step 0 counts as "not given" and the formula
        # uses (time_iter_ - 1), so callable boundary values see the time one step behind
        if not time_iter_:
            time_iter_ = self.current_step
        dt = float(self.transient_settings['time_step'])
        return self.transient_settings['starting_time'] + dt * (time_iter_ - 1)

    # ------------------------------------------------------------------ time loop
    def init_solver(self):
        self.trial_function = None          # no symbolic layer: the forms are fixed per solver
        self.test_function = None
        self.w_current = self.get_initial_field()
        self.w_prev = Function(self.function_space)
        self.w_prev.assign(self.w_current)
        self.w_pp = Function(self.function_space)
        self.w_pp.assign(self.w_current)

    def solve_current_step(self):
        F, Dirichlet_bcs_up = self.generate_form(self.current_step, self.trial_function, self.test_function, self.w_current, self.w_prev)
        self.w_pp.assign(self.w_prev)
        self.w_prev.assign(self.w_current)      # the form holds references: T_prev is the previous step's solution
        self.w_current = self.solve_form(F, self.w_current, Dirichlet_bcs_up)
        self.result = self.w_current

    def solve_transient(self):
        self.init_solver()
        ts = self.transient_settings
        self.current_time = ts['starting_time']
        self.current_step = 0
        if ts['transient']:
            t_end = ts['ending_time']
        else:
            t_end = self.current_time + 1

        sf = self.report_settings.get('saving_freq')
        result_filename = None
        if sf and sf > 0:
            result_filename = self.report_settings.get('result_filename') or None

        t_start = time.perf_counter()
        while (self.current_time < t_end):
            if ts['transient']:
                dt = self.get_time_step(self.current_step)
            else:
                dt = 1
            self.solve_current_step()
            self.logger.info("Current step = %d time = %g elapsed = %.3f s", self.current_step, self.current_time, time.perf_counter() - t_start)
            pf = self.report_settings.get('plotting_freq', 0)
            if pf and pf > 0 and self.current_step > 0 and (self.current_step % pf == 0):
                self.plot()
            if result_filename and self.current_step > 0 and (self.current_step % sf == 0):
                self.save(result_filename)
            if not self.transient_settings['transient']:
                break
            self.current_step += 1
            self.current_time += dt
        self.timings['solve_all'] = time.perf_counter() - t_start
        return self.w_current

    def solve(self):
        self.result = self.solve_transient()
        return self.result

    def plot(self):
        """Batch-safe: plots only 2D scalar results when matplotlib is importable and plotting is interactive."""
        if not self.report_settings.get('plotting_interactive', False) or os.environ.get('FENICSSOLVER_BATCH') or os.environ.get('BATCH'):
            return
        try:
            import matplotlib.pyplot as plt
            import matplotlib.tri as mtri
        except ImportError:
            return
        if self.dimension == 2 and self.function_space.ncomp == 1:
            c = self.mesh.coordinates()
            plt.tricontourf(mtri.Triangulation(c[:, 0], c[:, 1], self.mesh.cells()), self.result.compute_vertex_values(), 32)
            plt.colorbar()
            plt.show()

    def save(self, result_filename):
        """`File(result_filename) << (w_current, current_time)` (SolverBase.py:570-577): `.pvd` writes the ParaView
        collection plus one ASCII `.vtu` piece per call (`name000000.vtu`, ... as dolfin's VTKFile numbers them);
        also `.vtu` alone, legacy-ASCII `.vtk`, and `.npy` (nodal values, restartable through initial_values)."""
        values = self.w_current.values
        name = self.get_variable_name()
        if result_filename.endswith('.npy'):
            np.save(result_filename, values)
        elif result_filename.endswith('.vtk'):
            write_vtk(result_filename, self.mesh, values, name)
        elif result_filename.endswith('.vtu'):
            write_vtu(result_filename, self.mesh, values, name)
        elif result_filename.endswith('.pvd'):
            series = self.__dict__.setdefault('_pvd_series', {}).setdefault(result_filename, [])
            piece = '%s%06d.vtu' % (result_filename[:-4], len(series))
            write_vtu(piece, self.mesh, values, name)
            series.append((float(getattr(self, 'current_time', 0.0)), os.path.basename(piece)))
            write_pvd(result_filename, series)
        else:
            raise SolverError('result file must end in .pvd, .vtu, .vtk or .npy')

    # ------------------------------------------------------------------ linear solve (the hot path)
    def device_space(self):
        """The device-resident space (mesh + pattern), created on first use and reused by every step."""
        if self._space is None:
            self._space = DeviceSpace(self.mesh, self.function_space.ncomp, comm=self.comm, space=self.function_space)
            self.timings.update({'mesh_upload': self._space.timings['mesh'], 'symbolic': self._space.timings['symbolic']})
        return self._space

    def krylov_parameters(self):
        sp = dict(default_solver_parameters)
        sp.update(self.solver_settings.get('solver_parameters') or {})
        parity = sp.get('parity_mode', True)
        rtol = float(sp.get('relative_tolerance', PARITY_RTOL))
        maxit = int(sp.get('maximum_iterations', PARITY_MAXIT))
        if parity:
            rtol, maxit = min(rtol, PARITY_RTOL), max(maxit, PARITY_MAXIT)
        return {'rtol': rtol, 'atol': float(sp.get('absolute_tolerance', 0.0)), 'maxit': maxit,
                'method': sp.get('linear_solver'), 'precond': sp.get('preconditioner', 'jacobi'),
                'drop_zeros': {True: 1, False: 0, 'auto': 2}.get(sp.get('drop_zeros', 'auto'), 2), 'mg_sweeps': sp.get('mg_sweeps', 2)}

    def solve_linear_problem(self, F, u, Dirichlet_bcs):
        """assemble A and b, apply the Dirichlet conditions, solve (SolverBase.py:592-613).  F is the
        solver's form description with an assemble(space) -> (b, symmetric) method."""
        space = self.device_space()
        t0 = time.perf_counter()
        b, symmetric_form = F.assemble(space)
        kp = self.krylov_parameters()
        method = kp['method'] or ('cg' if symmetric_form else 'bicgstab')
        x = u.device_vector()
        if x is None or x.n != space.ndof_local:
            x = space.vector_from_function(u)
        dofs, vals = collect_dirichlet(Dirichlet_bcs, self.function_space)
        # symmetric elimination (assemble_system) keeps A SPD for CG; plain bc.apply for BiCGStab
        space.apply_dirichlet(b, dofs, vals, symmetric=(method == 'cg'), x=x)
        space.ctx.sync()
        self.timings['assemble'] = time.perf_counter() - t0
        t0 = time.perf_counter()
        # 'drop_zeros' (True / False / 'auto', default 'auto' = when >= 20 % of the stored entries are exactly zero): the Krylov
        # SpMVs skip the entries that are exactly zero after assembly (same products, same iterates; the assembled pattern,
        # which is the parity object, keeps them as dolfin/PETSc do)
        space.ctx.set_option("drop_zeros", int(kp['drop_zeros']))
        if kp['precond'] in ('gmg', 'amg', 'petsc_amg') and method == 'cg':
            # CG preconditioned by geometric multigrid on the nested box meshes: what solve_amg's CG + GAMG is for the
            # reference (SolverBase.py:643-672); the coarse levels are assembled from the same settings
            mg = self.multigrid_hierarchy(space)
            self.timings['mg_setup'] = time.perf_counter() - t0
            info = mg.solve(b, x, rtol=kp['rtol'], atol=kp['atol'], maxit=kp['maxit'], nu=int(kp.get('mg_sweeps', 2)))
            info['mg_levels'] = len(mg.matrices)
        else:
            if kp['precond'] in ('gmg', 'amg', 'petsc_amg'):
                raise SolverError('the multigrid preconditioner needs a symmetric problem (CG)')
            info = space.solve(b, x, method=method, rtol=kp['rtol'], atol=kp['atol'], maxit=kp['maxit'], precond=kp['precond'])
        self.timings['solve'] = time.perf_counter() - t0
        self.solve_info = info
        if info['converged'] != 1:
            self.logger.warning("%s did not converge: %s", method, info)
        self._last_x = x
        if self.parallel and self.solver_settings.get('gather_result', True):
            u.assign_array(space.gather_global(x))      # every rank gets the global vertex-ordered vector
        else:
            # one GPU: the solution stays on the device until someone asks for the array.  Distributed without gather_result:
            # the Function keeps this rank's part (owned + ghosts), which is what the next time step's T_prev term and the
            # Krylov start vector consume; the global array is only available through gather_result / local_result()
            u.set_device(x, local=self.parallel)
        return u

    def multigrid_hierarchy(self, space):
        """Level matrices for the multigrid preconditioner: the fine one is space.A as just assembled; every coarser
        box (cell counts halved while they stay even and >= 4) gets a shallow copy of this solver that marks its
        own boundary facets with the same predicates and assembles the same form there.  The level spaces (meshes,
        patterns) are kept across steps; the values are re-assembled because a transient run changes them."""
        import copy as _copy
        from . import _lib as L
        mesh = self.mesh
        slab_dist = self.parallel and space.part is None
        if not getattr(mesh, 'box', None) or (self.parallel and not slab_dist) or self.function_space.degree != 1:
            raise SolverError("preconditioner 'gmg' needs a generated box mesh (UnitCubeMesh/BoxMesh/...) and degree 1 (one GPU, or z-slabs)")
        n = [int(k) for k in mesh.box['n']]
        sizes = [n]
        while all(k % 2 == 0 and k >= 4 for k in sizes[-1]):
            sizes.append([k // 2 for k in sizes[-1]])
        if slab_dist and len(sizes) < 2:
            raise SolverError("the distributed multigrid preconditioner needs at least one coarser level")
        if slab_dist and mesh.tdim != 3:
            raise SolverError("the distributed multigrid preconditioner is implemented for 3-D boxes (z-slabs); use 'jacobi' for a distributed 2-D run")
        levels = self.__dict__.setdefault('_mg_levels', {})
        mats = [space.A]
        for nl in sizes[1:]:
            key = tuple(nl)
            c = levels.get(key)
            if c is None:
                c = _copy.copy(self)
                # coarse levels live whole on every rank (replicated) even when the fine level is slab-distributed
                c.settings = dict(self.settings)
                c.settings['solver_settings'] = dict(self.settings['solver_settings'], distributed=False)
                c.solver_settings = c.settings['solver_settings']
                c.comm, c.parallel = Comm(), False
                c.mesh = Mesh(box={'n': tuple(nl), 'p0': mesh.box['p0'], 'p1': mesh.box['p1']})
                c.subdomains = None
                c.generate_boundary_facets()
                ncomp = self.function_space.ncomp
                c.function_space = FunctionSpace(c.mesh, 'CG', 1) if ncomp == 1 else VectorFunctionSpace(c.mesh, 'CG', 1)
                c._space = DeviceSpace(c.mesh, ncomp, ctx=space.ctx, space=c.function_space)
                c.timings = {}
                c.__dict__.pop('_mg_levels', None)
                levels[key] = c
            zero = Function(c.function_space)
            Fc, bcs_c = c.generate_form(self.current_step, None, None, zero, zero)
            bc_, _sym = Fc.assemble(c._space)
            dofs, vals = collect_dirichlet(bcs_c, c.function_space)
            c._space.apply_dirichlet(bc_, dofs, np.zeros(len(dofs)), symmetric=True)
            mats.append(c._space.A)
        # the dampings depend on the operator only through lambda_max(D^-1 A) per level, which a re-assembly of the same
        # form on the same meshes (the next time step) does not change: reuse them, estimate once
        prev = self.__dict__.get('_mg_omega')
        reuse = prev if (prev is not None and prev[0] == [tuple(k) for k in sizes]
                         and self.solver_settings.get('solver_parameters', {}).get('mg_reuse_damping', True)) else None
        slab = None
        if slab_dist:
            space.activate()                                 # the fine space's slab layout carries the halo exchanges of the cycle
            layer0 = space.v_off // space.plane
            oz0 = layer0 + space.ghost_lo
            slab = (layer0, oz0, oz0 + space.owned_planes)
        self._mg = None            # release the previous hierarchy's level vectors first: the new one takes exactly those blocks back
        mg = L.Multigrid(space.ctx, mats, sizes, mesh.tdim, omega=None if reuse is None else reuse[1], slab=slab)
        self._mg_omega = ([tuple(k) for k in sizes], [mg.omega(l) for l in range(len(mats))])
        self._mg = mg
        return mg

    def local_result(self):
        """This rank's owned part of the last solution (host copy); with gather_result=False this is the
        only device->host transfer of a distributed solve."""
        return self.device_space().owned_values(self._last_x)

    def solve_amg(self, F, u, bcs):
        """The reference's 3D elasticity path: assemble_system + CG preconditioned by PETSc GAMG (SolverBase.py:643-672).
        Here: symmetric elimination + CG on the 3x3 block matrix, preconditioned by geometric multigrid when the mesh is a
        generated box with at least one coarser level (single GPU, degree 1) and the user named no preconditioner;
        Jacobi otherwise.  Converged to the parity tolerance either way (the reference stops at dolfin's default 1e-6)."""
        sp = self.solver_settings.setdefault('solver_parameters', {})
        auto = 'preconditioner' not in sp
        if auto and self._multigrid_levels() >= 2:
            sp['preconditioner'] = 'gmg'
        try:
            return self.solve_linear_problem(F, u, bcs)
        finally:
            if auto:
                sp.pop('preconditioner', None)

    def _multigrid_levels(self):
        mesh = self.mesh
        if not getattr(mesh, 'box', None) or self.function_space.degree != 1:
            return 0
        if self.parallel and (getattr(mesh, 'force_general_partition', False) or not getattr(mesh, 'slab_partition', True) or mesh.tdim != 3):
            return 0                                          # RCB-partitioned or 2-D slabs: no distributed geometric hierarchy
        n, levels = [int(k) for k in mesh.box['n']], 1
        while all(k % 2 == 0 and k >= 4 for k in n):
            n, levels = [k // 2 for k in n], levels + 1
        return levels

    def solve_nonlinear_problem(self, F, u_current, Dirichlet_bcs, J):
        """Newton's method on the device (NonlinearVariationalSolver, SolverBase.py:615-626).  Each iteration
        assembles the linear part (A, b) and lets the form add its nonlinear Jacobian / residual terms at the
        iterate x; the update solves J dx = -F(x) with homogeneous Dirichlet rows (x carries the boundary values
        from the start).  Stops on ||F|| <= newton_rtol * ||F_0|| (default 1e-11; dolfin's own default is 1e-9) or
        newton_atol.  `J` is unused: the forms know their own derivative."""
        space = self.device_space()
        dist = space.comm.nranks > 1
        kp = self.krylov_parameters()
        sp = self.solver_settings.get('solver_parameters') or {}
        n_rtol, n_atol = float(sp.get('newton_relative_tolerance', 1e-11)), float(sp.get('newton_absolute_tolerance', 1e-14))
        n_maxit = int(sp.get('newton_maximum_iterations', 50))
        dofs, vals = collect_dirichlet(Dirichlet_bcs, self.function_space)
        xg = u_current.array().copy()                         # global nodal values; every rank imposes the same Dirichlet data
        xg[dofs] = np.broadcast_to(vals, np.shape(dofs))
        x = space.vector_from_global(xg)                      # local part (owned + ghosts)
        y, dx = space.scratch_vector('newton_y'), space.scratch_vector('newton_dx')
        t0 = time.perf_counter()
        r0, history, lin_iters = None, [], 0
        for it in range(n_maxit + 1):
            b, symmetric_form = F.assemble(space)                 # linear part: space.A, b
            space.A.spmv(x, y)                                    # refreshes the ghosts of x first when distributed
            b.axpy(-1.0, y)                                       # b <- b - A_lin x
            F.add_newton_terms(space, x, b)                       # A += dR/dx, b -= R(x)
            method = kp['method'] or ('cg' if symmetric_form else 'bicgstab')
            space.apply_dirichlet(b, dofs, np.zeros(len(dofs)), symmetric=(method == 'cg'))
            rn = float(np.sqrt(b.dot(b)))                         # owned rows, all-reduced
            history.append(rn)
            r0 = rn if r0 is None else r0
            if rn <= max(n_rtol * r0, n_atol) or it == n_maxit:
                break
            dx.fill(0.0)
            info = space.solve(b, dx, method=method, rtol=kp['rtol'], atol=kp['atol'], maxit=kp['maxit'], precond=kp['precond'])
            lin_iters += info['iterations']
            if info['converged'] != 1:
                self.logger.warning("%s did not converge inside Newton iteration %d: %s", method, it, info)
            x.axpy(1.0, dx)
            if dist:
                x.halo()                                          # the update is only valid on the owned rows
        self.timings['solve'] = time.perf_counter() - t0
        converged = history[-1] <= max(n_rtol * r0, n_atol)
        self.solve_info = {'iterations': lin_iters, 'converged': int(converged), 'newton_iterations': len(history) - 1,
                           'newton_residuals': history, 'rnorm': history[-1], 'bnorm': r0, 'solve_ms': self.timings['solve'] * 1e3,
                           'spmv_ms': 0.0, 'operand_nnzb': 0}
        if not converged:
            self.logger.warning("Newton did not converge: residuals %s", history)
        self._last_x = x
        if dist and self.solver_settings.get('gather_result', True):
            u_current.assign_array(space.gather_global(x))
        else:
            u_current.set_device(x, local=dist)
        return u_current


class _NodeCoordinates:
    """What DirichletBC.dofs_and_values needs from the node coordinates — their number and rows picked by index — without
    materialising all of them on the host (a generated 256^3 box would cost 0.4 GB and a second of numpy for two faces)."""

    def __init__(self, space):
        self.space = space
        n = space.num_nodes() if hasattr(space, "num_nodes") else space.num_vertices()
        d = space.mesh().gdim if hasattr(space, "mesh") and callable(getattr(space, "mesh")) else space.gdim
        self.shape = (n, d)

    def __getitem__(self, idx):
        if hasattr(self.space, "node_coordinates_at"):
            return self.space.node_coordinates_at(idx)
        return self.space.vertex_coordinates(idx)


def collect_dirichlet(bcs, space):
    """DirichletBC list -> (global dofs, values); later conditions win on shared dofs, as repeated
    bc.apply calls do.  `space`: the FunctionSpace (a Mesh is accepted for P1)."""
    if not bcs:
        return np.zeros(0, dtype=np.int64), np.zeros(0)
    coords = _NodeCoordinates(space)
    dofs_all, vals_all = [], []
    for bc in bcs:
        if not isinstance(bc, DirichletBC):
            raise SolverError('only DirichletBC objects are supported (PointSource is not)')
        d, v = bc.dofs_and_values(coords)
        dofs_all.append(d)
        vals_all.append(v)
    d = np.concatenate(dofs_all)
    v = np.concatenate(vals_all)
    # keep the last occurrence of each dof; the selection depends only on the marked sets, so a transient
    # loop (new DirichletBC objects, same markers, every step) reuses it
    key = tuple((id(bc.markers), bc.marker_id, bc.component, getattr(bc.markers, "version", None)) for bc in bcs)
    cache = getattr(space, "__dict__", {}).setdefault("_bc_merge", {})
    hit = cache.get(key)
    if hit is None or hit[0].size != d.size:
        _, idx = np.unique(d[::-1], return_index=True)
        idx = d.size - 1 - idx
        if len(cache) > 16:
            cache.clear()
        hit = cache[key] = (d, idx, d[idx])
    return hit[2], v[hit[1]]


def mesh_from_dict(m):
    from . import dolfin_compat as dc
    kind = m.get('type')
    n = m.get('n')
    if kind == 'UnitSquareMesh':
        return dc.UnitSquareMesh(*n)
    if kind == 'UnitCubeMesh':
        return dc.UnitCubeMesh(*n)
    if kind == 'RectangleMesh':
        return dc.RectangleMesh(m['p0'], m['p1'], *n)
    if kind == 'BoxMesh':
        return dc.BoxMesh(m['p0'], m['p1'], *n)
    if 'coordinates' in m and 'cells' in m:
        return Mesh(np.asarray(m['coordinates']), np.asarray(m['cells']))
    raise SolverError('unknown mesh description {}'.format(kind))


def write_vtk(path, mesh, values, name):
    c, t = mesh.coordinates(), mesh.cells()
    nv, d = c.shape
    pts = np.zeros((nv, 3))
    pts[:, :d] = c
    with open(path, 'w') as f:
        f.write("# vtk DataFile Version 3.0\n%s\nASCII\nDATASET UNSTRUCTURED_GRID\nPOINTS %d double\n" % (name, nv))
        np.savetxt(f, pts, fmt="%.16g")
        nl = t.shape[1]
        f.write("CELLS %d %d\n" % (t.shape[0], t.shape[0] * (nl + 1)))
        np.savetxt(f, np.hstack([np.full((t.shape[0], 1), nl), t]), fmt="%d")
        f.write("CELL_TYPES %d\n" % t.shape[0])
        np.savetxt(f, np.full(t.shape[0], 10 if nl == 4 else 5), fmt="%d")
        f.write("POINT_DATA %d\n" % nv)
        vals = np.asarray(values)[:nv]                  # degree 2: the vertex nodes come first
        if vals.ndim == 1:
            f.write("SCALARS %s double 1\nLOOKUP_TABLE default\n" % name)
            np.savetxt(f, vals, fmt="%.16g")
        else:
            v3 = np.zeros((nv, 3))
            v3[:, :vals.shape[1]] = vals
            f.write("VECTORS %s double\n" % name)
            np.savetxt(f, v3, fmt="%.16g")


def write_vtu(path, mesh, values, name):
    """One ASCII VTK XML UnstructuredGrid piece, the layout dolfin's VTKFile writes for a P1 function: points,
    connectivity/offsets/types (5 = triangle, 10 = tetrahedron), one PointData array (vectors padded to 3
    components).  Degree-2 results are written at the vertices (dolfin does the same for `File << u`)."""
    c, t = mesh.coordinates(), mesh.cells()
    nv, d = c.shape
    nc, nl = t.shape
    pts = np.zeros((nv, 3))
    pts[:, :d] = c
    vals = np.asarray(values)[:nv]
    with open(path, 'w') as f:
        f.write('<?xml version="1.0"?>\n<VTKFile type="UnstructuredGrid" version="0.1">\n<UnstructuredGrid>\n')
        f.write('<Piece NumberOfPoints="%d" NumberOfCells="%d">\n' % (nv, nc))
        f.write('<Points>\n<DataArray type="Float64" NumberOfComponents="3" format="ascii">\n')
        np.savetxt(f, pts, fmt="%.16g")
        f.write('</DataArray>\n</Points>\n<Cells>\n<DataArray type="UInt32" Name="connectivity" format="ascii">\n')
        np.savetxt(f, t, fmt="%d")
        f.write('</DataArray>\n<DataArray type="UInt32" Name="offsets" format="ascii">\n')
        np.savetxt(f, (np.arange(1, nc + 1) * nl)[None, :], fmt="%d")
        f.write('</DataArray>\n<DataArray type="UInt8" Name="types" format="ascii">\n')
        np.savetxt(f, np.full((1, nc), 10 if nl == 4 else 5), fmt="%d")
        f.write('</DataArray>\n</Cells>\n')
        if vals.ndim == 1:
            f.write('<PointData Scalars="%s">\n<DataArray type="Float64" Name="%s" format="ascii">\n' % (name, name))
            np.savetxt(f, vals, fmt="%.16g")
        else:
            v3 = np.zeros((nv, 3))
            v3[:, :vals.shape[1]] = vals
            f.write('<PointData Vectors="%s">\n<DataArray type="Float64" Name="%s" NumberOfComponents="3" format="ascii">\n' % (name, name))
            np.savetxt(f, v3, fmt="%.16g")
        f.write('</DataArray>\n</PointData>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n')


def write_pvd(path, series):
    """ParaView collection: one DataSet per (time, piece file)."""
    with open(path, 'w') as f:
        f.write('<?xml version="1.0"?>\n<VTKFile type="Collection" version="0.1">\n<Collection>\n')
        for t, piece in series:
            f.write('<DataSet timestep="%.16g" part="0" file="%s" />\n' % (t, piece))
        f.write('</Collection>\n</VTKFile>\n')
