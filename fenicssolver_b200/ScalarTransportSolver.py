"""ScalarTransportSolver — heat / species / electrostatics: diffusion (+ advection) with Dirichlet,
flux, Neumann, HTC/Robin boundaries, steady or Crank-Nicolson transient.

Mirrors /root/reference/FenicsSolver/ScalarTransportSolver.py: material look-ups :73-129, boundary
types :142-211, body source :213-226, weak form :228-311, linear solve :378-383.  The UFL form is
replaced by a ScalarForm record that libfsb assembles:

  steady     A = K(k) + c C(v) + sum_i h_i M_F(i)
             b = sum_i int g_i q ds(i) + sum_i h_i Ta_i int q ds(i) + int S q dx
  transient  A = (c/dt) M + theta K(k) + c C(v) + sum h M_F        theta = 0.5 (Crank-Nicolson)
             b = (c/dt) M T_prev - (1-theta) K(k) T_prev + loads
(convection and the boundary/source loads are fully implicit, exactly as the reference writes them,
:297-311).  Radiation (:334-374) adds the nonlinear boundary term -m (Ta^4 - T^4) q ds over the whole exterior
surface; the problem is then solved by Newton's method (SolverBase.solve_nonlinear_problem), the Jacobian term
4 m T^3 u q ds and the residual being assembled on the device.  Out of scope on the device path: nonlinear
material functions, SUPG/IP stabilisation, point sources -> SolverError.
"""
from __future__ import annotations

import numbers

import numpy as np

from . import _lib
from .SolverBase import SolverBase, SolverError
from .dolfin_compat import Constant, DirichletBC, Expression, Function, Point, PointSource

supported_scalars = {'temperature', 'electric_potential', 'species_concentration'}
electric_permittivity_in_vacumm = 8.854187817e-12


class ScalarForm:
    """Plain-data description of the linear scalar transport form; assemble() runs it on the device."""

    def __init__(self, solver):
        self.solver = solver
        self.conductivity = 1.0          # number or dim x dim tensor
        self.capacity = 1.0
        self.transient = False
        self.dt = None
        self.theta = 0.5
        self.velocity = None             # constant vector, or nodal field [nverts, dim] (P1 interpolant)
        self.neumann = []                # (marker id, constant g)
        self.robin = []                  # (marker id, h, T_ambient)
        self.sources = []                # (value (number | nodal array), subdomain id | None)
        self.T_prev = None
        self.radiation = None            # (m = emissivity * Stefan-Boltzmann, T_ambient)
        self.point_sources = []          # PointSource objects
        self.conductivity_fn = None      # k(T) callable: the stiffness term moves into add_newton_terms
        self.conductivity_nodal = None   # k(x) at the vertices (Expression / array): P1 interpolant, own kernel
        self.supg_pe = None              # Peclet number of the SUPG test function q + tau v.grad q (None: Galerkin)

    def _k(self):
        k = self.conductivity
        if isinstance(k, np.ndarray) and k.ndim == 2:
            return 1.0, k
        return float(k), None

    def assemble(self, space):
        """-> (rhs DeviceVector, matrix_is_symmetric).  A is assembled into space.A."""
        s = self.solver
        A = space.A
        kscale, ktensor = self._k()
        c = float(self.capacity)
        vel = None if self.velocity is None else np.asarray(self.velocity, dtype=np.float64)
        vel_field = None
        if vel is not None and vel.ndim == 2:
            # a velocity FIELD: the convection matrix comes from its own kernel, nothing else sees the velocity
            if self.supg_pe:
                raise SolverError('SUPG with a velocity field is not implemented (constant velocity only)')
            vel_field, vel = space.local_nodal(vel, vel.shape[1]), None       # this rank's nodes (owned + ghosts) of the global field
        adv = c if vel is not None else 0.0
        b = space.scratch_vector('rhs')
        supg = self.supg_pe if vel is not None else None
        if self.transient:
            A.assemble_scalar(kscale=self.theta * kscale, ktensor=ktensor, mass=c / self.dt, adv=adv, vel=vel, overwrite=True)
            tp = self.T_prev.device_vector()
            if tp is None or tp.n != space.ndof_local:
                tp = space.vector_from_function(self.T_prev)
            elif space.comm.nranks > 1:
                tp.halo()
            _lib.apply_scalar(space.dmesh, tp, b, kscale=-(1.0 - self.theta) * kscale, ktensor=ktensor, mass=c / self.dt)
            if supg:
                _lib.assemble_scalar_supg(space.dmesh, A, vel, supg, mass=c / self.dt, adv=adv)
                _lib.assemble_scalar_supg(space.dmesh, None, vel, supg, mass=c / self.dt, x=tp, y=b)
        else:
            A.assemble_scalar(kscale=kscale, ktensor=ktensor, adv=adv, vel=vel, overwrite=True)      # A = form: first term after what was A.zero()
            if supg:
                _lib.assemble_scalar_supg(space.dmesh, A, vel, supg, adv=adv)
        if vel_field is not None:
            _lib.assemble_advection_nodal(space.dmesh, A, vel_field, scale=c)      # fully implicit, like the constant case (:311)
        if self.conductivity_nodal is not None:
            # k(x): int k_h grad u . grad v with k_h the P1 interpolant of the nodal values; the Newton kernel with k' = 0 is
            # exactly that bilinear form (and its action on T_prev for the explicit half of Crank-Nicolson)
            kd = space.vector_from_global(self.conductivity_nodal)
            zero = space.scratch_vector('zero')
            _lib.assemble_scalar_nonlinear_k(space.dmesh, A, None, zero, kd, zero, scale=self.theta if self.transient else 1.0)
            if self.transient:
                _lib.assemble_scalar_nonlinear_k(space.dmesh, None, b, tp, kd, zero, scale=1.0 - self.theta, rscale=-1.0)
        for marker, g in self.neumann:
            fv, op = space.local_facets(*s.boundary_facets.facets(marker))
            _lib.assemble_facet_load(space.dmesh, b, fv, g)
            if supg:
                _lib.assemble_facet_supg(space.dmesh, None, b, fv, op, vel, supg, g=g)
        for marker, h, Ta in self.robin:
            fv, op = space.local_facets(*s.boundary_facets.facets(marker))
            A.assemble_facet_mass(fv, h)
            _lib.assemble_facet_load(space.dmesh, b, fv, h * Ta)
            if supg:
                _lib.assemble_facet_supg(space.dmesh, A, b, fv, op, vel, supg, g=h * Ta, h=h)
        for value, sub_id in self.sources:
            if isinstance(value, np.ndarray):
                if sub_id is not None:
                    raise SolverError('nodal body source restricted to a subdomain is not implemented')
                if supg:
                    raise SolverError('SUPG with a nodal (Expression) body source is not implemented')
                _lib.assemble_source_nodal(space.dmesh, b, space.vector_from_global(value))
            else:
                tags = None
                if sub_id is not None:
                    if s.subdomains is None:
                        raise SolverError('body_source per subdomain needs cell markers (mesh_physical_region.xml)')
                    tags = space.local_cell_tags(s.subdomains.array())
                _lib.assemble_source(space.dmesh, b, float(value), cell_tags=tags, tag=sub_id or 0)
                if supg:
                    _lib.assemble_source_supg(space.dmesh, b, float(value), vel, supg, cell_tags=tags, tag=sub_id or 0)
        for ps in self.point_sources:
            nodes, w = ps.entries()
            if space.comm.nranks > 1:
                nodes, w = space.local_dofs(nodes, w)       # every rank adds the entries of its own rows (owner computes)
            b.add_entries(nodes, w)
        symmetric = vel is None and vel_field is None and (ktensor is None or np.allclose(ktensor, ktensor.T, rtol=0, atol=0))
        return b, symmetric and self.conductivity_fn is None       # the k'(T) Jacobian term is not symmetric

    def add_newton_terms(self, space, x, r):
        """Nonlinear part at the iterate x: space.A += dR/dT(x), r -= R(x) (r holds b - A_lin x)."""
        if self.radiation is not None:
            m, Ta = self.radiation
            fv, _ = space.local_facets(*self.solver.mesh.exterior_facets()[:2])
            _lib.assemble_facet_radiation(space.dmesh, space.A, r, x, fv, m, Ta, rscale=-1.0)
        if self.conductivity_fn is not None:
            Th = x.numpy()
            k, dk = nodal_function_and_derivative(self.conductivity_fn, Th)
            _lib.assemble_scalar_nonlinear_k(space.dmesh, space.A, r, x, _lib.DeviceVector.from_numpy(space.ctx, k),
                                             _lib.DeviceVector.from_numpy(space.ctx, dk), rscale=-1.0)


def nodal_function_and_derivative(fn, T):
    """k(T_a) and k'(T_a) at the nodal values.  Central differences, replaced by the complex-step derivative (exact to
    rounding for the arithmetic expressions material lambdas are made of) whenever the function accepts complex input
    and the two agree; a non-analytic function (abs, where, ...) keeps the central differences."""
    T = np.asarray(T, dtype=np.float64)
    k = np.broadcast_to(np.asarray(fn(T), dtype=np.float64), T.shape).copy()
    h = 1e-6 * np.maximum(np.abs(T), 1.0)
    dk = np.broadcast_to((np.asarray(fn(T + h), dtype=np.float64) - np.asarray(fn(T - h), dtype=np.float64)) / (2 * h), T.shape)
    try:
        cs = np.imag(np.broadcast_to(np.asarray(fn(T + 1e-30j)), T.shape)) / 1e-30
        if np.all(np.isfinite(cs)) and np.allclose(cs, dk, rtol=1e-5, atol=1e-7 * max(float(np.abs(dk).max()), 1e-300)):
            dk = cs
    except (TypeError, ValueError):
        pass
    return k, np.ascontiguousarray(dk, dtype=np.float64)


class ScalarTransportSolver(SolverBase):
    """general scalar transport (diffusion and advection) solver, exampled by heat transfer"""

    def __init__(self, s):
        SolverBase.__init__(self, s)
        if 'scalar_name' in self.settings:
            self.scalar_name = self.settings['scalar_name'].lower()
        else:
            self.scalar_name = "temperature"
        self.using_diffusion_form = False
        self.nonlinear = False
        self.nonlinear_material = False
        for v in self.material.values():
            if callable(v) and not isinstance(v, (Constant, Expression, Function)):
                self.nonlinear = True            # as the reference does at construction (:60-66); fields k(x) are linear

    def _material_number(self, c, T):
        from inspect import isfunction
        if isfunction(c):
            # a function of the unknown (:77-78): evaluated at nodal values when T is an array, else returned as is
            self.nonlinear_material = True
            return c(T) if isinstance(T, np.ndarray) else c
        return self.get_material_value(c)

    def capacity(self, T=None):
        if 'capacity' in self.material:
            c = self.material['capacity']
        elif self.scalar_name == "temperature":
            c = self.material['density'] * self.material['specific_heat_capacity']
        elif self.scalar_name == "electric_potential":
            c = electric_permittivity_in_vacumm
        elif self.scalar_name == "spicies_concentration":      # sic: the reference's spelling (:83)
            c = 1
        else:
            raise SolverError('material capacity property is not found for {}'.format(self.scalar_name))
        return self._material_number(c, T)

    def diffusivity(self, T=None):
        if 'diffusivity' in self.material:
            c = self.material['diffusivity']
        elif self.scalar_name == "temperature":
            c = self.material['thermal_conductivity'] / self.capacity()
        elif self.scalar_name == "electric_potential":
            c = self.material['relative_electric_permittivity']
        else:
            raise SolverError('conductivity material property is not found for {}'.format(self.scalar_name))
        return self._material_number(c, T)

    def conductivity(self, T=None):
        if 'conductivity' in self.material:
            c = self.material['conductivity']
        elif self.scalar_name == "temperature":
            c = self.material['thermal_conductivity']
        elif self.scalar_name == "electric_potential":
            c = self.material['relative_electric_permittivity'] * electric_permittivity_in_vacumm
        elif self.scalar_name == "spicies_concentration":
            c = self.material['diffusivity']
        else:
            c = self.diffusivity() * self.capacity()
        return self._material_number(c, T)

    def _constant(self, value, what):
        v = self.translate_value(value)
        if isinstance(v, numbers.Number):
            return float(v)
        raise SolverError('{} must be a constant on the device path (got {})'.format(what, type(value)))

    def update_boundary_conditions(self, time_iter_, T, Tq, ds):
        """-> (DirichletBC list, ScalarForm with the boundary integrals filled in); `ds` is the form."""
        F = ds
        capacity = self.capacity(T)
        bcs = []
        if 'point_source' in self.settings and self.settings['point_source']:
            ps = self.settings['point_source']
            # a PointSource, or a list of (point, magnitude) tuples (:150-158); they go into the right-hand side
            # before the Dirichlet rows are imposed, the order of the reference's bcs list
            if isinstance(ps, PointSource):
                F.point_sources.append(ps)
            else:
                for si in ps:
                    pt = si[0] if isinstance(si[0], Point) else Point(*np.atleast_1d(si[0]))
                    F.point_sources.append(PointSource(self.function_space, pt, si[1]))
        if 'surface_source' in self.settings and self.settings['surface_source']:
            raise SolverError('surface_source is broken in the reference scalar solver (undefined get_flux) and not implemented')

        for name, bc_settings in self.boundary_conditions.items():
            i = bc_settings['boundary_id']
            bc = self.get_boundary_variable(bc_settings)
            btype = bc['type']
            if btype == 'Dirichlet' or btype == 'fixedValue':
                if not isinstance(bc['value'], DirichletBC):
                    T_bc = self.translate_value(bc['value'])
                    bcs.append(DirichletBC(self.function_space, T_bc, self.boundary_facets, i))
                else:
                    bcs.append(bc['value'])
            elif btype == 'Neumann' or btype == 'fixedGradient':
                g = self._constant(bc['value'], 'Neumann gradient')
                F.neumann.append((i, g if self.using_diffusion_form else capacity * g))     # :181 as written
            elif btype == 'symmetry':
                pass
            elif btype == 'mixed' or btype == 'Robin':
                T_bc = self.translate_value(bc['value'])
                g = self._constant(bc['gradient'], 'Robin gradient')
                F.neumann.append((i, g if self.using_diffusion_form else capacity * g))
                bcs.append(DirichletBC(self.function_space, T_bc, self.boundary_facets, i))
            elif btype.lower().find('flux') >= 0 or btype == 'electric_current':
                g = self._constant(bc['value'], 'flux')
                F.neumann.append((i, g / capacity if self.using_diffusion_form else g))
            elif btype == 'HTC':
                Ta = self._constant(bc['ambient'], 'HTC ambient')
                htc = self._constant(bc['value'], 'HTC coefficient')
                if self.using_diffusion_form:
                    htc = htc / capacity
                F.robin.append((i, htc, Ta))
            else:
                raise SolverError('boundary type`{}` is not supported'.format(btype))
        return bcs, F

    def get_body_source_items(self, time_iter_, T, Tq, dx):
        bs = self.get_body_source()
        if bs is not None and isinstance(bs, dict):
            return [(v['value'], v['subdomain_id']) for k, v in bs.items()]
        elif bs is not None:
            return [(bs, None)]
        return None

    def generate_form(self, time_iter_, T, T_test, T_current, T_prev):
        F = ScalarForm(self)
        conductivity = self.conductivity(T)
        capacity = self.capacity(T)
        if callable(capacity):
            raise SolverError('nonlinear capacity is not supported (the reference says so too, :291)')
        if isinstance(conductivity, (Function, Expression)) or (isinstance(conductivity, np.ndarray) and conductivity.ndim == 1):
            # k(x): a scalar field given by an Expression, a Function or one value per vertex
            if self.function_space.degree != 1:
                raise SolverError('a conductivity field is implemented for degree-1 spaces')
            kn = self.translate_value(conductivity) if not isinstance(conductivity, np.ndarray) else conductivity
            kn = np.asarray(kn, dtype=np.float64)
            if kn.shape != (self.mesh.num_vertices(),):
                raise SolverError('a conductivity field must be scalar, one value per vertex (tensor fields are not implemented)')
            F.conductivity_nodal, conductivity = kn, 0.0
        elif callable(conductivity) and not isinstance(conductivity, Constant):
            # k(T): Newton on the device with k_h the P1 interpolant of the nodal values k(T_a)
            if self.function_space.degree != 1:
                raise SolverError('temperature-dependent conductivity is implemented for degree-1 spaces')
            if self.transient_settings['transient']:
                raise SolverError('temperature-dependent conductivity is implemented for steady problems')
            self.nonlinear = True
            F.conductivity_fn, conductivity = conductivity, 0.0
        F.conductivity, F.capacity = conductivity, capacity

        if not hasattr(self, 'convective_velocity'):
            if 'convective_velocity' in self.settings and self.settings['convective_velocity'] is not None:
                self.convective_velocity = self.settings['convective_velocity']
            else:
                self.convective_velocity = None
        if self.convective_velocity is not None:
            ads = self.settings.get('advection_settings') or {'stabilization_method': None}
            vel = self.translate_value(self.convective_velocity)
            nv = self.mesh.num_vertices()
            if isinstance(vel, np.ndarray) and vel.shape == (self.dimension,):
                pass                                           # Constant((vx, vy[, vz]))
            elif isinstance(vel, np.ndarray) and vel.size == nv * self.dimension and self.function_space.degree == 1:
                vel = vel.reshape(nv, self.dimension)          # Expression / Function / nodal array: P1 interpolant of the field
            else:
                raise SolverError('convective_velocity must be a constant vector or a field with dim values per vertex (degree 1)')
            F.velocity = vel
            method = ads.get('stabilization_method')
            if method in ('SPUG', 'SUPG'):               # 'SPUG' is the reference's spelling (:259)
                # SPUG_method == 2 (:268-270): the test function becomes q + tau v.grad(q) in every integral
                if self.function_space.degree != 1:
                    raise SolverError('SUPG is implemented for degree-1 spaces')
                F.supg_pe = float(ads['Pe'])
            elif method:
                raise SolverError('advection stabilisation `{}` is not implemented on the device path (interior penalty '
                                  'needs interior-facet assembly)'.format(method))

        if self.transient_settings['transient']:
            F.transient = True
            F.dt = self.get_time_step(time_iter_)
            F.theta = 0.5
            F.T_prev = T_prev

        bcs, F = self.update_boundary_conditions(time_iter_, T, T_test, F)
        bs_items = self.get_body_source_items(time_iter_, T, T_test, None)
        if bs_items:
            F.sources.extend(bs_items)

        if self.scalar_name == "temperature":
            if ('radiation_settings' in self.settings and self.settings['radiation_settings']):
                self.radiation_settings = self.settings['radiation_settings']
                self.has_radiation = True
            elif hasattr(self, 'radiation_settings') and self.radiation_settings:
                self.has_radiation = True
            else:
                self.has_radiation = False
            if self.has_radiation:
                if self.function_space.degree != 1:
                    raise SolverError('radiation is implemented for degree-1 spaces')
                if F.supg_pe:
                    raise SolverError('radiation together with SUPG is not implemented')
                self.nonlinear = True
                m_, Ta = self.radiation_coefficients()
                F.radiation = (m_, Ta)           # F -= radiation_flux(T)*Tq*ds: all exterior facets (:359)
        return F, bcs

    def radiation_coefficients(self):
        """(emissivity * Stefan constant, ambient temperature) with the reference's look-up order (:361-374)."""
        Stefan_constant = 5.670367e-8
        if 'emissivity' in self.material:
            emissivity = self.material['emissivity']
        elif 'emissivity' in self.radiation_settings:
            emissivity = self.radiation_settings['emissivity']
        else:
            emissivity = 1.0
        if 'ambient_temperature' in self.radiation_settings:
            T_ambient_radiaton = self.radiation_settings['ambient_temperature']
        else:
            T_ambient_radiaton = self.reference_values['temperature']
        return float(emissivity) * Stefan_constant, float(T_ambient_radiaton)

    def radiation_flux(self, T):
        m_, Ta = self.radiation_coefficients()
        return m_ * (Ta ** 4 - np.asarray(T, dtype=np.float64) ** 4)

    def solve_form(self, F, T_current, bcs):
        if self.nonlinear:
            return self.solve_nonlinear_problem(F, T_current, bcs, None)
        else:
            return self.solve_linear_problem(F, T_current, bcs)
