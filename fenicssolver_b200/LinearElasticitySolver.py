"""LinearElasticitySolver — small-strain elasticity on P1 vector spaces (3 dofs per node, interleaved).

Mirrors /root/reference/FenicsSolver/LinearElasticitySolver.py: sigma :62-69, boundary types :99-204,
form :206-245, solve :247-253.

  a(u,v) = int (2 mu sym(grad u) + lambda div(u) I) : grad v dx
  L(v)   = - ( sum_i int t_i.v ds(i) + int f.v dx )

The minus sign is the reference's: it appends the load integrals with `F += item`
(:242-243) and then solves lhs(F) == rhs(F), so tractions and body forces act with the opposite sign
to the scalar solver's convention.  This is reproduced by default (settings['reference_load_sign'] =
True); set it False for the conventional sign.  Out of scope: thermal stress, dynamics, modal
analysis, von Mises projection -> SolverError / not provided.
"""
from __future__ import annotations

import numbers

import numpy as np

from . import _lib
from .SolverBase import SolverBase, SolverError
from .dolfin_compat import Constant, DirichletBC


class ElasticityForm:
    def __init__(self, solver, mu, lmbda):
        self.solver, self.mu, self.lmbda = solver, mu, lmbda
        self.tractions = []      # (marker id, constant vector)
        self.pressures = []      # (marker id, scalar along the outward normal)
        self.body_forces = []    # constant vector | nodal array
        self.load_sign = -1.0

    def assemble(self, space):
        s = self.solver
        A = space.A
        A.zero()
        A.assemble_elasticity(self.mu, self.lmbda)
        b = space.scratch_vector('rhs')
        dim = s.dimension
        for marker, t in self.tractions:
            fv, _ = space.local_facets(*s.boundary_facets.facets(marker))
            _lib.assemble_facet_load(space.dmesh, b, fv, t, ncomp=dim, scale=self.load_sign)
        for marker, p in self.pressures:
            fv, op = space.local_facets(*s.boundary_facets.facets(marker))
            _lib.assemble_facet_load(space.dmesh, b, fv, p, ncomp=dim, scale=self.load_sign, opp=op, normal=True)
        for f in self.body_forces:
            if isinstance(f, np.ndarray) and f.shape == (dim,):
                _lib.assemble_source(space.dmesh, b, f, ncomp=dim, scale=self.load_sign)
            else:
                _lib.assemble_source_nodal(space.dmesh, b, space.vector_from_global(np.asarray(f).reshape(-1)), ncomp=dim, scale=self.load_sign)
        return b, True


class LinearElasticitySolver(SolverBase):
    def __init__(self, case_settings):
        case_settings['vector_name'] = 'displacement'
        SolverBase.__init__(self, case_settings)
        self.solving_modal = False
        self.solving_dynamics = False

    def lame_parameters(self):
        elasticity = self.material['elastic_modulus']
        nu = self.material['poisson_ratio']
        mu = elasticity / (2.0 * (1.0 + nu))
        lmbda = elasticity * nu / ((1.0 + nu) * (1.0 - 2.0 * nu))
        return mu, lmbda

    def get_flux(self, u, mag_vector):
        return mag_vector

    def _vector_constant(self, value, what):
        v = self.translate_value(value)
        if isinstance(v, np.ndarray) and v.shape == (self.dimension,):
            return v
        raise SolverError('{} must be a constant vector on the device path'.format(what))

    def update_boundary_conditions(self, time_iter_, u, v, ds):
        F = ds
        V = self.function_space
        bcs = []
        if 'point_source' in self.settings and self.settings['point_source']:
            raise SolverError('point_source is not implemented (and reads surface_source in the reference, :105-108)')
        if 'surface_source' in self.settings and self.settings['surface_source']:
            raise SolverError('surface_source on all boundaries is not implemented on the device path')

        for name, bc_settings in self.boundary_conditions.items():
            i = bc_settings['boundary_id']
            bc = self.get_boundary_variable(bc_settings)
            btype = bc['type']
            if btype == 'Dirichlet' or btype == 'displacement':
                bv = bc['value']
                if isinstance(bv, (tuple, list)) and len(bv) == self.dimension:
                    # per-component constraint, None = free axis (DirichletBC(V.sub(axis), ...), :122-130)
                    for axis_i, disp in enumerate(bv):
                        if disp is not None:
                            bcs.append(DirichletBC(V, self.translate_value(disp), self.boundary_facets, i, component=axis_i))
                else:
                    bcs.append(DirichletBC(V, self.translate_value(bv), self.boundary_facets, i))
            elif btype == 'force':
                val = bc['value']
                if isinstance(val, Constant):
                    val = val.values() if val.values().size > 1 else float(val)
                if isinstance(val, (tuple, list, np.ndarray)) and len(val) == self.dimension:
                    F.tractions.append((i, np.asarray(val, dtype=np.float64)))
                else:
                    bc_force = float(self.translate_value(val))
                    space = self.device_space()
                    fv, _ = space.local_facets(*self.boundary_facets.facets(i))
                    bc_area = _lib.facet_area(space.dmesh, fv)          # assemble(Constant(1)*ds(id)), :171
                    self.logger.info('boundary area (m2) for force boundary is %g', bc_area)
                    g = bc_force / bc_area
                    if bc.get('direction'):
                        F.tractions.append((i, np.asarray(self.translate_value(bc['direction'])) * g))
                    else:
                        F.pressures.append((i, g))
            elif btype == 'pressure':
                p = float(self.translate_value(bc['value']))
                if bc.get('direction'):
                    F.tractions.append((i, np.asarray(self.translate_value(bc['direction'])) * p))
                else:
                    F.pressures.append((i, p))
            elif btype == 'stress':
                g = self.translate_value(bc['value'])
                if isinstance(g, np.ndarray) and g.shape == (self.dimension,):
                    F.tractions.append((i, g))                      # Constant vector: used as is (:192-193)
                else:
                    raise SolverError('stress tensors (sigma.n) are not implemented on the device path')
            elif btype == 'Neumann':
                raise SolverError('Neumann boundary type`{}` is not supported'.format(btype))
            elif btype == 'symmetry':
                raise SolverError('symmetry boundary type`{}` is not supported'.format(btype))
            else:
                raise SolverError('boundary type`{}` is not supported'.format(btype))
        return bcs, F

    def generate_form(self, time_iter_, u, v, u_current, u_prev):
        mu, lmbda = self.lame_parameters()
        F = ElasticityForm(self, mu, lmbda)
        F.load_sign = -1.0 if self.settings.get('reference_load_sign', True) else 1.0
        if self.transient_settings['transient'] and self.solving_dynamics:
            raise SolverError('elastodynamics is not implemented')
        bcs, F = self.update_boundary_conditions(time_iter_, u, v, F)
        if self.body_source is not None:
            f = self.translate_value(self.body_source)
            F.body_forces.append(f)
        if self.settings.get('temperature_distribution') or getattr(self, 'temperature_distribution', None):
            raise SolverError('thermal stress is not implemented on the device path')
        return F, bcs

    def solve_form(self, F, u_, bcs):
        if self.dimension == 3:
            u_ = self.solve_amg(F, u_, bcs)
        else:
            u_ = self.solve_linear_problem(F, u_, bcs)
        return u_

    def displacement(self):
        return self.w_current

    def velocity(self):
        dt = self.get_time_step(self.current_step)
        return (self.w_current.values - self.w_prev.values) / dt
