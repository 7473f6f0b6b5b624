"""LinearElasticitySolver — small-strain elasticity on P1 vector spaces (3 dofs per node, interleaved).

Mirrors /root/reference/FenicsSolver/LinearElasticitySolver.py: sigma :62-69, boundary types :99-204,
form :206-245, solve :247-253.

  a(u,v) = int (2 mu sym(grad u) + lambda div(u) I) : grad v dx
  L(v)   = - ( sum_i int t_i.v ds(i) + int f.v dx )

The minus sign is the reference's: it appends the load integrals with `F += item`
(:242-243) and then solves lhs(F) == rhs(F), so tractions and body forces act with the opposite sign
to the scalar solver's convention.  This is reproduced by default (settings['reference_load_sign'] =
True); set it False for the conventional sign.  Thermal stress enters with the conventional sign, as in the
reference (`F -= inner(stress_t, grad(v))*dx`, :238).  von_Mises(u) is the L2 projection onto P1 (:71-76).
Out of scope: elastodynamics, modal analysis -> SolverError.
"""
from __future__ import annotations

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
        self.thermal = None      # (beta, T: number | nodal array, T_ref)
        self.stress_tensors = [] # (marker id, dim x dim tensor S): traction S.n per facet
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
            fv, op = space.local_facets(*(s.mesh.exterior_facets()[:2] if marker is None else s.boundary_facets.facets(marker)))
            _lib.assemble_facet_load(space.dmesh, b, fv, p, ncomp=dim, scale=self.load_sign, opp=op, normal=True)
        for f in self.body_forces:
            if isinstance(f, np.ndarray) and f.shape == (dim,):
                _lib.assemble_source(space.dmesh, b, f, ncomp=dim, scale=self.load_sign)
            else:
                _lib.assemble_source_nodal(space.dmesh, b, space.vector_from_global(np.asarray(f).reshape(-1)), ncomp=dim, scale=self.load_sign)
        for marker, S in self.stress_tensors:
            dofs, vals = facet_traction_entries(s.mesh, space.fs, *s.boundary_facets.facets(marker), S)
            ld, lv = space.local_dofs(dofs, self.load_sign * vals)
            b.add_entries(ld, lv)
        if self.thermal is not None:
            beta, T, T_ref = self.thermal
            if isinstance(T, np.ndarray):
                Td = space.local_nodal(T, 1) if space.comm.nranks > 1 else _lib.DeviceVector.from_numpy(space.ctx, T)
                _lib.assemble_thermal_load(space.dmesh, b, beta, T=Td, T_ref=T_ref)
            else:
                _lib.assemble_thermal_load(space.dmesh, b, beta, T_const=float(T), T_ref=T_ref)
        return b, True


def facet_traction_entries(mesh, fs, fverts, opp, S):
    """(dofs, values) of  int (S.n).v ds  over the given exterior facets, n the outward unit normal of each facet
    (LinearElasticitySolver.py:190-196, `dot(g, mesh_normal)`).  Boundary-only work, done on the host and added to the
    right-hand side with fsb_vec_add_entries.  P1: |F|/d per facet vertex; P2: the facet integrals of the degree-2
    basis (segment 1/6, 1/6, 2/3; triangle 0 at the vertices, 1/3 at the edge nodes)."""
    c = mesh.coordinates()
    fverts = np.asarray(fverts, dtype=np.int64)
    d = fverts.shape[1]
    X = c[fverts]
    if d == 2:
        t = X[:, 1] - X[:, 0]
        meas = np.linalg.norm(t, axis=1)
        n = np.stack([t[:, 1], -t[:, 0]], axis=1)
    else:
        n = np.cross(X[:, 1] - X[:, 0], X[:, 2] - X[:, 0])
        meas = 0.5 * np.linalg.norm(n, axis=1)
    n = n / np.linalg.norm(n, axis=1, keepdims=True)
    flip = np.einsum("ij,ij->i", n, c[np.asarray(opp, dtype=np.int64)] - X[:, 0]) > 0
    n[flip] *= -1
    trac = n @ np.asarray(S, dtype=np.float64).T                     # (S.n) per facet
    degree = getattr(fs, "degree", 1)
    if degree == 1:
        nodes, w = fverts, np.full(d, 1.0 / d)
    else:
        nodes = fs.facet_nodes(fverts).astype(np.int64)
        w = np.array([1 / 6, 1 / 6, 2 / 3]) if d == 2 else np.array([0.0, 0.0, 0.0, 1 / 3, 1 / 3, 1 / 3])
    vals = meas[:, None, None] * w[None, :, None] * trac[:, None, :]  # [nf, nodes per facet, dim]
    dofs = nodes[:, :, None] * d + np.arange(d)[None, None, :]
    return dofs.ravel(), vals.ravel()


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

    def thermal_stress(self, T):
        """Isotropic thermal stress magnitude E/(1-2nu) * tec * (T - T_ref) (times the identity), :78-85.
        T: number or nodal array."""
        elasticity = self.material['elastic_modulus']
        nu = self.material['poisson_ratio']
        tec = self.material['thermal_expansion_coefficient']
        return elasticity / (1.0 - 2.0 * nu) * tec * (T - self.reference_values['temperature'])

    def von_Mises(self, u):
        """project(sqrt(3/2 s:s), FunctionSpace(mesh, 'P', 1)), s the deviatoric stress of u (:71-76): the load
        int vm phi_a and the P1 mass matrix are assembled on the device and solved by Jacobi-CG."""
        from .backend import DeviceSpace
        from .dolfin_compat import Function, FunctionSpace
        space = self.device_space()
        dist = space.comm.nranks > 1
        mu, lmbda = self.lame_parameters()
        V1 = FunctionSpace(self.mesh, 'P', 1)
        proj = getattr(self, '_vm_space', None)
        if proj is None:
            # the scalar P1 space of the projection: on the displacement's own device mesh when that is degree 1 on one GPU;
            # distributed, a scalar DeviceSpace over the same partition (same slab / same RCB cut of the same nodes)
            if space.degree == 1 and not dist:
                proj = (space.dmesh, _lib.DeviceMatrix.create(space.dmesh, 1), None)
            else:
                s1 = DeviceSpace(self.mesh, 1, ctx=space.ctx, comm=space.comm)
                proj = (s1.dmesh, s1.A, s1)
            self._vm_space = proj
        dmesh1, M, s1 = proj
        if dist and space.degree != 1:
            raise SolverError('von_Mises projection of a distributed degree-2 displacement is not implemented')
        ud = u.device_vector() if isinstance(u, Function) else None
        if ud is None or ud.n != space.ndof_local:
            ud = space.vector_from_global(u.array() if isinstance(u, Function) else np.asarray(u).reshape(-1))
        elif dist:
            space.activate()
            ud.halo()                                       # the cell loop reads the displacement at the ghost nodes
        nv = s1.nv_local if s1 is not None else self.mesh.num_vertices()
        b = _lib.DeviceVector(space.ctx, nv)
        _lib.assemble_von_mises_load(space.dmesh, ud, mu, lmbda, b)
        M.zero()
        M.assemble_scalar(kscale=0.0, mass=1.0)
        x = _lib.DeviceVector(space.ctx, nv)
        if s1 is not None:
            info = s1.solve(b, x, method='cg', rtol=1e-12, maxit=100000)
        else:
            info = M.solve(b, x, method='cg', rtol=1e-12, maxit=100000)
        if info['converged'] != 1:
            self.logger.warning('von Mises projection did not converge: %s', info)
        out = Function(V1)
        if dist:
            out.assign_array(s1.gather_global(x))           # every rank gets the global vertex-ordered field
        else:
            out.set_device(x)
        return out

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
            # dot(mesh_normal*gS, v)*ds over the whole exterior surface (:110-115); a given 'direction' is read but never
            # used by the reference (the branch adds nothing), which is reproduced
            ss = self.settings['surface_source']
            if not ss.get('direction'):
                F.pressures.append((None, float(self.translate_value(self.get_flux(u, ss['value'])))))

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
                elif isinstance(g, np.ndarray) and g.size == self.dimension ** 2:
                    # a stress tensor: g = dot(g, mesh_normal) (:194-195), one traction per facet
                    F.stress_tensors.append((i, g.reshape(self.dimension, self.dimension)))
                else:
                    raise SolverError('stress must be a constant vector or a constant dim x dim tensor')
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
        # thermal stress (:230-238): settings['temperature_distribution'] or an attribute set by a coupler
        if not hasattr(self, 'temperature_distribution'):
            td = self.settings.get('temperature_distribution')
            if td is not None and not (isinstance(td, (int, float)) and not td):
                self.temperature_distribution = td
        td = getattr(self, 'temperature_distribution', None)
        if td is not None:
            T = self.translate_value(td, self.function_space)
            if isinstance(T, np.ndarray):
                T = np.asarray(T, dtype=np.float64).reshape(-1)
                if T.size != self.function_space.num_nodes():
                    raise SolverError('temperature_distribution must be a number or one value per node of the space')
            beta = self.thermal_stress(1.0 + self.reference_values['temperature'])      # E/(1-2nu)*tec
            F.thermal = (beta, T, float(self.reference_values['temperature']))
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
