/* fsb.h — C-ABI of the B200-native FEM assemble-and-solve library (libfsb.so).
 *
 * This is the drop-in boundary for the FenicsSolver hot path.  The reference has no FFI of its
 * own: its seam is the Python layer above dolfin, so each entry point below names the dolfin call
 * it replaces and the reference line that makes that call (paths relative to
 * /root/reference/FenicsSolver/).  The Python host (fenicssolver_b200/_lib.py) binds these with
 * ctypes; INTEGRATION.md shows the stub a FenicsSolver maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative fsb_status on failure; nothing throws
 *     across the ABI; fsb_last_error(ctx) returns the message of the last failure on that ctx.
 *   - handles are opaque; the caller owns all host memory (copied before return), the library owns
 *     all device memory until the matching *_destroy.
 *   - one context per GPU, calls on one context are serialised by the caller; kernels run on the
 *     context's stream; a call blocks only where it returns host-visible data.
 *   - DoF numbering is vertex order: scalar dof = v, vector dof = ncomp*v + c.  Matrices are
 *     block-CSR (block size = ncomp) with int64 row_ptr, int32 col_idx sorted ascending,
 *     structural zeros kept; fsb_mat_download_csr returns the equivalent scalar CSR.
 *   - distributed runs (fsb_dist_*): each rank holds a z-slab of a box mesh in natural plane order
 *     [ghost plane | owned planes | ghost plane]; vectors have local length, reductions and SpMV
 *     run over the owned rows, ghosts are refreshed by the halo exchange inside the solvers.
 */
#ifndef FSB_H
#define FSB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fsb_ctx fsb_ctx;
typedef struct fsb_mesh fsb_mesh;
typedef struct fsb_mat fsb_mat;
typedef struct fsb_mg fsb_mg;     /* geometric multigrid hierarchy (box meshes) */
typedef struct fsb_vec fsb_vec;

typedef enum {
  FSB_OK = 0,
  FSB_ERR_CUDA = -1,     /* a CUDA runtime call failed */
  FSB_ERR_ARG = -2,      /* bad argument */
  FSB_ERR_NOMEM = -3,    /* device allocation failed */
  FSB_ERR_BREAKDOWN = -4,/* Krylov breakdown (zero/NaN pivot scalar) */
  FSB_ERR_NCCL = -5,     /* NCCL missing or a collective failed */
  FSB_ERR_STATE = -6     /* object used in the wrong state */
} fsb_status;

/* result of a Krylov solve (PETScKrylovSolver.solve return + monitor, SolverBase.py:663-670) */
typedef struct {
  int32_t iterations;
  int32_t converged;     /* 1 converged, 0 hit maxit, -1 breakdown */
  double rnorm;          /* final ||M^-1 r||_2 (recurrence residual, preconditioned norm) */
  double bnorm;          /* ||M^-1 b||_2 */
  double solve_ms;       /* device time of the iteration loop (CUDA events on the ctx stream) */
  double spmv_ms;        /* accumulated device time of the SpMV launches when profiling is on, else 0 */
  int64_t operand_nnzb;  /* blocks in the matrix the SpMVs ran on (< the assembled count with "drop_zeros") */
} fsb_solve_info;

/* ---- context --------------------------------------------------------------------------------- */
/* stream: a cudaStream_t the caller owns (e.g. torch.cuda.Stream().cuda_stream) or NULL for a
 * library-owned non-blocking stream. */
int fsb_init(int device, void* stream, fsb_ctx** ctx);
void fsb_destroy(fsb_ctx* ctx);
const char* fsb_last_error(fsb_ctx* ctx);
int fsb_sync(fsb_ctx* ctx);
int fsb_device_info(fsb_ctx* ctx, int32_t* sm_count, int64_t* free_bytes, int64_t* total_bytes);
/* tuning/diagnostic switches: "asm_mode" (0 search+atomics, 1 = default: position-map+atomics, 2: row-gather kernels
 * without atomics for the degree-1 scalar forms — bitwise reproducible sums, about 2x slower — position-map+atomics elsewhere),
 * "spmv_mode" (0 TMA-staged tiles, 1 plain row-per-thread), "profile" (0/1), "graph" (0/1),
 * "check_every" (iterations between host convergence polls), "drop_zeros" (the Krylov SpMVs run on a
 * compacted copy without the blocks that are exactly zero after assembly; the assembled CSR, the parity object, is untouched:
 * 0 never, 1 always, 2 = default: when at least 20 % of the stored blocks are exactly zero),
 * "alloc_cache_mb" (bound on released device blocks kept for exact-size reuse; 0 releases them and disables it). */
int fsb_set_option(fsb_ctx* ctx, const char* name, int64_t value);
int64_t fsb_launch_count(fsb_ctx* ctx);   /* kernels launched by this library on ctx so far */

/* ---- mesh: dolfin Mesh / UnitCubeMesh / BoxMesh / RectangleMesh ------------------------------- */
/* Mesh(filename) (SolverBase.py:224): host arrays; cells[ncells][tdim+1] sorted ascending per cell. */
int fsb_mesh_upload(fsb_ctx* ctx, int32_t gdim, int32_t tdim, int64_t nverts, const double* xyz,
                    int64_t ncells, const int32_t* cells, fsb_mesh** mesh);
/* The same for a contiguous part of a larger mesh (a z-slab of a host box mesh in a distributed run): `cells` holds GLOBAL vertex
 * ids, `xyz` the coordinates of the vertices [vertex_offset, vertex_offset + nverts); ids are shifted to local on the device,
 * so the caller can pass slices of its (pinned) global arrays without a host pass over them. */
int fsb_mesh_upload_part(fsb_ctx* ctx, int32_t gdim, int32_t tdim, int64_t nverts, const double* xyz,
                         int64_t ncells, const int32_t* cells, int64_t vertex_offset, fsb_mesh** mesh);
/* UnitSquareMesh/RectangleMesh (tdim 2, diagonal "right") and UnitCubeMesh/BoxMesh (tdim 3)
 * generated on the device in dolfin's layout (examples/test_heat_transfer.py:34,
 * examples/test_linear_elasticity.py:42).  layer0/layer1 select the cell layers [layer0,layer1)
 * along the last axis (a z-slab; pass 0,n[tdim-1] for the whole mesh); vertex ids are local to the
 * slab: global id - layer0*plane_size. */
int fsb_mesh_box(fsb_ctx* ctx, int32_t tdim, const int32_t* n, const double* p0, const double* p1,
                 int32_t layer0, int32_t layer1, fsb_mesh** mesh);
/* FunctionSpace(mesh, "Lagrange", 2) (examples/test_linear_elasticity.py:105-106): the mesh together with
 * a degree-2 node layout.  cell_nodes[ncells][nl], nl = 6 (triangles) / 10 (tetrahedra): the tdim+1 vertices,
 * sorted ascending, then the edge nodes in UFC order (triangle (1,2)(0,2)(0,1); tetrahedron
 * (2,3)(1,3)(1,2)(0,3)(0,2)(0,1)); node ids < nverts are vertices, nnodes = nverts + nedges.  Every entry point
 * below accepts such a mesh: matrices/vectors then have nnodes*ncomp rows, and facet lists hold each facet's P2
 * nodes (its tdim vertices followed by its edge nodes) instead of its vertices only. */
int fsb_mesh_upload_p2(fsb_ctx* ctx, int32_t gdim, int32_t tdim, int64_t nverts, const double* xyz,
                       int64_t ncells, const int32_t* cell_nodes, int64_t nnodes, fsb_mesh** mesh);
int fsb_mesh_sizes(fsb_mesh* mesh, int32_t* gdim, int32_t* tdim, int64_t* nverts, int64_t* ncells);
int fsb_mesh_download(fsb_mesh* mesh, double* xyz, int32_t* cells);
/* Boundary detection and facet numbering (K1): what mesh.init(tdim-1), `facet.exterior()` and dolfin's global facet
 * numbering give FacetFunction / SubDomain.mark / MeshFunction(mesh, "..._facet_region.xml") in SolverBase.py:229,236,277-283.
 * The first call computes, on the device, the facets held by exactly one cell; *nbf = their number, *nfacets = the number
 * of distinct facets of the mesh (either may be NULL).  _get copies them to host arrays (any may be NULL), in lexicographic
 * order of the sorted vertex tuples: fverts[nbf][tdim], opp[nbf] = the cell's vertex opposite the facet, cell[nbf], and
 * facet_id[nbf] = the facet's index in dolfin's numbering (rank of its vertex tuple among all distinct facets). */
int fsb_mesh_exterior_facets(fsb_mesh* mesh, int64_t* nbf, int64_t* nfacets);
int fsb_mesh_exterior_facets_get(fsb_mesh* mesh, int32_t* fverts, int32_t* opp, int32_t* cell, int64_t* facet_id);
/* What SubDomain.mark evaluates `inside` on (SolverBase.py:281-282: all vertices of a facet and its midpoint): *nbv = number of
 * distinct boundary vertices; bverts[nbv] ascending; finv[nbf][tdim] = the facets' vertices as indices into bverts;
 * bxyz[nbv][gdim] their coordinates; mid[nbf][gdim] the facet midpoints.  Call once with NULL arrays for the size. */
int fsb_mesh_boundary_geometry(fsb_mesh* mesh, int64_t* nbv, int32_t* bverts, int32_t* finv, double* bxyz, double* mid);
void fsb_mesh_destroy(fsb_mesh* mesh);

/* ---- vectors: dolfin GenericVector ----------------------------------------------------------- */
int fsb_vec_create(fsb_ctx* ctx, int64_t n, fsb_vec** v);
int fsb_vec_fill(fsb_vec* v, double value);
int fsb_vec_upload(fsb_vec* v, const double* host, int64_t n);
int fsb_vec_download(fsb_vec* v, double* host, int64_t n);
/* Page-locked host memory for results (what vector().get_local() hands to the caller): a block released with fsb_host_free is
 * kept and reused by the next request of the same size; fsb_vec_download into such a block is one direct DMA. */
int fsb_host_alloc(fsb_ctx* ctx, int64_t bytes, void** ptr);
int fsb_host_free(fsb_ctx* ctx, void* ptr);
int fsb_vec_copy(fsb_vec* dst, fsb_vec* src);
int fsb_vec_axpy(fsb_vec* y, double a, fsb_vec* x);            /* y += a x */
/* v[idx[i]] += vals[i] (host lists; repeated indices accumulate): PointSource.apply(b), SolverBase.py:598-602 */
int fsb_vec_add_entries(fsb_vec* v, int64_t n, const int64_t* idx, const double* vals);
int fsb_vec_size(fsb_vec* v, int64_t* n);
void* fsb_vec_ptr(fsb_vec* v);                                 /* device pointer (for interop) */
void fsb_vec_destroy(fsb_vec* v);

/* ---- matrix: DofMap + SparsityPatternBuilder + GenericMatrix ---------------------------------- */
/* builds the topology-based pattern of a P1 space with ncomp components on `mesh` (what dolfin does
 * inside assemble()/LinearVariationalSolver, SolverBase.py:595,608-612,644); values are zero. */
int fsb_mat_create(fsb_mesh* mesh, int32_t ncomp, fsb_mat** A);
/* arbitrary scalar CSR from the host (for SpMV / Krylov use without a mesh). */
int fsb_mat_from_csr(fsb_ctx* ctx, int64_t nrows, const int64_t* row_ptr, const int32_t* col_idx,
                     const double* vals, fsb_mat** A);
int fsb_mat_sizes(fsb_mat* A, int64_t* nrows, int64_t* nnz, int32_t* bs, int64_t* nnzb);
int fsb_mat_download_csr(fsb_mat* A, int64_t* row_ptr, int32_t* col_idx, double* vals);
int fsb_mat_zero(fsb_mat* A);
int fsb_mat_set_owned_rows(fsb_mat* A, int64_t row0, int64_t row1);  /* block rows [row0,row1) this rank owns */
void fsb_mat_destroy(fsb_mat* A);

/* ---- assembly: FFC tabulate_tensor + MatSetValues/VecSetValues ADD_VALUES --------------------- */
/* A += kscale * int (K grad u).grad v  +  mass * int u v  +  adv * int (vel.grad u) v     (P1 scalar)
 * ktensor: gdim*gdim row-major conductivity tensor or NULL for identity; vel: gdim or NULL.
 * Restates ScalarTransportSolver.py:284-285 (F_static), :292 (transient mass), :311 (convection). */
int fsb_assemble_scalar(fsb_mesh* mesh, fsb_mat* A, double kscale, const double* ktensor,
                        double mass, double adv, const double* vel);
/* A = the same form (what `A = assemble(a)` on a fresh tensor is): equal to fsb_mat_zero + fsb_assemble_scalar; the row-gather
 * kernel writes every row once and skips the zero-fill. */
int fsb_assemble_scalar_set(fsb_mesh* mesh, fsb_mat* A, double kscale, const double* ktensor,
                            double mass, double adv, const double* vel);
/* y += (kscale*K + mass*M + adv*C(vel)) x   cell by cell without forming the matrix: the
 * Crank-Nicolson right-hand side (1/dt) c M T_prev - (1-theta) K T_prev, ScalarTransportSolver.py:292-293 */
int fsb_apply_scalar(fsb_mesh* mesh, fsb_vec* x, fsb_vec* y, double kscale, const double* ktensor,
                     double mass, double adv, const double* vel);
/* A += int sigma(u):grad(v), sigma = 2 mu sym(grad u) + lambda div(u) I   (LinearElasticitySolver.py:62-69,215) */
int fsb_assemble_elasticity(fsb_mesh* mesh, fsb_mat* A, double mu, double lambda);
/* b += scale * int S.v dx, S constant[ncomp]  (ScalarTransportSolver.py:213-226; LinearElasticitySolver.py:227-228);
 * cell_tags/tag (host int32[ncells] or NULL): restrict to dx(tag). */
int fsb_assemble_source(fsb_mesh* mesh, fsb_vec* b, int32_t ncomp, const double* S, double scale,
                        const int32_t* cell_tags, int32_t tag);
/* b += scale * int S_h v dx with S_h the P1 interpolant of the nodal vector S (same layout as b). */
int fsb_assemble_source_nodal(fsb_mesh* mesh, fsb_vec* b, int32_t ncomp, fsb_vec* S, double scale);
/* exterior-facet terms over nf facets given by host vertex lists fverts[nf][tdim] (+ opposite vertex
 * opp[nf] when the outward normal is needed):
 *   mode 0: b += scale * int g.v ds            g constant[ncomp]              (ScalarTransportSolver.py:179-208)
 *   mode 1: b += scale * int (g[0] n).v ds     pressure / normal force        (LinearElasticitySolver.py:176-189)  */
int fsb_assemble_facet_load(fsb_mesh* mesh, fsb_vec* b, int32_t ncomp, int64_t nf, const int32_t* fverts,
                            const int32_t* opp, int32_t mode, const double* g, double scale);
/* A += h * int u v ds over the facets: the HTC/Robin matrix term (ScalarTransportSolver.py:201-208). */
int fsb_assemble_facet_mass(fsb_mesh* mesh, fsb_mat* A, int64_t nf, const int32_t* fverts, double h);
/* assemble(Constant(1)*ds(id))  (LinearElasticitySolver.py:171) */
int fsb_facet_area(fsb_mesh* mesh, int64_t nf, const int32_t* fverts, double* area);

/* b += scale * beta * int (T - T_ref) div(v) dx: the thermal-stress load of sigma_t = beta (T - T_ref) I, beta =
 * E/(1-2 nu) * expansion coefficient (LinearElasticitySolver.py:78-85 thermal_stress, :232-238).  T: nodal scalar
 * field on the space's nodes, or NULL for the constant T_const.  b has ncomp == dim. */
int fsb_assemble_thermal_load(fsb_mesh* mesh, fsb_vec* b, double beta, fsb_vec* T, double T_const, double T_ref,
                              double scale);
/* b_a += int vonMises(u) phi_a dx with phi_a the P1 vertex basis: right-hand side of project(von_Mises, P1)
 * (LinearElasticitySolver.py:71-76); u: displacement on the mesh's nodes (ncomp == dim), b: nverts entries. */
int fsb_assemble_von_mises_load(fsb_mesh* mesh, fsb_vec* u, double mu, double lambda, fsb_vec* b);
/* Newton terms of the radiation boundary flux m (Ta^4 - T^4) over the given exterior facets
 * (ScalarTransportSolver.py:334-359 F -= radiation_flux(T)*Tq*ds, :361-374): A += int 4 m T^3 u v ds (A may be NULL),
 * r += rscale * int m (T^4 - Ta^4) v ds (r may be NULL).  Degree-1 spaces. */
int fsb_assemble_facet_radiation(fsb_mesh* mesh, fsb_mat* A, fsb_vec* r, fsb_vec* T, int64_t nf, const int32_t* fverts,
                                 double m, double T_ambient, double rscale);

/* Conductivity as a function of the unknown, k(T) (ScalarTransportSolver.py:228-233; examples/test_heat_transfer.py:53-56),
 * Newton terms at the iterate T with k, dk the nodal values k(T_a), k'(T_a) (k_h = their P1 interpolant):
 * r += rscale * scale * int k_h grad T . grad v,  A += scale * int (k_h grad u + k'_h u grad T) . grad v.  Degree 1. */
int fsb_assemble_scalar_nonlinear_k(fsb_mesh* mesh, fsb_mat* A, fsb_vec* r, fsb_vec* T, fsb_vec* k, fsb_vec* dk,
                                    double scale, double rscale);

/* Convection by a velocity field given at the vertices (dim values per vertex, P1 interpolant v_h; the reference interpolates
 * Expression velocities into a Lagrange space, ScalarTransportSolver.py:130-139): A += scale * int (v_h.grad u) v dx, or
 * y += (that) x when A is NULL.  Degree 1. */
int fsb_assemble_advection_nodal(fsb_mesh* mesh, fsb_mat* A, fsb_vec* x, fsb_vec* y, fsb_vec* vel, double scale);
/* SUPG stabilisation (ScalarTransportSolver.py:252-274, method 2): test function Tq = q + tau vel.grad(q),
 * tau = 0.5 h / (4/(Pe h) + 2 |vel|), h = 2 * cell circumradius.  The three calls add ONLY the extra terms that the
 * tau vel.grad(q) part of the test function produces (degree 1, constant vel); the Galerkin terms come from the calls above.
 *   scalar_supg:  A (or y += .. x when A is NULL) gets  int (mass u + adv vel.grad u) tau vel.grad(v)
 *   source_supg:  b += int S tau vel.grad(v) dx          (cell_tags/tag as in fsb_assemble_source)
 *   facet_supg:   b += int g tau vel.grad(v) ds,  A += int h u tau vel.grad(v) ds   over facets (vertex lists + opposite
 *                 vertex, which together name the adjacent cell whose gradient is used)  */
int fsb_assemble_scalar_supg(fsb_mesh* mesh, fsb_mat* A, fsb_vec* x, fsb_vec* y, double mass, double adv,
                             const double* vel, double pe);
int fsb_assemble_source_supg(fsb_mesh* mesh, fsb_vec* b, double S, const double* vel, double pe,
                             const int32_t* cell_tags, int32_t tag);
int fsb_assemble_facet_supg(fsb_mesh* mesh, fsb_mat* A, fsb_vec* b, int64_t nf, const int32_t* fverts,
                            const int32_t* opp, double g, double h, const double* vel, double pe);

/* ---- DirichletBC.apply / assemble_system ------------------------------------------------------ */
/* symmetric=0: zero row, unit diagonal, b=g (bc.apply(A,b), SolverBase.py:598-602, 608);
 * symmetric=1: additionally b -= A[:,bc] g and zero the column (assemble_system, SolverBase.py:644).
 * x (optional) gets x[dof]=g so a Krylov start vector satisfies the BCs.  dofs/vals are host arrays. */
int fsb_apply_dirichlet(fsb_mat* A, fsb_vec* b, fsb_vec* x, int64_t nbc, const int64_t* dofs,
                        const double* vals, int32_t symmetric);

/* ---- Krylov: PETSc KSPCG / KSPBCGS + PCJacobi -------------------------------------------------- */
int fsb_spmv(fsb_mat* A, fsb_vec* x, fsb_vec* y);               /* y = A x over the owned rows */
int fsb_dot(fsb_vec* x, fsb_vec* y, double* result);            /* owned range, allreduced when distributed */
/* precond: 0 none, 1 Jacobi (M = diag A).  Convergence is tested on the preconditioned residual, as PETSc's
 * KSP does by default: ||M^-1 r||_2 <= max(rtol*||M^-1 b||_2, atol).  x holds the start vector on entry and
 * the solution on exit.  (SolverBase.py:603-612, 663-670) */
int fsb_solve_cg(fsb_mat* A, fsb_vec* b, fsb_vec* x, double rtol, double atol, int32_t maxit,
                 int32_t precond, fsb_solve_info* info);
int fsb_solve_bicgstab(fsb_mat* A, fsb_vec* b, fsb_vec* x, double rtol, double atol, int32_t maxit,
                       int32_t precond, fsb_solve_info* info);

/* ---- multigrid-preconditioned CG: the reference's 3-D elasticity path is CG + PETSc GAMG (SolverBase.py:643-672 solve_amg) --
 * Geometric multigrid on generated box meshes, whose triangulation is nested under halving the cell counts: level 0 is
 * the fine matrix, A[l] the same form assembled (and Dirichlet-eliminated with fsb_apply_dirichlet, so constrained dofs
 * are known) on the box with ncells[l][0..2] = ncells[l-1]/2 cells per axis.  The matrices stay owned by the caller and must
 * outlive the hierarchy.  P1 interpolation along the coarse edges / its transpose, Chebyshev smoothing of degree nu on D^-1 A
 * over [lambda_max/10, lambda_max] (lambda_max from a power iteration capped by the Gershgorin bound, per level; reported
 * through fsb_mg_omega as omega = 4/(3 lambda_max)), damped Jacobi on the coarsest level, V(nu,nu) cycle as the
 * preconditioner of CG; convergence test as fsb_solve_cg.
 * Single GPU. */
int fsb_mg_create(fsb_ctx* ctx, int32_t nlevels, fsb_mat** A, const int32_t* ncells /*[nlevels][3]*/, int32_t tdim,
                  const double* omega /* per-level values to reuse from an earlier hierarchy (entries <= 0 or NULL: estimate) */,
                  fsb_mg** mg);
/* The same hierarchy for a slab-distributed fine level (one process per GPU, fsb_dist_set_slab active): A[0] is this rank's slab of
 * the fine matrix (local plane 0 = global vertex plane layer0, owned planes [owned_z0, owned_z1) of the fine grid), ncells[0..3) the
 * GLOBAL fine cell counts; A[1..] are whole coarse-level matrices, replicated on every rank.  Fine-level smoothing, residuals and the
 * outer CG run distributed (ghost-plane halo, all-reduced dot products); the restricted residual is summed over the ranks by one
 * all-reduce and every rank cycles the coarse levels itself.  What KSPCG + PCGAMG on an MPI communicator are to solve_amg
 * (SolverBase.py:643-672) when the reference runs under mpirun. */
int fsb_mg_create_slab(fsb_ctx* ctx, int32_t nlevels, fsb_mat** A, const int32_t* ncells, int32_t tdim, const double* omega,
                       int32_t layer0, int32_t owned_z0, int32_t owned_z1, fsb_mg** mg);
int fsb_mg_omega(fsb_mg* mg, int32_t level, double* omega);      /* 4 / (3 lambda_max estimate) of a level */
int fsb_mg_apply(fsb_mg* mg, fsb_vec* r, fsb_vec* z, int32_t nu);  /* z = one V(nu,nu) cycle applied to r (zero start) */
int fsb_solve_cg_mg(fsb_mg* mg, fsb_vec* b, fsb_vec* x, double rtol, double atol, int32_t maxit, int32_t nu,
                    fsb_solve_info* info);
void fsb_mg_destroy(fsb_mg* mg);

/* ---- distributed: mesh partition + PETSc VecScatter/MPI_Allreduce (SolverBase.py:102-118, 634) -- */
#define FSB_NCCL_UID_BYTES 128
int fsb_dist_unique_id(void* uid128);                           /* rank 0 creates, host broadcasts */
int fsb_dist_init(fsb_ctx* ctx, int32_t rank, int32_t nranks, const void* uid128);
/* declare the slab layout of vectors on this ctx: ghost planes below/above (0 or 1) and the number
 * of owned vertex planes; the plane size follows from each vector's length; neighbours are
 * rank-1 / rank+1. */
int fsb_dist_set_slab(fsb_ctx* ctx, int32_t ghost_lo, int32_t ghost_hi, int64_t owned_planes);
/* general node partition (unstructured meshes, degree-2 spaces): vectors are numbered [n_owned owned nodes | ghosts],
 * the ghosts owned by neighbour i being the contiguous range [recv_off[i], recv_off[i]+recv_cnt[i]); send_idx holds,
 * per neighbour (send_ptr[i]..send_ptr[i+1]), the owned local nodes that neighbour reads, in the order of its ghost range.
 * Vectors with ncomp values per node are handled by their length.  Replaces a slab declaration on this ctx (the
 * peer-memory CG path needs the slab layout; here the Krylov solvers use NCCL send/recv + all-reduce). */
int fsb_dist_set_halo(fsb_ctx* ctx, int64_t n_owned, int64_t n_local, int32_t nneigh, const int32_t* neigh_rank,
                      const int64_t* send_ptr, const int64_t* send_idx, const int64_t* recv_off, const int64_t* recv_cnt);
int fsb_dist_halo(fsb_vec* v);                                  /* refresh ghost planes of v */
int fsb_dist_allreduce_max(fsb_ctx* ctx, double* value);        /* host scalar, for timing */

#ifdef __cplusplus
}
#endif
#endif /* FSB_H */
