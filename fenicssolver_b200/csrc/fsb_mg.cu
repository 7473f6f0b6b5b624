// Geometric multigrid preconditioner for generated box meshes (the reference's 3-D elasticity path is CG preconditioned by
// algebraic multigrid, SolverBase.py:643-672 solve_amg; every BASELINE config runs on UnitCubeMesh/BoxMesh).
//
// dolfin's box triangulation (6 Kuhn tetrahedra per hexahedron, 2 triangles per rectangle) is invariant under uniform
// refinement: the mesh with n/2 cells per axis is nested in the mesh with n, so the P1 spaces are nested and the
// prolongation is plain P1 interpolation along the coarse edges.  A fine vertex with index parity d = (i&1, j&1, k&1) is
// the coarse vertex C = (i,j,k)/2 when d = 0 and otherwise the midpoint of the coarse edge (C, C + d), which is always an
// edge of the Kuhn triangulation:
//     x_f(2C + d) = x_c(C)                       d = 0
//                 = (x_c(C) + x_c(C + d)) / 2    d != 0
// and the restriction is its transpose, gathered per coarse vertex.  The level operators are re-discretisations (the host
// assembles the same form on every level; for constant coefficients they equal the Galerkin products), Dirichlet rows
// are identity on every level and the transfer operators are masked there.  Smoother: Chebyshev iteration of degree nu on
// D^-1 A over [lambda_max / 10, lambda_max] (what PETSc's GAMG uses by default; ~20 % fewer PCG iterations than nu damped-Jacobi
// sweeps at the same SpMV count), lambda_max from a power iteration capped by the Gershgorin bound per level; the same
// polynomial before and after the coarse correction, so the V-cycle is a symmetric positive operator and plain PCG applies;
// damped-Jacobi sweeps on the coarsest level.  All matrix products are the library's SpMV.
#include "fsb_internal.cuh"
#include <cmath>

namespace {

struct Level {
  fsb_mat* A = nullptr;
  fsb_mat* S = nullptr;         // SpMV operand: A, or its compacted copy without the exactly-zero blocks (drop_zeros)
  int dims[3] = {1, 1, 1};      // vertices per axis
  int64_t nnode = 0, n = 0;     // nodes, scalar dofs
  double omega = 0.6;
  double *dinv = nullptr, *x = nullptr, *b = nullptr, *r = nullptr, *y = nullptr;
  int64_t off = 0, len = 0;     // scalar dofs the element-wise kernels touch: all of them, or (slab-distributed fine level) the owned rows
};

__global__ void k_mg_dinv(int64_t n, int bs, const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx,
                          const double* __restrict__ vals, double* __restrict__ dinv) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t R = r / bs;
    const int i = (int)(r % bs);
    const int64_t base = row_ptr[R];
    const int len = (int)(row_ptr[R + 1] - base);
    const int p = row_find(col_idx + base, 0, len, (int32_t)R);
    const double d = vals[(base + p) * bs * bs + i * bs + i];
    dinv[r] = d != 0.0 ? 1.0 / d : 1.0;
  }
}

// x = w dinv b
__global__ void k_mg_jacobi0(int64_t n, double w, const double* __restrict__ dinv, const double* __restrict__ b, double* __restrict__ x) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] = w * dinv[i] * b[i];
}
// x += w dinv (b - y), y = A x
__global__ void k_mg_jacobi(int64_t n, double w, const double* __restrict__ dinv, const double* __restrict__ b, const double* __restrict__ y,
                            double* __restrict__ x) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] += w * dinv[i] * (b[i] - y[i]);
}
// Chebyshev step after y = A x:  r = b - y ; d = c1 d + c2 dinv r ; x += d     (first step: c1 = 0)
__global__ void k_mg_cheb(int64_t n, double c1, double c2, const double* __restrict__ dinv, const double* __restrict__ b,
                          const double* __restrict__ y, double* __restrict__ d, double* __restrict__ x) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double di = (c1 != 0.0 ? c1 * d[i] : 0.0) + c2 * dinv[i] * (b[i] - y[i]);
    d[i] = di;
    x[i] += di;
  }
}
// first Chebyshev step from a zero iterate: d = x = c2 dinv b
__global__ void k_mg_cheb0(int64_t n, double c2, const double* __restrict__ dinv, const double* __restrict__ b, double* __restrict__ d,
                           double* __restrict__ x) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double di = c2 * dinv[i] * b[i];
    d[i] = di;
    x[i] = di;
  }
}
__global__ void k_mg_residual(int64_t n, const double* __restrict__ b, const double* __restrict__ y, double* __restrict__ r) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) r[i] = b[i] - y[i];
}

// start vector of the power iteration: pseudo-random signs (a smooth start has almost no component along the
// oscillatory eigenvectors that carry lambda_max)
__global__ void k_mg_randsign(int64_t n, int64_t gbase, double* __restrict__ x) {      // signs by GLOBAL dof index: the same start vector on any partition
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    unsigned long long h = (unsigned long long)(i + gbase) * 0x9E3779B97F4A7C15ull;
    h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
    x[i] = (h & 1) ? 1.0 : -1.0;
  }
}

// Gershgorin bound on lambda_max(D^-1 A): max over scalar rows of sum_j |a_ij| / |a_ii| (positive doubles compare like
// their bit patterns, so atomicMax on the bits works)
__global__ void k_mg_gershgorin(int64_t r0, int64_t n, int bs, const int64_t* __restrict__ row_ptr, const double* __restrict__ vals,
                                const double* __restrict__ dinv, unsigned long long* __restrict__ out) {
  double m = 0.0;
  for (int64_t r = r0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < r0 + n; r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t R = r / bs;
    const int i = (int)(r % bs);
    double s = 0.0;
    for (int64_t k = row_ptr[R]; k < row_ptr[R + 1]; ++k)
      for (int j = 0; j < bs; ++j) s += fabs(vals[k * bs * bs + i * bs + j]);
    m = fmax(m, s * fabs(dinv[r]));
  }
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}

struct Dims { int f[3], c[3]; };

// b_c(C) = r_f(2C) + 1/2 sum_{d != 0} ( r_f(2C + d) + r_f(2C - d) ), zero on constrained coarse dofs.
// Slab form: only the coarse planes [ck0, ck1) are produced (those whose fine plane 2K this rank owns) and r_f is this rank's slab,
// whose local plane 0 is the global fine plane fz0 (ghost planes refreshed by the caller); ck0 = 0, ck1 = all, fz0 = 0 otherwise.
__global__ void k_mg_restrict(Dims g, int bs, const double* __restrict__ rf, const uint8_t* __restrict__ bc_c, double* __restrict__ bc_out,
                              int ck0, int ck1, int fz0) {
  const int64_t cplane = (int64_t)g.c[0] * g.c[1];
  const int64_t t0 = cplane * ck0 * bs, t1 = cplane * ck1 * bs;
  for (int64_t t = t0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < t1; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t node = t / bs;
    const int comp = (int)(t % bs);
    if (bc_c && bc_c[t]) { bc_out[t] = 0.0; continue; }
    const int I = (int)(node % g.c[0]), J = (int)((node / g.c[0]) % g.c[1]), K = (int)(node / ((int64_t)g.c[0] * g.c[1]));
    const int fi = 2 * I, fj = 2 * J, fk = 2 * K;
    auto at = [&](int i, int j, int k) -> double {
      if (i < 0 || j < 0 || k < 0 || i >= g.f[0] || j >= g.f[1] || k >= g.f[2]) return 0.0;
      return rf[(i + (int64_t)g.f[0] * (j + (int64_t)g.f[1] * (k - fz0))) * bs + comp];
    };
    double s = at(fi, fj, fk);
    for (int d = 1; d < 8; ++d) {
      const int di = d & 1, dj = (d >> 1) & 1, dk = (d >> 2) & 1;
      if ((dj && g.f[1] == 1) || (dk && g.f[2] == 1)) continue;
      s += 0.5 * (at(fi + di, fj + dj, fk + dk) + at(fi - di, fj - dj, fk - dk));
    }
    bc_out[t] = s;
  }
}

// x_f += P x_c, nothing on constrained fine dofs.  Slab form: x_f / bc_f are this rank's slab (local plane 0 = global fine plane fz0)
// and only its local scalar dofs [t0, t1) — the owned rows — are updated; x_c is the whole coarse vector.
__global__ void k_mg_prolong_add(Dims g, int bs, const double* __restrict__ xc, const uint8_t* __restrict__ bc_f, double* __restrict__ xf,
                                 int64_t t0, int64_t t1, int fz0) {
  for (int64_t t = t0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < t1; t += (int64_t)gridDim.x * blockDim.x) {
    if (bc_f && bc_f[t]) continue;
    const int64_t node = t / bs;
    const int comp = (int)(t % bs);
    const int i = (int)(node % g.f[0]), j = (int)((node / g.f[0]) % g.f[1]), k = (int)(node / ((int64_t)g.f[0] * g.f[1])) + fz0;
    const int di = i & 1, dj = j & 1, dk = k & 1;
    const int I = i >> 1, J = j >> 1, K = k >> 1;
    const int64_t c0 = I + (int64_t)g.c[0] * (J + (int64_t)g.c[1] * K);
    const int64_t c1 = (I + di) + (int64_t)g.c[0] * ((J + dj) + (int64_t)g.c[1] * (K + dk));
    const double v = (di | dj | dk) ? 0.5 * (xc[c0 * bs + comp] + xc[c1 * bs + comp]) : xc[c0 * bs + comp];
    xf[t] += v;
  }
}

__global__ void k_mg_axpy(int64_t n, double a, const double* __restrict__ x, double* __restrict__ y) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] += a * x[i];
}
// p = z + beta p
__global__ void k_mg_xpay(int64_t n, double beta, const double* __restrict__ z, double* __restrict__ p) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = z[i] + beta * p[i];
}
__global__ void k_mg_scale_dinv(int64_t n, const double* __restrict__ dinv, const double* __restrict__ x, double* __restrict__ y) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = dinv[i] * x[i];
}

// deterministic two-stage dot product: per-CTA partials, then one CTA sums them in index order
__global__ void k_mg_dot(int64_t n, const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ w, double* __restrict__ partials) {
  __shared__ double red[32];
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    s += (w ? w[i] * w[i] : 1.0) * x[i] * y[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}
// x += a p ; r -= a q ; per-CTA partial of |D^-1 r|^2 (finished by k_mg_dot_final)
__global__ void k_mg_update(int64_t n, double a, const double* __restrict__ p, const double* __restrict__ q, const double* __restrict__ dinv,
                            double* __restrict__ x, double* __restrict__ r, double* __restrict__ partials) {
  __shared__ double red[32];
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    x[i] += a * p[i];
    const double ri = r[i] - a * q[i];
    r[i] = ri;
    const double zi = dinv[i] * ri;
    s += zi * zi;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}
__global__ void k_mg_dot_final(int nparts, const double* __restrict__ partials, double* __restrict__ out) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) s += partials[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) *out = s;
}

}  // namespace

struct fsb_mg {
  fsb_ctx* ctx = nullptr;
  int tdim = 3, bs = 1, nu = 2, coarse_sweeps = 24;
  std::vector<Level> lv;
  double *p = nullptr, *q = nullptr, *z = nullptr, *r = nullptr;     // fine-level PCG vectors
  // Slab-distributed fine level (one process per GPU): level 0 is this rank's z-slab of the fine problem (owned planes
  // [oz0, oz1) of the global fine grid, local plane 0 = global plane fz0, ghost planes refreshed by halo exchanges), every
  // coarser level is REPLICATED — each rank holds and cycles the whole coarse hierarchy, which is 1/7 of a fine level's work,
  // so the fine smoothing/residuals (7/8 of a cycle) scale with the ranks and no coarse level ever has fewer planes than ranks.
  // The restricted residual is assembled by one all-reduce of the level-1 right-hand side (every rank contributes the coarse
  // planes under its owned fine planes, zeros elsewhere: the sum is exact).
  bool dist = false;
  int fz0 = 0, oz0 = 0, oz1 = 0;
};

static unsigned mg_grid(fsb_ctx* ctx, int64_t n) { return fsb_grid(n, 256, (int64_t)ctx->sm_count * 8); }

static int mg_dot(fsb_mg* mg, int64_t n, const double* x, const double* y, const double* w, double* host_out, bool slab = false) {
  fsb_ctx* ctx = mg->ctx;
  const unsigned g = std::min<unsigned>(mg_grid(ctx, n), kMaxPartials);
  k_mg_dot<<<g, 256, 0, ctx->stream>>>(n, x, y, w, ctx->d_partials);
  FSB_LAUNCH_CHECK(ctx);
  k_mg_dot_final<<<1, 256, 0, ctx->stream>>>((int)g, ctx->d_partials, ctx->d_scalars + 48);
  FSB_LAUNCH_CHECK(ctx);
  if (slab) {           // the caller passed the owned range of its slab: the global sum is one all-reduce away
    int rc = fsb_dist_allreduce_sum_dev(ctx, ctx->d_scalars + 48, 1);
    if (rc) return rc;
  }
  FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(host_out, ctx->d_scalars + 48, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FSB_OK;
}

// y = A x on the level's rows (the owned rows of a slab-distributed fine level, after refreshing the ghost planes of x)
static int mg_spmv(fsb_mg* mg, Level& L, double* x, double* y) {
  if (mg->dist && &L == &mg->lv[0]) {
    int rc = fsb_dist_halo_raw(mg->ctx, x, L.n);
    if (rc) return rc;
  }
  return fsb_launch_spmv(L.S ? L.S : L.A, x, y, nullptr, 0, nullptr, nullptr);
}

// The level matrices were (re)assembled since the last cycle: refresh the operands the smoothers and residuals multiply by.
// With drop_zeros (default 'auto') a level whose stored blocks are >= 20 % exact zeros (every level of a right-angled box
// hierarchy) gets the compacted copy: a V(2,2) cycle runs five SpMVs per level, the count + compact passes cost two.
static int mg_prepare_operands(fsb_mg* mg) {
  for (Level& L : mg->lv) {
    L.S = L.A;
    if (mg->ctx->drop_zeros) {
      int rc = fsb_mat_squeeze(L.A, &L.S);
      if (rc) return rc;
    }
  }
  return FSB_OK;
}

static int mg_smooth(fsb_mg* mg, Level& L, int sweeps, bool zero_start) {
  fsb_ctx* ctx = mg->ctx;
  const unsigned g = mg_grid(ctx, L.len);
  const int64_t o = L.off;
  for (int s = 0; s < sweeps; ++s) {
    if (s == 0 && zero_start) {
      k_mg_jacobi0<<<g, 256, 0, ctx->stream>>>(L.len, L.omega, L.dinv + o, L.b + o, L.x + o);
    } else {
      int rc = mg_spmv(mg, L, L.x, L.y);
      if (rc) return rc;
      k_mg_jacobi<<<g, 256, 0, ctx->stream>>>(L.len, L.omega, L.dinv + o, L.b + o, L.y + o, L.x + o);
    }
    FSB_LAUNCH_CHECK(ctx);
  }
  return FSB_OK;
}

// Chebyshev smoother of degree `deg` for D^-1 A on [lmax / kChebRatio, lmax], lmax = 4 / (3 omega) (the level's estimate);
// L.r is the direction vector (free while smoothing)
static constexpr double kChebRatio = 10.0;
static int mg_chebyshev(fsb_mg* mg, Level& L, int deg, bool zero_start) {
  fsb_ctx* ctx = mg->ctx;
  const unsigned g = mg_grid(ctx, L.len);
  const int64_t o = L.off;
  const double lmax = 4.0 / (3.0 * L.omega), lmin = lmax / kChebRatio;
  const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;
  double rho0 = 1.0 / sigma;
  for (int k = 0; k < deg; ++k) {
    double c1 = 0.0, c2 = 1.0 / theta;
    if (k > 0) {
      const double rho1 = 1.0 / (2.0 * sigma - rho0);
      c1 = rho1 * rho0; c2 = 2.0 * rho1 / delta;
      rho0 = rho1;
    }
    if (k == 0 && zero_start) {
      k_mg_cheb0<<<g, 256, 0, ctx->stream>>>(L.len, c2, L.dinv + o, L.b + o, L.r + o, L.x + o);
    } else {
      int rc = mg_spmv(mg, L, L.x, L.y);
      if (rc) return rc;
      k_mg_cheb<<<g, 256, 0, ctx->stream>>>(L.len, c1, c2, L.dinv + o, L.b + o, L.y + o, L.r + o, L.x + o);
    }
    FSB_LAUNCH_CHECK(ctx);
  }
  return FSB_OK;
}

// x_l = V(b_l), zero initial guess
static int mg_vcycle(fsb_mg* mg, int l) {
  fsb_ctx* ctx = mg->ctx;
  Level& L = mg->lv[l];
  int rc;
  if (l + 1 == (int)mg->lv.size()) return mg_smooth(mg, L, mg->coarse_sweeps, true);
  if ((rc = mg_chebyshev(mg, L, mg->nu, true))) return rc;
  if ((rc = mg_spmv(mg, L, L.x, L.y))) return rc;
  k_mg_residual<<<mg_grid(ctx, L.len), 256, 0, ctx->stream>>>(L.len, L.b + L.off, L.y + L.off, L.r + L.off);
  FSB_LAUNCH_CHECK(ctx);
  Level& C = mg->lv[l + 1];
  Dims g;
  for (int a = 0; a < 3; ++a) { g.f[a] = L.dims[a]; g.c[a] = C.dims[a]; }
  const bool slab = mg->dist && l == 0;
  if (slab) {
    // the coarse planes under this rank's owned fine planes need r on one ghost plane per side; the other planes of the
    // (replicated) coarse right-hand side come from the other ranks through one all-reduce of zeros + one contribution each
    if ((rc = fsb_dist_halo_raw(ctx, L.r, L.n))) return rc;
    FSB_CHECK_CUDA(ctx, cudaMemsetAsync(C.b, 0, sizeof(double) * C.n, ctx->stream));
    const int ck0 = (mg->oz0 + 1) / 2, ck1 = (mg->oz1 + 1) / 2;
    if (ck1 > ck0) {
      k_mg_restrict<<<mg_grid(ctx, (int64_t)(ck1 - ck0) * g.c[0] * g.c[1] * mg->bs), 256, 0, ctx->stream>>>(g, mg->bs, L.r, C.A->bc_flag, C.b, ck0, ck1, mg->fz0);
      FSB_LAUNCH_CHECK(ctx);
    }
    if (C.n > 0x7fffffffll) FSB_FAIL(ctx, FSB_ERR_ARG, "coarse level too large for one all-reduce");
    if ((rc = fsb_dist_allreduce_sum_dev(ctx, C.b, (int)C.n))) return rc;
  } else {
    k_mg_restrict<<<mg_grid(ctx, C.n), 256, 0, ctx->stream>>>(g, mg->bs, L.r, C.A->bc_flag, C.b, 0, g.c[2], 0);
    FSB_LAUNCH_CHECK(ctx);
  }
  if ((rc = mg_vcycle(mg, l + 1))) return rc;
  k_mg_prolong_add<<<mg_grid(ctx, L.len), 256, 0, ctx->stream>>>(g, mg->bs, C.x, L.A->bc_flag, L.x, L.off, L.off + L.len, slab ? mg->fz0 : 0);
  FSB_LAUNCH_CHECK(ctx);
  return mg_chebyshev(mg, L, mg->nu, false);
}

// lambda_max(D^-1 A): power iteration from a random-sign vector (approaches from below: 10 % safety), capped by the
// Gershgorin bound (rigorous, exact for the P1 Laplacian's interior rows); damping 4 / (3 lambda_max)
static int mg_estimate_omega(fsb_mg* mg, Level& L) {
  fsb_ctx* ctx = mg->ctx;
  const unsigned g = mg_grid(ctx, L.len);
  const int64_t o = L.off;
  const bool slab = mg->dist && &L == &mg->lv[0];
  unsigned long long* d_bits = reinterpret_cast<unsigned long long*>(ctx->d_scalars + 49);
  FSB_CHECK_CUDA(ctx, cudaMemsetAsync(d_bits, 0, sizeof(unsigned long long), ctx->stream));
  k_mg_gershgorin<<<g, 256, 0, ctx->stream>>>(o, L.len, mg->bs, L.A->row_ptr, L.A->vals, L.dinv, d_bits);
  FSB_LAUNCH_CHECK(ctx);
  double bound = 0.0;
  FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(&bound, d_bits, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (slab) {
    int rc = fsb_dist_allreduce_max(ctx, &bound);
    if (rc) return rc;
  }
  if (slab) FSB_CHECK_CUDA(ctx, cudaMemsetAsync(L.x, 0, sizeof(double) * L.n, ctx->stream));
  // signs by global dof index (local index + the slab's offset in the global numbering): every partition starts from the same vector
  k_mg_randsign<<<g, 256, 0, ctx->stream>>>(L.len, slab ? (int64_t)mg->fz0 * L.dims[0] * L.dims[1] * mg->bs + o : 0, L.x + o);
  FSB_LAUNCH_CHECK(ctx);
  double lam = 0.0;
  for (int it = 0; it < 30; ++it) {
    int rc = mg_spmv(mg, L, L.x, L.y);
    if (rc) return rc;
    k_mg_scale_dinv<<<g, 256, 0, ctx->stream>>>(L.len, L.dinv + o, L.y + o, L.r + o);          // r = D^-1 A x
    FSB_LAUNCH_CHECK(ctx);
    double xx, rr;
    if ((rc = mg_dot(mg, L.len, L.x + o, L.x + o, nullptr, &xx, slab)) || (rc = mg_dot(mg, L.len, L.r + o, L.r + o, nullptr, &rr, slab))) return rc;
    if (!(xx > 0.0) || !(rr > 0.0)) break;
    lam = std::sqrt(rr / xx);
    FSB_CHECK_CUDA(ctx, cudaMemsetAsync(L.x, 0, sizeof(double) * L.n, ctx->stream));
    k_mg_axpy<<<g, 256, 0, ctx->stream>>>(L.len, 1.0 / std::sqrt(rr), L.r + o, L.x + o);
    FSB_LAUNCH_CHECK(ctx);
  }
  double est = 1.1 * lam;
  if (bound > 0.0 && (est <= 0.0 || est > bound)) est = bound;
  if (!(est > 0.0)) est = 2.0;
  L.omega = 4.0 / (3.0 * est);
  return FSB_OK;
}

extern "C" void fsb_mg_destroy(fsb_mg* mg) {
  if (!mg) return;
  for (Level& L : mg->lv) {
    fsb_dfree(mg->ctx, L.dinv); fsb_dfree(mg->ctx, L.x); fsb_dfree(mg->ctx, L.b); fsb_dfree(mg->ctx, L.r); fsb_dfree(mg->ctx, L.y);
  }
  fsb_dfree(mg->ctx, mg->p); fsb_dfree(mg->ctx, mg->q); fsb_dfree(mg->ctx, mg->z); fsb_dfree(mg->ctx, mg->r);
  delete mg;
}

static int mg_create_impl(fsb_ctx* ctx, int32_t nlevels, fsb_mat** A, const int32_t* ncells, int32_t tdim, const double* omega,
                          bool dist, int layer0, int oz0, int oz1, fsb_mg** out) {
  if (!ctx || !A || !ncells || !out || nlevels < 1 || (tdim != 2 && tdim != 3)) return FSB_ERR_ARG;
  if (fsb_dist_active(ctx) && !dist) FSB_FAIL(ctx, FSB_ERR_STATE, "a distributed context needs fsb_mg_create_slab (slab-distributed fine level)");
  if (dist && (!fsb_dist_active(ctx) || nlevels < 2)) FSB_FAIL(ctx, FSB_ERR_STATE, "fsb_mg_create_slab needs an initialised distributed context and a coarse level");
  if (dist && tdim != 3) FSB_FAIL(ctx, FSB_ERR_ARG, "the slab-distributed multigrid is implemented for 3-D boxes (slabs along z)");
  fsb_mg* mg = new fsb_mg();
  mg->ctx = ctx; mg->tdim = tdim; mg->bs = A[0]->bs;
  mg->dist = dist; mg->fz0 = layer0; mg->oz0 = oz0; mg->oz1 = oz1;
  mg->lv.resize(nlevels);
  int rc = FSB_OK;
  for (int l = 0; l < nlevels && !rc; ++l) {
    Level& L = mg->lv[l];
    L.A = A[l];
    if (!L.A || L.A->bs != mg->bs) { ctx->err = "multigrid levels must share the block size"; rc = FSB_ERR_ARG; break; }
    int64_t nn = 1;
    for (int a = 0; a < 3; ++a) { L.dims[a] = a < tdim ? ncells[l * 3 + a] + 1 : 1; nn *= L.dims[a]; }
    const bool slab = dist && l == 0;
    if (slab) {
      // the fine matrix holds this rank's slab: whole planes of the global grid, the owned ones at rows [own0, own1)
      const int64_t plane = nn / L.dims[tdim - 1];
      if (L.A->nbrows % plane != 0 || L.A->own0 != (int64_t)(oz0 - layer0) * plane || L.A->own1 != (int64_t)(oz1 - layer0) * plane ||
          layer0 < 0 || oz0 < layer0 || oz1 <= oz0 || oz1 > L.dims[tdim - 1]) {
        ctx->err = "fine-level slab does not match its matrix (planes / owned rows)"; rc = FSB_ERR_ARG; break;
      }
      nn = L.A->nbrows;
    } else if (nn != L.A->nbrows) { ctx->err = "level matrix does not match its box dimensions"; rc = FSB_ERR_ARG; break; }
    if (l > 0)
      for (int a = 0; a < tdim; ++a)
        if (ncells[(l - 1) * 3 + a] != 2 * ncells[l * 3 + a]) { ctx->err = "levels must halve the cell counts"; rc = FSB_ERR_ARG; }
    if (rc) break;
    L.nnode = nn; L.n = nn * mg->bs;
    L.off = slab ? L.A->own0 * mg->bs : 0;
    L.len = slab ? (L.A->own1 - L.A->own0) * mg->bs : L.n;
    if ((rc = fsb_dmalloc(ctx, &L.dinv, (size_t)L.n)) || (rc = fsb_dmalloc(ctx, &L.x, (size_t)L.n)) || (rc = fsb_dmalloc(ctx, &L.b, (size_t)L.n)) ||
        (rc = fsb_dmalloc(ctx, &L.r, (size_t)L.n)) || (rc = fsb_dmalloc(ctx, &L.y, (size_t)L.n)))
      break;
    if (slab) {          // ghost rows of the work vectors are read by halo sends before anything writes them
      for (double* v : {L.x, L.b, L.r, L.y}) cudaMemsetAsync(v, 0, sizeof(double) * L.n, ctx->stream);
    }
    k_mg_dinv<<<mg_grid(ctx, L.n), 256, 0, ctx->stream>>>(L.n, mg->bs, L.A->row_ptr, L.A->col_idx, L.A->vals, L.dinv);
    ctx->launches++;
    if (omega && omega[l] > 0.0) L.omega = omega[l];
    else if ((rc = mg_estimate_omega(mg, L))) break;
  }
  const int64_t n0 = mg->lv[0].n;
  if (!rc) rc = fsb_dmalloc(ctx, &mg->p, (size_t)n0);
  if (!rc) rc = fsb_dmalloc(ctx, &mg->q, (size_t)n0);
  if (!rc) rc = fsb_dmalloc(ctx, &mg->z, (size_t)n0);
  if (!rc) rc = fsb_dmalloc(ctx, &mg->r, (size_t)n0);
  if (!rc && dist) for (double* v : {mg->p, mg->q, mg->z, mg->r}) cudaMemsetAsync(v, 0, sizeof(double) * n0, ctx->stream);
  if (rc) { fsb_mg_destroy(mg); return rc; }
  *out = mg;
  return FSB_OK;
}

extern "C" int fsb_mg_create(fsb_ctx* ctx, int32_t nlevels, fsb_mat** A, const int32_t* ncells, int32_t tdim, const double* omega,
                             fsb_mg** out) {
  return mg_create_impl(ctx, nlevels, A, ncells, tdim, omega, false, 0, 0, 0, out);
}

extern "C" int fsb_mg_create_slab(fsb_ctx* ctx, int32_t nlevels, fsb_mat** A, const int32_t* ncells, int32_t tdim, const double* omega,
                                  int32_t layer0, int32_t owned_z0, int32_t owned_z1, fsb_mg** out) {
  return mg_create_impl(ctx, nlevels, A, ncells, tdim, omega, true, layer0, owned_z0, owned_z1, out);
}

extern "C" int fsb_mg_omega(fsb_mg* mg, int32_t level, double* omega) {
  if (!mg || !omega || level < 0 || level >= (int)mg->lv.size()) return FSB_ERR_ARG;
  *omega = mg->lv[level].omega;
  return FSB_OK;
}

// z = V(r) on the fine level: the cycle reads r as the level's right-hand side and builds the iterate directly in z
static int mg_apply(fsb_mg* mg, const double* r, double* z) {
  Level& L = mg->lv[0];
  double *b0 = L.b, *x0 = L.x;
  L.b = const_cast<double*>(r);      // never written by the cycle
  L.x = z;
  const int rc = mg_vcycle(mg, 0);
  L.b = b0; L.x = x0;
  return rc;
}

extern "C" int fsb_mg_apply(fsb_mg* mg, fsb_vec* r, fsb_vec* z, int32_t nu) {
  if (!mg || !r || !z) return FSB_ERR_ARG;
  if (r->n != mg->lv[0].n || z->n != r->n || r == z) FSB_FAIL(mg->ctx, FSB_ERR_ARG, "vector sizes do not match the fine level");
  mg->nu = nu > 0 ? nu : 2;
  int rc = mg_prepare_operands(mg);
  if (rc) return rc;
  return mg_apply(mg, r->d, z->d);
}

extern "C" int fsb_solve_cg_mg(fsb_mg* mg, fsb_vec* b, fsb_vec* x, double rtol, double atol, int32_t maxit, int32_t nu,
                               fsb_solve_info* info) {
  if (!mg || !b || !x || !info) return FSB_ERR_ARG;
  fsb_ctx* ctx = mg->ctx;
  Level& L = mg->lv[0];
  const int64_t n = L.n;
  if (b->n != n || x->n != n) FSB_FAIL(ctx, FSB_ERR_ARG, "vector sizes do not match the fine level");
  memset(info, 0, sizeof(*info));
  mg->nu = nu > 0 ? nu : 2;

  cudaEvent_t e0, e1;
  FSB_CHECK_CUDA(ctx, cudaEventCreate(&e0));
  FSB_CHECK_CUDA(ctx, cudaEventCreate(&e1));
  struct EvGuard { cudaEvent_t a, b; ~EvGuard() { cudaEventDestroy(a); cudaEventDestroy(b); } } guard{e0, e1};
  FSB_CHECK_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
  // element-wise work and dot products run over the level's own rows: all of them, or the owned rows of this rank's slab
  const int64_t o = L.off, len = L.len;
  const bool slab = mg->dist;
  const unsigned g = mg_grid(ctx, len);
  double *p = mg->p, *q = mg->q, *z = mg->z, *r = mg->r;
  int rc;
  if ((rc = mg_prepare_operands(mg))) return rc;        // inside the timed solve
  info->operand_nnzb = L.S->nnzb;
  // convergence on the Jacobi-scaled residual ||D^-1 r|| <= max(rtol ||D^-1 b||, atol): the same norm as fsb_solve_cg
  double bb, rr, rz, pq;
  if ((rc = mg_dot(mg, len, b->d + o, b->d + o, L.dinv + o, &bb, slab))) return rc;
  if ((rc = mg_spmv(mg, L, x->d, q))) return rc;
  k_mg_residual<<<g, 256, 0, ctx->stream>>>(len, b->d + o, q + o, r + o);
  FSB_LAUNCH_CHECK(ctx);
  if ((rc = mg_dot(mg, len, r + o, r + o, L.dinv + o, &rr, slab))) return rc;
  const double tol2 = std::max(rtol * rtol * bb, atol * atol);
  int it = 0, outcome = 0;
  if (rr <= tol2) outcome = 1;
  else if (!(rr == rr)) outcome = -1;
  if (!outcome) {
    if ((rc = mg_apply(mg, r, z))) return rc;
    FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(p + o, z + o, sizeof(double) * len, cudaMemcpyDeviceToDevice, ctx->stream));
    if ((rc = mg_dot(mg, len, r + o, z + o, nullptr, &rz, slab))) return rc;
    while (it < maxit) {
      // q = A p with p.q fused into the SpMV (no separate pass over p and q); distributed: ghost planes of p first, p.q all-reduced
      if (slab && (rc = fsb_dist_halo_raw(ctx, p, n))) return rc;
      if ((rc = fsb_launch_spmv(L.S, p, q, p, 0, ctx->d_scalars + 50, nullptr))) return rc;
      if (slab && (rc = fsb_dist_allreduce_sum_dev(ctx, ctx->d_scalars + 50, 1))) return rc;
      FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(&pq, ctx->d_scalars + 50, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
      FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      if (!(pq == pq) || pq == 0.0 || !(rz == rz)) { outcome = -1; break; }
      const double alpha = rz / pq;
      {
        const unsigned gp = std::min<unsigned>(g, kMaxPartials);
        k_mg_update<<<gp, 256, 0, ctx->stream>>>(len, alpha, p + o, q + o, L.dinv + o, x->d + o, r + o, ctx->d_partials);
        FSB_LAUNCH_CHECK(ctx);
        k_mg_dot_final<<<1, 256, 0, ctx->stream>>>((int)gp, ctx->d_partials, ctx->d_scalars + 48);
        FSB_LAUNCH_CHECK(ctx);
        if (slab && (rc = fsb_dist_allreduce_sum_dev(ctx, ctx->d_scalars + 48, 1))) return rc;
        FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(&rr, ctx->d_scalars + 48, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      }
      ++it;
      if (rr <= tol2) { outcome = 1; break; }
      if (!(rr == rr)) { outcome = -1; break; }
      if ((rc = mg_apply(mg, r, z))) return rc;
      double rz_new;
      if ((rc = mg_dot(mg, len, r + o, z + o, nullptr, &rz_new, slab))) return rc;
      const double beta = rz_new / rz;
      rz = rz_new;
      k_mg_xpay<<<g, 256, 0, ctx->stream>>>(len, beta, z + o, p + o);
      FSB_LAUNCH_CHECK(ctx);
    }
  }
  FSB_CHECK_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
  FSB_CHECK_CUDA(ctx, cudaEventSynchronize(e1));
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  info->iterations = it;
  info->converged = outcome;
  info->rnorm = std::sqrt(std::max(rr, 0.0));
  info->bnorm = std::sqrt(std::max(bb, 0.0));
  info->solve_ms = ms;
  if (outcome < 0) FSB_FAIL(ctx, FSB_ERR_BREAKDOWN, "multigrid-preconditioned CG breakdown (non-finite or zero recurrence scalar)");
  return FSB_OK;
}
