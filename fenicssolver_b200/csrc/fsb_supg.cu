// SUPG stabilisation of the advection term (ScalarTransportSolver.py:252-274, "SPUG" method 2): the test function
// becomes  Tq = q + tau v.grad(q),  tau = 0.5 h / (4/(Pe h) + 2 |v|),  h = 2 * circumradius of the cell.
// For P1 and a constant velocity, v.grad(phi_a) = v.G_a is constant per cell, so every integral of the form picks up
// s_a = tau (v.G_a) times the integral of the rest (grad Tq = grad q: the diffusion term does not change):
//   cells   A_ab += s_a [ mass |T|/(D+1) + adv (v.G_b) |T| ]          b_a += S s_a |T|
//   facets  b_a  += g s_a |F|        A_ab += h s_a |F|/D  (b on the facet)   for every node a of the adjacent cell
// The entry points below add ONLY these extra terms; the Galerkin parts come from the regular assembly calls.
#include "fsb_internal.cuh"
#include "fsb_p1.cuh"

struct SupgParams {
  double vel[3];
  double vnorm;
  double pe;
};

template <int D>
__device__ __forceinline__ double circumradius(const double* __restrict__ xyz, const int (&v)[D + 1], double vol) {
  double X[D + 1][D];
#pragma unroll
  for (int a = 0; a <= D; ++a)
#pragma unroll
    for (int i = 0; i < D; ++i) X[a][i] = __ldg(xyz + (int64_t)v[a] * D + i);
  auto dist = [&](int p, int q) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i) s += (X[p][i] - X[q][i]) * (X[p][i] - X[q][i]);
    return sqrt(s);
  };
  if constexpr (D == 2) {
    return dist(0, 1) * dist(1, 2) * dist(0, 2) / (4.0 * vol);
  } else {
    // products of opposite edge lengths
    const double pa = dist(0, 1) * dist(2, 3), pb = dist(0, 2) * dist(1, 3), pc = dist(0, 3) * dist(1, 2);
    const double q = (pa + pb + pc) * (pa + pb - pc) * (pa - pb + pc) * (-pa + pb + pc);
    return sqrt(fmax(q, 0.0)) / (24.0 * vol);
  }
}

// s_a = tau (v . G_a) for the nodes of the cell
template <int D>
__device__ __forceinline__ void supg_weights(const double* __restrict__ xyz, const int (&v)[D + 1], const Geo<D>& g,
                                             const SupgParams& p, double (&s)[D + 1]) {
  const double h = 2.0 * circumradius<D>(xyz, v, g.vol);
  const double tau = 0.5 * h / (4.0 / (p.pe * h) + 2.0 * p.vnorm);
#pragma unroll
  for (int a = 0; a <= D; ++a) {
    double d = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i) d += p.vel[i] * g.G[a][i];
    s[a] = tau * d;
  }
}

template <int D, bool ACTION>
__global__ void __launch_bounds__(128)
k_scalar_supg(int64_t ncells, const int32_t* __restrict__ cells, const double* __restrict__ xyz, SupgParams p, double mass,
              double adv, const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx, double* __restrict__ vals,
              const uint8_t* __restrict__ posmap, const double* __restrict__ x, double* __restrict__ y) {
  constexpr int NL = D + 1;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncells; c += (int64_t)gridDim.x * blockDim.x) {
    int v[NL];
    load_cell<D>(cells, c, v);
    Geo<D> g;
    p1_geometry<D>(xyz, v, g);
    double s[NL], col[NL];
    supg_weights<D>(xyz, v, g, p, s);
#pragma unroll
    for (int b = 0; b < NL; ++b) {
      double vg = 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i) vg += p.vel[i] * g.G[b][i];
      col[b] = g.vol * (mass / (double)NL + adv * vg);          // int (mass phi_b + adv v.grad phi_b)
    }
    if (ACTION) {
      double t = 0.0;
#pragma unroll
      for (int b = 0; b < NL; ++b) t += col[b] * __ldg(x + v[b]);
#pragma unroll
      for (int a = 0; a < NL; ++a) atomicAdd(y + v[a], s[a] * t);
    } else {
      int64_t base[NL];
      int pos[NL][NL];
      entry_positions<D>(posmap, c, v, row_ptr, col_idx, base, pos);
#pragma unroll
      for (int a = 0; a < NL; ++a)
#pragma unroll
        for (int b = 0; b < NL; ++b) add_nz(vals + base[a] + pos[a][b], s[a] * col[b]);
    }
  }
}

template <int D>
__global__ void k_source_supg(int64_t ncells, const int32_t* __restrict__ cells, const double* __restrict__ xyz, SupgParams p,
                              double S, const int32_t* __restrict__ tags, int tag, double* __restrict__ b) {
  constexpr int NL = D + 1;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncells; c += (int64_t)gridDim.x * blockDim.x) {
    if (tags && tags[c] != tag) continue;
    int v[NL];
    load_cell<D>(cells, c, v);
    Geo<D> g;
    p1_geometry<D>(xyz, v, g);
    double s[NL];
    supg_weights<D>(xyz, v, g, p, s);
#pragma unroll
    for (int a = 0; a < NL; ++a) atomicAdd(b + v[a], S * g.vol * s[a]);
  }
}

// the cell adjacent to an exterior facet is (facet vertices, opposite vertex)
template <int D>
__global__ void k_facet_supg(int64_t nf, const int32_t* __restrict__ fverts, const int32_t* __restrict__ opp,
                             const double* __restrict__ xyz, SupgParams p, double gload, double h,
                             const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx, double* __restrict__ vals,
                             double* __restrict__ b) {
  constexpr int NL = D + 1;
  for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nf; f += (int64_t)gridDim.x * blockDim.x) {
    const int32_t* fv = fverts + f * D;
    int v[NL];
    for (int a = 0; a < D; ++a) v[a] = fv[a];
    v[D] = opp[f];
    // the affine geometry does not depend on the vertex order, the weights s_a follow the order used here
    Geo<D> g;
    p1_geometry<D>(xyz, v, g);
    double s[NL];
    supg_weights<D>(xyz, v, g, p, s);
    double n[3], x0[3];
    const double meas = facet_geom<D>(xyz, fv, n, x0);
    for (int a = 0; a < NL; ++a) {
      if (b && gload != 0.0) atomicAdd(b + v[a], gload * s[a] * meas);
      if (vals && h != 0.0) {
        const int64_t base = row_ptr[v[a]];
        const int len = (int)(row_ptr[v[a] + 1] - base);
        for (int q = 0; q < D; ++q) atomicAdd(vals + base + row_find(col_idx + base, 0, len, fv[q]), h * s[a] * meas / (double)D);
      }
    }
  }
}

static int fill_supg(fsb_mesh* mesh, const double* vel, double pe, SupgParams& p) {
  fsb_ctx* ctx = mesh->ctx;
  if (mesh->degree != 1) FSB_FAIL(ctx, FSB_ERR_ARG, "SUPG is implemented for degree-1 spaces");
  if (!vel || !(pe > 0.0)) FSB_FAIL(ctx, FSB_ERR_ARG, "SUPG needs a velocity and a positive Peclet number");
  memset(&p, 0, sizeof(p));
  double n2 = 0.0;
  for (int i = 0; i < mesh->tdim; ++i) { p.vel[i] = vel[i]; n2 += vel[i] * vel[i]; }
  p.vnorm = sqrt(n2);
  p.pe = pe;
  return FSB_OK;
}

extern "C" int fsb_assemble_scalar_supg(fsb_mesh* mesh, fsb_mat* A, fsb_vec* x, fsb_vec* y, double mass, double adv,
                                        const double* vel, double pe) {
  if (!mesh || (!A && (!x || !y))) return FSB_ERR_ARG;
  fsb_ctx* ctx = mesh->ctx;
  SupgParams p;
  int rc = fill_supg(mesh, vel, pe, p);
  if (rc) return rc;
  if (A && (A->bs != 1 || A->nbrows != mesh->nnodes)) FSB_FAIL(ctx, FSB_ERR_ARG, "matrix does not belong to a scalar P1 space on this mesh");
  if (!A && (x->n != mesh->nnodes || y->n != mesh->nnodes || x == y)) FSB_FAIL(ctx, FSB_ERR_ARG, "vector sizes do not match the mesh");
  const uint8_t* pm = (A && ctx->asm_mode >= 1 && A->mesh == mesh) ? A->posmap : nullptr;
  const unsigned grid = fsb_grid(mesh->ncells, 128, (int64_t)ctx->sm_count * 64);
  if (mesh->tdim == 3) {
    if (A) k_scalar_supg<3, false><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cells, mesh->xyz, p, mass, adv, A->row_ptr, A->col_idx, A->vals, pm, nullptr, nullptr);
    else k_scalar_supg<3, true><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cells, mesh->xyz, p, mass, adv, nullptr, nullptr, nullptr, nullptr, x->d, y->d);
  } else {
    if (A) k_scalar_supg<2, false><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cells, mesh->xyz, p, mass, adv, A->row_ptr, A->col_idx, A->vals, pm, nullptr, nullptr);
    else k_scalar_supg<2, true><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cells, mesh->xyz, p, mass, adv, nullptr, nullptr, nullptr, nullptr, x->d, y->d);
  }
  FSB_LAUNCH_CHECK(ctx);
  return FSB_OK;
}

extern "C" int fsb_assemble_source_supg(fsb_mesh* mesh, fsb_vec* b, double S, const double* vel, double pe,
                                        const int32_t* cell_tags, int32_t tag) {
  if (!mesh || !b) return FSB_ERR_ARG;
  fsb_ctx* ctx = mesh->ctx;
  SupgParams p;
  int rc = fill_supg(mesh, vel, pe, p);
  if (rc) return rc;
  if (b->n != mesh->nnodes) FSB_FAIL(ctx, FSB_ERR_ARG, "rhs size does not match the mesh");
  int32_t* d_tags = nullptr;
  if (cell_tags) {
    rc = fsb_dmalloc(ctx, &d_tags, (size_t)mesh->ncells);
    if (rc) return rc;
    FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(d_tags, cell_tags, sizeof(int32_t) * mesh->ncells, cudaMemcpyHostToDevice, ctx->stream));
  }
  const unsigned grid = fsb_grid(mesh->ncells, 128, (int64_t)ctx->sm_count * 64);
  if (mesh->tdim == 3) k_source_supg<3><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cells, mesh->xyz, p, S, d_tags, tag, b->d);
  else k_source_supg<2><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cells, mesh->xyz, p, S, d_tags, tag, b->d);
  FSB_LAUNCH_CHECK(ctx);
  if (d_tags) {
    FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    fsb_dfree(ctx, d_tags);
  }
  return FSB_OK;
}

extern "C" int fsb_assemble_facet_supg(fsb_mesh* mesh, fsb_mat* A, fsb_vec* b, int64_t nf, const int32_t* fverts,
                                       const int32_t* opp, double g, double h, const double* vel, double pe) {
  if (!mesh || (!A && !b) || (nf > 0 && (!fverts || !opp))) return FSB_ERR_ARG;
  fsb_ctx* ctx = mesh->ctx;
  SupgParams p;
  int rc = fill_supg(mesh, vel, pe, p);
  if (rc) return rc;
  if ((b && b->n != mesh->nnodes) || (A && (A->bs != 1 || A->nbrows != mesh->nnodes)))
    FSB_FAIL(ctx, FSB_ERR_ARG, "SUPG facet terms need the scalar matrix / rhs of this mesh");
  if (nf == 0) return FSB_OK;
  const int D = mesh->tdim;
  int32_t *d_fv = nullptr, *d_opp = nullptr;
  rc = fsb_dmalloc(ctx, &d_fv, (size_t)nf * D);
  if (!rc) rc = fsb_dmalloc(ctx, &d_opp, (size_t)nf);
  if (rc) { fsb_dfree(ctx, d_fv); return rc; }
  FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(d_fv, fverts, sizeof(int32_t) * nf * D, cudaMemcpyHostToDevice, ctx->stream));
  FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(d_opp, opp, sizeof(int32_t) * nf, cudaMemcpyHostToDevice, ctx->stream));
  const unsigned grid = fsb_grid(nf, 128, (int64_t)ctx->sm_count * 16);
  if (D == 3)
    k_facet_supg<3><<<grid, 128, 0, ctx->stream>>>(nf, d_fv, d_opp, mesh->xyz, p, g, h, A ? A->row_ptr : nullptr, A ? A->col_idx : nullptr,
                                                   A ? A->vals : nullptr, b ? b->d : nullptr);
  else
    k_facet_supg<2><<<grid, 128, 0, ctx->stream>>>(nf, d_fv, d_opp, mesh->xyz, p, g, h, A ? A->row_ptr : nullptr, A ? A->col_idx : nullptr,
                                                   A ? A->vals : nullptr, b ? b->d : nullptr);
  FSB_LAUNCH_CHECK(ctx);
  FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  fsb_dfree(ctx, d_fv);
  fsb_dfree(ctx, d_opp);
  return FSB_OK;
}
