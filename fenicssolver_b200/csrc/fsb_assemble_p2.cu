// Degree-2 Lagrange (P2) element kernels on affine simplices: 6-node triangles, 10-node tetrahedra.
// (examples/test_linear_elasticity.py:105-106 runs on VectorFunctionSpace(mesh, "Lagrange", 2).)
//
// On an affine cell every P2 form is a contraction of the cell geometry (volume, gradients G_c of the
// barycentric coordinates) with constant reference tensors
//     R[i][j][c][e] = 1/|T| int dphi_i/dl_c dphi_j/dl_e      M[i][j] = 1/|T| int phi_i phi_j
//     S[i][j][e]    = 1/|T| int phi_i dphi_j/dl_e            F[i]    = 1/|T| int phi_i
// (and the same M, F one dimension lower for the facet terms).  The tables are integrated on the host
// with a collapsed Gauss-Legendre rule that is exact for these polynomials and read by the kernels
// through uniform (warp-broadcast) loads.  One thread per cell; each local row is formed in registers
// and scattered with REDG.E.ADD.F64 at positions from the uint8 position map (or an in-row search).
// Local node order: vertices (sorted), then edges in UFC order  tri (1,2)(0,2)(0,1), tet (2,3)(1,3)(1,2)(0,3)(0,2)(0,1).
#include "fsb_internal.cuh"
#include <cmath>

namespace {

// ------------------------------------------------------------------------------------ reference tables (host)
struct P2Layout {
  int D, NL, NN, NF;                 // dimension, vertices per cell, nodes per cell, nodes per facet
  size_t oR, oM, oS, oF, oMf, oFf, total;
  explicit P2Layout(int d) : D(d), NL(d + 1), NN((d + 1) * (d + 2) / 2), NF(d * (d + 1) / 2) {
    oR = 0; oM = oR + (size_t)NN * NN * NL * NL; oS = oM + (size_t)NN * NN; oF = oS + (size_t)NN * NN * NL;
    oMf = oF + NN; oFf = oMf + (size_t)NF * NF; total = oFf + NF;
  }
};

const int kEdges1[1][2] = {{0, 1}};
const int kEdges2[3][2] = {{1, 2}, {0, 2}, {0, 1}};
const int kEdges3[6][2] = {{2, 3}, {1, 3}, {1, 2}, {0, 3}, {0, 2}, {0, 1}};
const int (*edges_of(int d))[2] { return d == 1 ? kEdges1 : (d == 2 ? kEdges2 : kEdges3); }

// Gauss-Legendre nodes/weights on [0,1]
void gauss01(int n, std::vector<double>& x, std::vector<double>& w) {
  x.resize(n); w.resize(n);
  for (int i = 0; i < n; ++i) {
    double z = std::cos(M_PI * (i + 0.75) / (n + 0.5)), pp = 0;
    for (int it = 0; it < 100; ++it) {
      double p1 = 1.0, p2 = 0.0;
      for (int j = 1; j <= n; ++j) { double p3 = p2; p2 = p1; p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j; }
      pp = n * (z * p1 - p2) / (z * z - 1.0);
      double z1 = z; z = z1 - p1 / pp;
      if (std::fabs(z - z1) < 1e-15) break;
    }
    x[i] = 0.5 * (1.0 - z);
    w[i] = 1.0 / ((1.0 - z * z) * pp * pp);     // = 0.5 * 2/((1-z^2) pp^2)
  }
}

// P2 basis and d/dl_c at barycentric point l (d+1 entries)
void p2_eval(int d, const double* l, double* phi, double* dphi /*[NN][NL]*/) {
  const int NL = d + 1, NN = (d + 1) * (d + 2) / 2;
  for (int i = 0; i < NN * NL; ++i) dphi[i] = 0.0;
  for (int a = 0; a < NL; ++a) { phi[a] = l[a] * (2.0 * l[a] - 1.0); dphi[a * NL + a] = 4.0 * l[a] - 1.0; }
  const int (*E)[2] = edges_of(d);
  for (int k = 0; k < NN - NL; ++k) {
    const int a = E[k][0], b = E[k][1];
    phi[NL + k] = 4.0 * l[a] * l[b];
    dphi[(NL + k) * NL + a] = 4.0 * l[b];
    dphi[(NL + k) * NL + b] = 4.0 * l[a];
  }
}

// 1/|T| int over the reference d-simplex of f(l), by a collapsed (Duffy) Gauss rule with n points per axis
template <typename Fn>
void integrate_simplex(int d, int n, Fn&& f) {
  std::vector<double> gx, gw;
  gauss01(n, gx, gw);
  double fact = 1.0;
  for (int k = 2; k <= d; ++k) fact *= k;          // 1/|T_ref| = d!
  if (d == 1) {
    for (int i = 0; i < n; ++i) { double l[2] = {1.0 - gx[i], gx[i]}; f(l, gw[i]); }
  } else if (d == 2) {
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        const double u = gx[i], v = gx[j], x = u, y = v * (1.0 - u);
        double l[3] = {1.0 - x - y, x, y};
        f(l, fact * gw[i] * gw[j] * (1.0 - u));
      }
  } else {
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j)
        for (int k = 0; k < n; ++k) {
          const double u = gx[i], v = gx[j], t = gx[k], x = u, y = v * (1.0 - u), z = t * (1.0 - u) * (1.0 - v);
          double l[4] = {1.0 - x - y - z, x, y, z};
          f(l, fact * gw[i] * gw[j] * gw[k] * (1.0 - u) * (1.0 - u) * (1.0 - v));
        }
  }
}

std::vector<double> build_tables(int D) {
  P2Layout L(D);
  std::vector<double> T(L.total, 0.0);
  {
    const int NL = L.NL, NN = L.NN;
    integrate_simplex(D, 6, [&](const double* l, double w) {
      double phi[10], dphi[40];
      p2_eval(D, l, phi, dphi);
      for (int i = 0; i < NN; ++i) {
        T[L.oF + i] += w * phi[i];
        for (int j = 0; j < NN; ++j) {
          T[L.oM + i * NN + j] += w * phi[i] * phi[j];
          for (int e = 0; e < NL; ++e) {
            T[L.oS + (i * NN + j) * NL + e] += w * phi[i] * dphi[j * NL + e];
            for (int c = 0; c < NL; ++c) T[L.oR + ((i * NN + j) * NL + c) * NL + e] += w * dphi[i * NL + c] * dphi[j * NL + e];
          }
        }
      }
    });
  }
  {
    const int NF = L.NF;
    integrate_simplex(D - 1, 6, [&](const double* l, double w) {
      double phi[10], dphi[40];
      p2_eval(D - 1, l, phi, dphi);
      for (int i = 0; i < NF; ++i) {
        T[L.oFf + i] += w * phi[i];
        for (int j = 0; j < NF; ++j) T[L.oMf + i * NF + j] += w * phi[i] * phi[j];
      }
    });
  }
  return T;
}

int ensure_tables(fsb_mesh* mesh) {
  if (mesh->p2_tables) return FSB_OK;
  fsb_ctx* ctx = mesh->ctx;
  std::vector<double> T = build_tables(mesh->tdim);
  int rc = fsb_dmalloc(ctx, &mesh->p2_tables, T.size());
  if (rc) return rc;
  FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(mesh->p2_tables, T.data(), sizeof(double) * T.size(), cudaMemcpyHostToDevice, ctx->stream));
  FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FSB_OK;
}

// ------------------------------------------------------------------------------------ device helpers
struct Form {
  double kscale, K[9], mass, adv, vel[3];
};
struct V3 { double v[3]; };

template <int D>
struct Cell {
  double vol;
  double G[D + 1][D];
};

template <int D>
__device__ __forceinline__ void cell_geometry(const double* __restrict__ xyz, const int* v, Cell<D>& g) {
  double X[D + 1][D];
#pragma unroll
  for (int a = 0; a <= D; ++a)
#pragma unroll
    for (int i = 0; i < D; ++i) X[a][i] = __ldg(xyz + (int64_t)v[a] * D + i);
  if constexpr (D == 3) {
    double a[3], b[3], c[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { a[i] = X[1][i] - X[0][i]; b[i] = X[2][i] - X[0][i]; c[i] = X[3][i] - X[0][i]; }
    const double bc[3] = {b[1] * c[2] - b[2] * c[1], b[2] * c[0] - b[0] * c[2], b[0] * c[1] - b[1] * c[0]};
    const double ca[3] = {c[1] * a[2] - c[2] * a[1], c[2] * a[0] - c[0] * a[2], c[0] * a[1] - c[1] * a[0]};
    const double ab[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    const double det = a[0] * bc[0] + a[1] * bc[1] + a[2] * bc[2], inv = 1.0 / det;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      g.G[1][i] = bc[i] * inv; g.G[2][i] = ca[i] * inv; g.G[3][i] = ab[i] * inv;
      g.G[0][i] = -(g.G[1][i] + g.G[2][i] + g.G[3][i]);
    }
    g.vol = fabs(det) * (1.0 / 6.0);
  } else {
    const double a0 = X[1][0] - X[0][0], a1 = X[1][1] - X[0][1], b0 = X[2][0] - X[0][0], b1 = X[2][1] - X[0][1];
    const double det = a0 * b1 - a1 * b0, inv = 1.0 / det;
    g.G[1][0] = b1 * inv; g.G[1][1] = -b0 * inv;
    g.G[2][0] = -a1 * inv; g.G[2][1] = a0 * inv;
    g.G[0][0] = -(g.G[1][0] + g.G[2][0]); g.G[0][1] = -(g.G[1][1] + g.G[2][1]);
    g.vol = fabs(det) * 0.5;
  }
}

// offset of column `col` in row `row` (position map when available, else in-row search)
__device__ __forceinline__ int entry_pos(const uint8_t* __restrict__ posmap, int64_t c, int nn, int i, int j, const int32_t* __restrict__ cols,
                                         int len, int32_t col) {
  if (posmap) return __ldg(posmap + (c * nn + i) * nn + j);
  return row_find(cols, 0, len, col);
}

// ------------------------------------------------------------------------------------ cell kernels
// A += kscale K(k) + mass M + adv C(vel)   or (ACTION)   y += (...) x
template <int D, bool ACTION>
__global__ void __launch_bounds__(128)
k_p2_scalar(int64_t ncells, const int32_t* __restrict__ cell_nodes, const double* __restrict__ xyz, Form f,
            const double* __restrict__ tab, const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx,
            double* __restrict__ vals, const uint8_t* __restrict__ posmap, const double* __restrict__ x, double* __restrict__ y) {
  constexpr int NL = D + 1, NN = (D + 1) * (D + 2) / 2;
  const double* __restrict__ tR = tab;
  const double* __restrict__ tM = tR + NN * NN * NL * NL;
  const double* __restrict__ tS = tM + NN * NN;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncells; c += (int64_t)gridDim.x * blockDim.x) {
    int nd[NN];
#pragma unroll
    for (int a = 0; a < NN; ++a) nd[a] = __ldg(cell_nodes + c * NN + a);
    Cell<D> g;
    cell_geometry<D>(xyz, nd, g);
    double gkg[NL][NL], vgl[NL];
#pragma unroll
    for (int a = 0; a < NL; ++a) {
      double KG[D];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) s += f.K[i * D + j] * g.G[a][j];
        KG[i] = s;
      }
      vgl[a] = 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i) vgl[a] += f.vel[i] * g.G[a][i];
      vgl[a] *= f.adv * g.vol;
#pragma unroll
      for (int e = 0; e < NL; ++e) {           // gkg[e][a] = G_e . K G_a
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) s += g.G[e][i] * KG[i];
        gkg[e][a] = f.kscale * g.vol * s;
      }
    }
    const double mw = f.mass * g.vol;
    double xl[NN];
    if (ACTION) {
#pragma unroll
      for (int j = 0; j < NN; ++j) xl[j] = __ldg(x + nd[j]);
    }
    for (int i = 0; i < NN; ++i) {
      const int64_t base = ACTION ? 0 : __ldg(row_ptr + nd[i]);
      const int len = ACTION ? 0 : (int)(__ldg(row_ptr + nd[i] + 1) - base);
      double yi = 0.0;
      for (int j = 0; j < NN; ++j) {
        const double* r = tR + (i * NN + j) * NL * NL;
        double s = mw * __ldg(tM + i * NN + j);
#pragma unroll
        for (int cc = 0; cc < NL; ++cc)
#pragma unroll
          for (int e = 0; e < NL; ++e) s += __ldg(r + cc * NL + e) * gkg[cc][e];
#pragma unroll
        for (int e = 0; e < NL; ++e) s += __ldg(tS + (i * NN + j) * NL + e) * vgl[e];
        if (ACTION) yi += s * xl[j];
        else add_nz(vals + base + entry_pos(posmap, c, NN, i, j, col_idx + base, len, nd[j]), s);
      }
      if (ACTION) atomicAdd(y + nd[i], yi);
    }
  }
}

// block (p,q)[a][b] = |T| ( mu (tr W) d_ab + mu W[b][a] + lambda W[a][b] ),  W[a][b] = sum_ce R[p][q][c][e] G_c[a] G_e[b]
template <int D>
__global__ void __launch_bounds__(128)
k_p2_elasticity(int64_t ncells, const int32_t* __restrict__ cell_nodes, const double* __restrict__ xyz, double mu, double lambda,
                const double* __restrict__ tab, const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx,
                double* __restrict__ vals, const uint8_t* __restrict__ posmap) {
  constexpr int NL = D + 1, NN = (D + 1) * (D + 2) / 2;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncells; c += (int64_t)gridDim.x * blockDim.x) {
    int nd[NN];
#pragma unroll
    for (int a = 0; a < NN; ++a) nd[a] = __ldg(cell_nodes + c * NN + a);
    Cell<D> g;
    cell_geometry<D>(xyz, nd, g);
    for (int p = 0; p < NN; ++p) {
      const int64_t base = __ldg(row_ptr + nd[p]);
      const int len = (int)(__ldg(row_ptr + nd[p] + 1) - base);
      for (int q = 0; q < NN; ++q) {
        const double* r = tab + (p * NN + q) * NL * NL;
        double H[NL][D];
#pragma unroll
        for (int cc = 0; cc < NL; ++cc)
#pragma unroll
          for (int b = 0; b < D; ++b) {
            double s = 0.0;
#pragma unroll
            for (int e = 0; e < NL; ++e) s += __ldg(r + cc * NL + e) * g.G[e][b];
            H[cc][b] = s;
          }
        double W[D][D], tr = 0.0;
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
          for (int b = 0; b < D; ++b) {
            double s = 0.0;
#pragma unroll
            for (int cc = 0; cc < NL; ++cc) s += g.G[cc][a] * H[cc][b];
            W[a][b] = s;
            if (a == b) tr += s;
          }
        double* blk = vals + (base + entry_pos(posmap, c, NN, p, q, col_idx + base, len, nd[q])) * (D * D);
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
          for (int b = 0; b < D; ++b)
            add_nz(blk + a * D + b, g.vol * (mu * ((a == b ? tr : 0.0) + W[b][a]) + lambda * W[a][b]));
      }
    }
  }
}

// b_i += scale |T| F[i] S   (constant S), optionally only cells with tags[c] == tag
template <int D>
__global__ void k_p2_source_const(int64_t ncells, const int32_t* __restrict__ cell_nodes, const double* __restrict__ xyz, int ncomp,
                                  V3 S, double scale, const double* __restrict__ tF, const int32_t* __restrict__ tags, int tag,
                                  double* __restrict__ b) {
  constexpr int NN = (D + 1) * (D + 2) / 2;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncells; c += (int64_t)gridDim.x * blockDim.x) {
    if (tags && tags[c] != tag) continue;
    int nd[NN];
    for (int a = 0; a < NN; ++a) nd[a] = cell_nodes[c * NN + a];
    Cell<D> g;
    cell_geometry<D>(xyz, nd, g);
    for (int i = 0; i < NN; ++i) {
      const double w = scale * g.vol * tF[i];
      for (int k = 0; k < ncomp; ++k) atomicAdd(b + (int64_t)nd[i] * ncomp + k, w * S.v[k]);
    }
  }
}

// b += scale M_e S_nodes  (S given at the P2 nodes)
template <int D>
__global__ void k_p2_source_nodal(int64_t ncells, const int32_t* __restrict__ cell_nodes, const double* __restrict__ xyz, int ncomp,
                                  const double* __restrict__ S, double scale, const double* __restrict__ tM, double* __restrict__ b) {
  constexpr int NN = (D + 1) * (D + 2) / 2;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncells; c += (int64_t)gridDim.x * blockDim.x) {
    int nd[NN];
    for (int a = 0; a < NN; ++a) nd[a] = cell_nodes[c * NN + a];
    Cell<D> g;
    cell_geometry<D>(xyz, nd, g);
    for (int k = 0; k < ncomp; ++k)
      for (int i = 0; i < NN; ++i) {
        double s = 0.0;
        for (int j = 0; j < NN; ++j) s += tM[i * NN + j] * S[(int64_t)nd[j] * ncomp + k];
        atomicAdd(b + (int64_t)nd[i] * ncomp + k, scale * g.vol * s);
      }
  }
}

// ------------------------------------------------------------------------------------ facet kernels
template <int D>
__device__ __forceinline__ double facet_measure(const double* __restrict__ xyz, const int32_t* fv, double (&n)[3], double (&x0)[3]) {
  if constexpr (D == 3) {
    double p[3][3];
    for (int a = 0; a < 3; ++a) for (int i = 0; i < 3; ++i) p[a][i] = xyz[(int64_t)fv[a] * 3 + i];
    const double a[3] = {p[1][0] - p[0][0], p[1][1] - p[0][1], p[1][2] - p[0][2]};
    const double b[3] = {p[2][0] - p[0][0], p[2][1] - p[0][1], p[2][2] - p[0][2]};
    n[0] = a[1] * b[2] - a[2] * b[1]; n[1] = a[2] * b[0] - a[0] * b[2]; n[2] = a[0] * b[1] - a[1] * b[0];
    for (int i = 0; i < 3; ++i) x0[i] = p[0][i];
    return 0.5 * sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
  } else {
    const double t0 = xyz[(int64_t)fv[1] * 2] - xyz[(int64_t)fv[0] * 2], t1 = xyz[(int64_t)fv[1] * 2 + 1] - xyz[(int64_t)fv[0] * 2 + 1];
    n[0] = t1; n[1] = -t0; n[2] = 0.0;
    x0[0] = xyz[(int64_t)fv[0] * 2]; x0[1] = xyz[(int64_t)fv[0] * 2 + 1]; x0[2] = 0.0;
    return sqrt(t0 * t0 + t1 * t1);
  }
}

template <int D>
__global__ void k_p2_facet_load(int64_t nf, const int32_t* __restrict__ fnodes, const int32_t* __restrict__ opp, const double* __restrict__ xyz,
                                int ncomp, int mode, V3 gval, double scale, const double* __restrict__ tFf, double* __restrict__ b) {
  constexpr int NF = D * (D + 1) / 2;
  for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nf; f += (int64_t)gridDim.x * blockDim.x) {
    const int32_t* fn = fnodes + f * NF;
    double n[3], x0[3];
    const double meas = facet_measure<D>(xyz, fn, n, x0);
    double gl[3] = {gval.v[0], gval.v[1], gval.v[2]};
    if (mode == 1) {
      const double nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
      double d = 0.0;
      for (int i = 0; i < D; ++i) d += n[i] * (xyz[(int64_t)opp[f] * D + i] - x0[i]);
      const double sgn = d > 0 ? -1.0 : 1.0;
      for (int i = 0; i < D; ++i) gl[i] = gval.v[0] * sgn * n[i] / nn;
    }
    for (int i = 0; i < NF; ++i)
      for (int k = 0; k < ncomp; ++k) atomicAdd(b + (int64_t)fn[i] * ncomp + k, scale * meas * tFf[i] * gl[k]);
  }
}

template <int D>
__global__ void k_p2_facet_mass(int64_t nf, const int32_t* __restrict__ fnodes, const double* __restrict__ xyz, double h,
                                const double* __restrict__ tMf, const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx,
                                double* __restrict__ vals) {
  constexpr int NF = D * (D + 1) / 2;
  for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nf; f += (int64_t)gridDim.x * blockDim.x) {
    const int32_t* fn = fnodes + f * NF;
    double n[3], x0[3];
    const double w = h * facet_measure<D>(xyz, fn, n, x0);
    for (int i = 0; i < NF; ++i) {
      const int64_t base = row_ptr[fn[i]];
      const int len = (int)(row_ptr[fn[i] + 1] - base);
      for (int j = 0; j < NF; ++j) atomicAdd(vals + base + row_find(col_idx + base, 0, len, fn[j]), w * tMf[i * NF + j]);
    }
  }
}


// P2 basis gradients d(phi_j)/d(lambda_e) at the barycentric point l
template <int D>
__device__ __forceinline__ void p2_dphi(const double* l, double (&dphi)[(D + 1) * (D + 2) / 2][D + 1]) {
  constexpr int NL = D + 1, NN = (D + 1) * (D + 2) / 2;
#pragma unroll
  for (int j = 0; j < NN; ++j)
#pragma unroll
    for (int e = 0; e < NL; ++e) dphi[j][e] = 0.0;
#pragma unroll
  for (int a = 0; a < NL; ++a) dphi[a][a] = 4.0 * l[a] - 1.0;
  if constexpr (D == 2) {
    dphi[3][1] = 4.0 * l[2]; dphi[3][2] = 4.0 * l[1];
    dphi[4][0] = 4.0 * l[2]; dphi[4][2] = 4.0 * l[0];
    dphi[5][0] = 4.0 * l[1]; dphi[5][1] = 4.0 * l[0];
  } else {
    dphi[4][2] = 4.0 * l[3]; dphi[4][3] = 4.0 * l[2];
    dphi[5][1] = 4.0 * l[3]; dphi[5][3] = 4.0 * l[1];
    dphi[6][1] = 4.0 * l[2]; dphi[6][2] = 4.0 * l[1];
    dphi[7][0] = 4.0 * l[3]; dphi[7][3] = 4.0 * l[0];
    dphi[8][0] = 4.0 * l[2]; dphi[8][2] = 4.0 * l[0];
    dphi[9][0] = 4.0 * l[1]; dphi[9][1] = 4.0 * l[0];
  }
}

// b[(a,i)] += scale beta int (T_h - T_ref) d(phi_a)/dx_i = scale beta |T| sum_j dT_j sum_e S[j][a][e] G_e[i]
template <int D>
__global__ void k_p2_thermal_load(int64_t ncells, const int32_t* __restrict__ cell_nodes, const double* __restrict__ xyz,
                                  const double* __restrict__ tS, const double* __restrict__ T, double T_const, double T_ref,
                                  double w, double* __restrict__ b) {
  constexpr int NL = D + 1, NN = (D + 1) * (D + 2) / 2;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncells; c += (int64_t)gridDim.x * blockDim.x) {
    int nd[NN];
    for (int a = 0; a < NN; ++a) nd[a] = cell_nodes[c * NN + a];
    Cell<D> g;
    cell_geometry<D>(xyz, nd, g);
    double dT[NN];
    for (int j = 0; j < NN; ++j) dT[j] = (T ? T[nd[j]] : T_const) - T_ref;
    for (int a = 0; a < NN; ++a) {
      double se[NL];
      for (int e = 0; e < NL; ++e) {
        double s = 0.0;
        for (int j = 0; j < NN; ++j) s += dT[j] * tS[(j * NN + a) * NL + e];
        se[e] = s;
      }
      for (int i = 0; i < D; ++i) {
        double s = 0.0;
        for (int e = 0; e < NL; ++e) s += se[e] * g.G[e][i];
        atomicAdd(b + (int64_t)nd[a] * D + i, w * g.vol * s);
      }
    }
  }
}

struct QRule { int n; double l[64][4]; double w[64]; };

// b_a += int vm(u_h) lambda_a dx over the vertices a of each cell, u_h degree 2: quadrature rule `q`
template <int D>
__global__ void k_p2_von_mises_load(int64_t ncells, const int32_t* __restrict__ cell_nodes, const double* __restrict__ xyz,
                                    const double* __restrict__ u, double mu, double lambda, const QRule* __restrict__ q,
                                    double* __restrict__ b) {
  constexpr int NL = D + 1, NN = (D + 1) * (D + 2) / 2;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncells; c += (int64_t)gridDim.x * blockDim.x) {
    int nd[NN];
    for (int a = 0; a < NN; ++a) nd[a] = cell_nodes[c * NN + a];
    Cell<D> g;
    cell_geometry<D>(xyz, nd, g);
    double ul[NN][D];
    for (int j = 0; j < NN; ++j)
      for (int i = 0; i < D; ++i) ul[j][i] = u[(int64_t)nd[j] * D + i];
    double acc[NL];
    for (int a = 0; a < NL; ++a) acc[a] = 0.0;
    for (int p = 0; p < q->n; ++p) {
      double l[NL];
      for (int a = 0; a < NL; ++a) l[a] = q->l[p][a];
      double dphi[NN][NL];
      p2_dphi<D>(l, dphi);
      double H[D][D];                       // grad u [i][k] = sum_j u_j[i] sum_e dphi_j/dl_e G_e[k]
      for (int i = 0; i < D; ++i)
        for (int k = 0; k < D; ++k) H[i][k] = 0.0;
      for (int j = 0; j < NN; ++j) {
        double gj[D];
        for (int k = 0; k < D; ++k) {
          double s = 0.0;
          for (int e = 0; e < NL; ++e) s += dphi[j][e] * g.G[e][k];
          gj[k] = s;
        }
        for (int i = 0; i < D; ++i)
          for (int k = 0; k < D; ++k) H[i][k] += ul[j][i] * gj[k];
      }
      const double vm = fsb_von_mises<D>(H, mu, lambda);
      for (int a = 0; a < NL; ++a) acc[a] += q->w[p] * vm * l[a];
    }
    for (int a = 0; a < NL; ++a) atomicAdd(b + nd[a], g.vol * acc[a]);
  }
}

void fill_form(Form& f, int D, double kscale, const double* ktensor, double mass, double adv, const double* vel) {
  memset(&f, 0, sizeof(f));
  f.kscale = kscale; f.mass = mass; f.adv = adv;
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < D; ++j) f.K[i * D + j] = ktensor ? ktensor[i * D + j] : (i == j ? 1.0 : 0.0);
  if (vel) for (int i = 0; i < D; ++i) f.vel[i] = vel[i];
}

}  // namespace

// ------------------------------------------------------------------------------------ entry points used by fsb_assemble.cu
int fsb_p2_scalar(fsb_mesh* mesh, fsb_mat* A, fsb_vec* x, fsb_vec* y, double kscale, const double* ktensor, double mass, double adv,
                  const double* vel) {
  fsb_ctx* ctx = mesh->ctx;
  int rc = ensure_tables(mesh);
  if (rc) return rc;
  Form f;
  fill_form(f, mesh->tdim, kscale, ktensor, mass, adv, vel);
  const unsigned grid = fsb_grid(mesh->ncells, 128, (int64_t)ctx->sm_count * 64);
  const bool action = A == nullptr;
  const uint8_t* pm = (!action && ctx->asm_mode >= 1 && A->mesh == mesh) ? A->posmap : nullptr;
  if (mesh->tdim == 3) {
    if (action) k_p2_scalar<3, true><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cell_nodes, mesh->xyz, f, mesh->p2_tables, nullptr, nullptr, nullptr, nullptr, x->d, y->d);
    else k_p2_scalar<3, false><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cell_nodes, mesh->xyz, f, mesh->p2_tables, A->row_ptr, A->col_idx, A->vals, pm, nullptr, nullptr);
  } else {
    if (action) k_p2_scalar<2, true><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cell_nodes, mesh->xyz, f, mesh->p2_tables, nullptr, nullptr, nullptr, nullptr, x->d, y->d);
    else k_p2_scalar<2, false><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cell_nodes, mesh->xyz, f, mesh->p2_tables, A->row_ptr, A->col_idx, A->vals, pm, nullptr, nullptr);
  }
  FSB_LAUNCH_CHECK(ctx);
  return FSB_OK;
}

int fsb_p2_elasticity(fsb_mesh* mesh, fsb_mat* A, double mu, double lambda) {
  fsb_ctx* ctx = mesh->ctx;
  int rc = ensure_tables(mesh);
  if (rc) return rc;
  const unsigned grid = fsb_grid(mesh->ncells, 128, (int64_t)ctx->sm_count * 64);
  const uint8_t* pm = (ctx->asm_mode >= 1 && A->mesh == mesh) ? A->posmap : nullptr;
  if (mesh->tdim == 3) k_p2_elasticity<3><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cell_nodes, mesh->xyz, mu, lambda, mesh->p2_tables, A->row_ptr, A->col_idx, A->vals, pm);
  else k_p2_elasticity<2><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cell_nodes, mesh->xyz, mu, lambda, mesh->p2_tables, A->row_ptr, A->col_idx, A->vals, pm);
  FSB_LAUNCH_CHECK(ctx);
  return FSB_OK;
}

int fsb_p2_source(fsb_mesh* mesh, double* b, int ncomp, const double* S_const, const double* S_nodal, double scale, const int32_t* d_tags,
                  int tag) {
  fsb_ctx* ctx = mesh->ctx;
  int rc = ensure_tables(mesh);
  if (rc) return rc;
  P2Layout L(mesh->tdim);
  const unsigned grid = fsb_grid(mesh->ncells, 128, (int64_t)ctx->sm_count * 64);
  if (S_nodal) {
    if (mesh->tdim == 3) k_p2_source_nodal<3><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cell_nodes, mesh->xyz, ncomp, S_nodal, scale, mesh->p2_tables + L.oM, b);
    else k_p2_source_nodal<2><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cell_nodes, mesh->xyz, ncomp, S_nodal, scale, mesh->p2_tables + L.oM, b);
  } else {
    V3 s{{0, 0, 0}};
    for (int k = 0; k < ncomp; ++k) s.v[k] = S_const[k];
    if (mesh->tdim == 3) k_p2_source_const<3><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cell_nodes, mesh->xyz, ncomp, s, scale, mesh->p2_tables + L.oF, d_tags, tag, b);
    else k_p2_source_const<2><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cell_nodes, mesh->xyz, ncomp, s, scale, mesh->p2_tables + L.oF, d_tags, tag, b);
  }
  FSB_LAUNCH_CHECK(ctx);
  return FSB_OK;
}

int fsb_p2_facet_load(fsb_mesh* mesh, double* b, int ncomp, int64_t nf, const int32_t* d_fnodes, const int32_t* d_opp, int mode,
                      const double* g, double scale) {
  fsb_ctx* ctx = mesh->ctx;
  int rc = ensure_tables(mesh);
  if (rc) return rc;
  P2Layout L(mesh->tdim);
  V3 gv{{0, 0, 0}};
  for (int k = 0; k < (mode == 1 ? 1 : ncomp); ++k) gv.v[k] = g[k];
  const unsigned grid = fsb_grid(nf, 128, (int64_t)ctx->sm_count * 16);
  if (mesh->tdim == 3) k_p2_facet_load<3><<<grid, 128, 0, ctx->stream>>>(nf, d_fnodes, d_opp, mesh->xyz, ncomp, mode, gv, scale, mesh->p2_tables + L.oFf, b);
  else k_p2_facet_load<2><<<grid, 128, 0, ctx->stream>>>(nf, d_fnodes, d_opp, mesh->xyz, ncomp, mode, gv, scale, mesh->p2_tables + L.oFf, b);
  FSB_LAUNCH_CHECK(ctx);
  return FSB_OK;
}

int fsb_p2_facet_mass(fsb_mesh* mesh, fsb_mat* A, int64_t nf, const int32_t* d_fnodes, double h) {
  fsb_ctx* ctx = mesh->ctx;
  int rc = ensure_tables(mesh);
  if (rc) return rc;
  P2Layout L(mesh->tdim);
  const unsigned grid = fsb_grid(nf, 128, (int64_t)ctx->sm_count * 16);
  if (mesh->tdim == 3) k_p2_facet_mass<3><<<grid, 128, 0, ctx->stream>>>(nf, d_fnodes, mesh->xyz, h, mesh->p2_tables + L.oMf, A->row_ptr, A->col_idx, A->vals);
  else k_p2_facet_mass<2><<<grid, 128, 0, ctx->stream>>>(nf, d_fnodes, mesh->xyz, h, mesh->p2_tables + L.oMf, A->row_ptr, A->col_idx, A->vals);
  FSB_LAUNCH_CHECK(ctx);
  return FSB_OK;
}

void fsb_simplex_rule(int d, int n, std::vector<double>& bary, std::vector<double>& w) {
  bary.clear(); w.clear();
  integrate_simplex(d, n, [&](const double* l, double wt) {
    for (int a = 0; a <= d; ++a) bary.push_back(l[a]);
    w.push_back(wt);
  });
}

int fsb_p2_thermal_load(fsb_mesh* mesh, double* b, const double* T, double T_const, double T_ref, double w) {
  fsb_ctx* ctx = mesh->ctx;
  int rc = ensure_tables(mesh);
  if (rc) return rc;
  P2Layout L(mesh->tdim);
  const unsigned grid = fsb_grid(mesh->ncells, 128, (int64_t)ctx->sm_count * 64);
  if (mesh->tdim == 3) k_p2_thermal_load<3><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cell_nodes, mesh->xyz, mesh->p2_tables + L.oS, T, T_const, T_ref, w, b);
  else k_p2_thermal_load<2><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cell_nodes, mesh->xyz, mesh->p2_tables + L.oS, T, T_const, T_ref, w, b);
  FSB_LAUNCH_CHECK(ctx);
  return FSB_OK;
}

// the projection integrand sqrt(3/2 s:s) is not a polynomial: a collapsed Gauss rule with 4 points per axis
// (exact to degree 5, the degree UFL estimates for this integrand on a degree-2 displacement)
int fsb_p2_von_mises_load(fsb_mesh* mesh, const double* u, double mu, double lambda, double* b) {
  fsb_ctx* ctx = mesh->ctx;
  const int D = mesh->tdim;
  std::vector<double> bary, w;
  fsb_simplex_rule(D, 4, bary, w);
  QRule q;
  memset(&q, 0, sizeof(q));
  q.n = (int)w.size();
  for (int p = 0; p < q.n; ++p) {
    for (int a = 0; a <= D; ++a) q.l[p][a] = bary[(size_t)p * (D + 1) + a];
    q.w[p] = w[p];
  }
  QRule* dq = nullptr;
  int rc = fsb_dmalloc(ctx, &dq, 1);
  if (rc) return rc;
  FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(dq, &q, sizeof(q), cudaMemcpyHostToDevice, ctx->stream));
  const unsigned grid = fsb_grid(mesh->ncells, 64, (int64_t)ctx->sm_count * 64);
  if (D == 3) k_p2_von_mises_load<3><<<grid, 64, 0, ctx->stream>>>(mesh->ncells, mesh->cell_nodes, mesh->xyz, u, mu, lambda, dq, b);
  else k_p2_von_mises_load<2><<<grid, 64, 0, ctx->stream>>>(mesh->ncells, mesh->cell_nodes, mesh->xyz, u, mu, lambda, dq, b);
  FSB_LAUNCH_CHECK(ctx);
  FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));     // q lives on the host stack
  fsb_dfree(ctx, dq);
  return FSB_OK;
}
