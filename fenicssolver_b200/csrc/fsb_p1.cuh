// P1 (degree-1 simplex) device helpers shared by the element kernels: cell load, affine geometry, entry positions,
// facet geometry.
#pragma once
#include "fsb_internal.cuh"

template <int D>
struct Geo {
  double vol;
  double G[D + 1][D];
};

template <int D>
__device__ __forceinline__ void load_cell(const int32_t* __restrict__ cells, int64_t c, int (&v)[D + 1]) {
  if constexpr (D == 3) {
    int4 q = __ldg(reinterpret_cast<const int4*>(cells) + c);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  } else {
#pragma unroll
    for (int a = 0; a <= D; ++a) v[a] = __ldg(cells + c * (D + 1) + a);
  }
}

template <int D>
__device__ __forceinline__ void p1_geometry(const double* __restrict__ xyz, const int (&v)[D + 1], Geo<D>& g) {
  double X[D + 1][D];
#pragma unroll
  for (int a = 0; a <= D; ++a)
#pragma unroll
    for (int i = 0; i < D; ++i) X[a][i] = __ldg(xyz + (int64_t)v[a] * D + i);
  if constexpr (D == 3) {
    double a[3], b[3], c[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { a[i] = X[1][i] - X[0][i]; b[i] = X[2][i] - X[0][i]; c[i] = X[3][i] - X[0][i]; }
    double bc[3] = {b[1] * c[2] - b[2] * c[1], b[2] * c[0] - b[0] * c[2], b[0] * c[1] - b[1] * c[0]};
    double ca[3] = {c[1] * a[2] - c[2] * a[1], c[2] * a[0] - c[0] * a[2], c[0] * a[1] - c[1] * a[0]};
    double ab[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    double det = a[0] * bc[0] + a[1] * bc[1] + a[2] * bc[2];
    double inv = 1.0 / det;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      g.G[1][i] = bc[i] * inv; g.G[2][i] = ca[i] * inv; g.G[3][i] = ab[i] * inv;
      g.G[0][i] = -(g.G[1][i] + g.G[2][i] + g.G[3][i]);
    }
    g.vol = fabs(det) * (1.0 / 6.0);
  } else {
    double a0 = X[1][0] - X[0][0], a1 = X[1][1] - X[0][1], b0 = X[2][0] - X[0][0], b1 = X[2][1] - X[0][1];
    double det = a0 * b1 - a1 * b0, inv = 1.0 / det;
    g.G[1][0] = b1 * inv; g.G[1][1] = -b0 * inv;
    g.G[2][0] = -a1 * inv; g.G[2][1] = a0 * inv;
    g.G[0][0] = -(g.G[1][0] + g.G[2][0]); g.G[0][1] = -(g.G[1][1] + g.G[2][1]);
    g.vol = fabs(det) * 0.5;
  }
}

// positions of the (D+1)^2 local entries: offset of column v[b] inside row v[a]
template <int D>
__device__ __forceinline__ void entry_positions(const uint8_t* __restrict__ posmap, int64_t c, const int (&v)[D + 1],
                                                const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx,
                                                int64_t (&base)[D + 1], int (&pos)[D + 1][D + 1]) {
  constexpr int NL = D + 1;
#pragma unroll
  for (int a = 0; a < NL; ++a) base[a] = __ldg(row_ptr + v[a]);
  if (posmap) {
    if constexpr (D == 3) {
      uint4 q = __ldg(reinterpret_cast<const uint4*>(posmap) + c);
      unsigned w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int a = 0; a < NL; ++a)
#pragma unroll
        for (int b = 0; b < NL; ++b) pos[a][b] = (w[a] >> (8 * b)) & 0xff;
    } else {
#pragma unroll
      for (int a = 0; a < NL; ++a)
#pragma unroll
        for (int b = 0; b < NL; ++b) pos[a][b] = __ldg(posmap + c * NL * NL + a * NL + b);
    }
  } else {
#pragma unroll
    for (int a = 0; a < NL; ++a) {
      const int len = (int)(__ldg(row_ptr + v[a] + 1) - base[a]);
      int lo = 0;
#pragma unroll
      for (int b = 0; b < NL; ++b) {
        lo = row_find(col_idx + base[a], lo, len, v[b]);
        pos[a][b] = lo++;
      }
    }
  }
}

// facet measure (edge length / triangle area) and the un-normalised normal
template <int D>
__device__ __forceinline__ double facet_geom(const double* __restrict__ xyz, const int32_t* fv, double (&n)[3], double (&x0)[3]) {
  if constexpr (D == 3) {
    double p[3][3];
    for (int a = 0; a < 3; ++a) for (int i = 0; i < 3; ++i) p[a][i] = xyz[(int64_t)fv[a] * 3 + i];
    double a[3] = {p[1][0] - p[0][0], p[1][1] - p[0][1], p[1][2] - p[0][2]};
    double b[3] = {p[2][0] - p[0][0], p[2][1] - p[0][1], p[2][2] - p[0][2]};
    n[0] = a[1] * b[2] - a[2] * b[1]; n[1] = a[2] * b[0] - a[0] * b[2]; n[2] = a[0] * b[1] - a[1] * b[0];
    for (int i = 0; i < 3; ++i) x0[i] = p[0][i];
    return 0.5 * sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
  } else {
    double t0 = xyz[(int64_t)fv[1] * 2] - xyz[(int64_t)fv[0] * 2], t1 = xyz[(int64_t)fv[1] * 2 + 1] - xyz[(int64_t)fv[0] * 2 + 1];
    n[0] = t1; n[1] = -t0; n[2] = 0.0;
    x0[0] = xyz[(int64_t)fv[0] * 2]; x0[1] = xyz[(int64_t)fv[0] * 2 + 1]; x0[2] = 0.0;
    return sqrt(t0 * t0 + t1 * t1);
  }
}

