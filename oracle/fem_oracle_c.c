/* fem_oracle_c.c — plain-C (OpenMP) restatement of the 3D heat hot path for CPU-scale baselines.
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY: built into oracle/_build/libfem_oracle.so by oracle/Makefile
 * and loaded by tests/ (pinned against oracle/fem_oracle.py, which is pinned against the reference's
 * fixtures) and by bench.py's cpu_baseline / --impl reference legs.  Nothing under
 * fenicssolver_b200/ links or loads it.  PARITY UNPINNED at the dolfin boundary (see fem_oracle.py).
 *
 * It follows what the reference's scalar path makes dolfin/PETSc do, on all host cores:
 *   mesh          UnitCubeMesh/BoxMesh layout             examples/test_heat_transfer.py:33-34
 *   pattern       SparsityPatternBuilder, sorted columns  SolverBase.py:608-612 (inside assemble)
 *   assemble      cell loop, K_e = |T| k G G^T, ADD_VALUES ScalarTransportSolver.py:284-285
 *   source        |T|/4 S per vertex                       ScalarTransportSolver.py:301-303
 *   Dirichlet     assemble_system (symmetric elimination)  SolverBase.py:644
 *   solve         Jacobi-preconditioned CG (KSPCG+PCJACOBI) SolverBase.py:663-670
 * with the same recurrences as fem_oracle.pcg_jacobi and the CUDA path, so iteration counts agree.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int fo_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* launchers such as torchrun export OMP_NUM_THREADS=1 into every rank; the CPU baseline is meant to use every host core it
 * is allowed to run on, so the caller sets the count explicitly (bench.py passes the size of the affinity mask) */
void fo_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

static const int HEX_TETS[6][4] = {{0, 1, 3, 7}, {0, 1, 5, 7}, {0, 4, 5, 7}, {0, 2, 3, 7}, {0, 4, 6, 7}, {0, 2, 6, 7}};

/* coords[nverts][3], cells[ncells][4] (sorted per cell) in the dolfin BoxMesh layout */
void fo_box_mesh(const int32_t n[3], const double p0[3], const double p1[3], double* coords, int32_t* cells) {
  const int64_t px = n[0] + 1, py = px * (n[1] + 1);
  const int64_t nverts = py * (n[2] + 1);
#pragma omp parallel for schedule(static)
  for (int64_t v = 0; v < nverts; ++v) {
    int64_t iz = v / py, rem = v % py, iy = rem / px, ix = rem % px;
    coords[3 * v + 0] = p0[0] + (double)ix * (p1[0] - p0[0]) / n[0];
    coords[3 * v + 1] = p0[1] + (double)iy * (p1[1] - p0[1]) / n[1];
    coords[3 * v + 2] = p0[2] + (double)iz * (p1[2] - p0[2]) / n[2];
  }
  const int64_t nhex = (int64_t)n[0] * n[1] * n[2];
#pragma omp parallel for schedule(static)
  for (int64_t h = 0; h < nhex; ++h) {
    int64_t cz = h / ((int64_t)n[0] * n[1]), rem = h % ((int64_t)n[0] * n[1]), cy = rem / n[0], cx = rem % n[0];
    int64_t v0 = cx + cy * px + cz * py;
    int64_t c[8] = {v0, v0 + 1, v0 + px, v0 + px + 1, v0 + py, v0 + py + 1, v0 + py + px, v0 + py + px + 1};
    for (int k = 0; k < 6; ++k)
      for (int a = 0; a < 4; ++a) cells[(6 * h + k) * 4 + a] = (int32_t)c[HEX_TETS[k][a]];
  }
}

static int cmp_i32(const void* a, const void* b) {
  int32_t x = *(const int32_t*)a, y = *(const int32_t*)b;
  return (x > y) - (x < y);
}

/* Pattern, pass 1: row_ptr[nverts+1] (returns nnz); pass 2 (col_idx != NULL): sorted unique columns.
 * v2c_ptr / v2c are the vertex->cell adjacency built by fo_vertex_cells. */
void fo_vertex_cells(int64_t nverts, int64_t ncells, const int32_t* cells, int64_t* v2c_ptr, int32_t* v2c) {
  memset(v2c_ptr, 0, sizeof(int64_t) * (nverts + 1));
  for (int64_t i = 0; i < ncells * 4; ++i) v2c_ptr[cells[i] + 1]++;
  for (int64_t v = 0; v < nverts; ++v) v2c_ptr[v + 1] += v2c_ptr[v];
  int64_t* cur = (int64_t*)malloc(sizeof(int64_t) * nverts);
  memcpy(cur, v2c_ptr, sizeof(int64_t) * nverts);
  for (int64_t c = 0; c < ncells; ++c)
    for (int a = 0; a < 4; ++a) v2c[cur[cells[4 * c + a]]++] = (int32_t)c;
  free(cur);
}

int64_t fo_csr_pattern(int64_t nverts, const int32_t* cells, const int64_t* v2c_ptr, const int32_t* v2c,
                       int64_t* row_ptr, int32_t* col_idx) {
  const int fill = col_idx != NULL;
#pragma omp parallel
  {
    int cap = 256;
    int32_t* buf = (int32_t*)malloc(sizeof(int32_t) * cap);
#pragma omp for schedule(static)
    for (int64_t r = 0; r < nverts; ++r) {
      int64_t m = (v2c_ptr[r + 1] - v2c_ptr[r]) * 4;
      if (m > cap) { cap = (int)m * 2; buf = (int32_t*)realloc(buf, sizeof(int32_t) * cap); }
      int k = 0;
      for (int64_t p = v2c_ptr[r]; p < v2c_ptr[r + 1]; ++p)
        for (int a = 0; a < 4; ++a) buf[k++] = cells[4 * (int64_t)v2c[p] + a];
      qsort(buf, k, sizeof(int32_t), cmp_i32);
      int u = 0;
      for (int i = 0; i < k; ++i)
        if (i == 0 || buf[i] != buf[i - 1]) buf[u++] = buf[i];
      if (!fill) row_ptr[r + 1] = u;
      else memcpy(col_idx + row_ptr[r], buf, sizeof(int32_t) * u);
    }
    free(buf);
  }
  if (!fill) {
    row_ptr[0] = 0;
    for (int64_t r = 0; r < nverts; ++r) row_ptr[r + 1] += row_ptr[r];
  }
  return row_ptr[nverts];
}

static inline int64_t find_col(const int32_t* cols, int64_t lo, int64_t hi, int32_t c) {
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (cols[mid] < c) lo = mid + 1; else hi = mid;
  }
  return lo;
}

/* vals += k * stiffness ; b += S * load   (one cell loop, like dolfin's assemble of a and L) */
void fo_assemble_heat(int64_t ncells, const int32_t* cells, const double* coords, double k, double S,
                      const int64_t* row_ptr, const int32_t* col_idx, double* vals, double* b) {
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < ncells; ++c) {
    const int32_t* v = cells + 4 * c;
    const double* x0 = coords + 3 * (int64_t)v[0];
    double a[3], bb[3], cc[3];
    for (int i = 0; i < 3; ++i) {
      a[i] = coords[3 * (int64_t)v[1] + i] - x0[i];
      bb[i] = coords[3 * (int64_t)v[2] + i] - x0[i];
      cc[i] = coords[3 * (int64_t)v[3] + i] - x0[i];
    }
    double bc[3] = {bb[1] * cc[2] - bb[2] * cc[1], bb[2] * cc[0] - bb[0] * cc[2], bb[0] * cc[1] - bb[1] * cc[0]};
    double ca[3] = {cc[1] * a[2] - cc[2] * a[1], cc[2] * a[0] - cc[0] * a[2], cc[0] * a[1] - cc[1] * a[0]};
    double ab[3] = {a[1] * bb[2] - a[2] * bb[1], a[2] * bb[0] - a[0] * bb[2], a[0] * bb[1] - a[1] * bb[0]};
    double det = a[0] * bc[0] + a[1] * bc[1] + a[2] * bc[2], inv = 1.0 / det;
    double G[4][3];
    for (int i = 0; i < 3; ++i) {
      G[1][i] = bc[i] * inv; G[2][i] = ca[i] * inv; G[3][i] = ab[i] * inv;
      G[0][i] = -(G[1][i] + G[2][i] + G[3][i]);
    }
    const double vol = fabs(det) / 6.0, kw = k * vol, sw = S * vol * 0.25;
    for (int p = 0; p < 4; ++p) {
      const int64_t base = row_ptr[v[p]], end = row_ptr[v[p] + 1];
      int64_t lo = base;
      for (int q = 0; q < 4; ++q) {
        lo = find_col(col_idx, lo, end, v[q]);
        const double e = kw * (G[p][0] * G[q][0] + G[p][1] * G[q][1] + G[p][2] * G[q][2]);
#pragma omp atomic
        vals[lo] += e;
        ++lo;
      }
      if (b) {
#pragma omp atomic
        b[v[p]] += sw;
      }
    }
  }
}

/* assemble_system-style symmetric Dirichlet elimination; flag/g are dense per-dof arrays */
void fo_apply_dirichlet_sym(int64_t n, const int64_t* row_ptr, const int32_t* col_idx, double* vals, double* b,
                            const uint8_t* flag, const double* g) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < n; ++r) {
    if (flag[r]) {
      for (int64_t k = row_ptr[r]; k < row_ptr[r + 1]; ++k) vals[k] = (col_idx[k] == r) ? 1.0 : 0.0;
      b[r] = g[r];
    } else {
      double corr = 0.0;
      for (int64_t k = row_ptr[r]; k < row_ptr[r + 1]; ++k)
        if (flag[col_idx[k]]) { corr += vals[k] * g[col_idx[k]]; vals[k] = 0.0; }
      b[r] -= corr;
    }
  }
}

static void spmv(int64_t n, const int64_t* rp, const int32_t* ci, const double* va, const double* x, double* y) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < n; ++r) {
    double s = 0.0;
    for (int64_t k = rp[r]; k < rp[r + 1]; ++k) s += va[k] * x[ci[k]];
    y[r] = s;
  }
}

void fo_spmv(int64_t n, const int64_t* rp, const int32_t* ci, const double* va, const double* x, double* y) { spmv(n, rp, ci, va, x, y); }

/* Jacobi-PCG; x holds the start vector.  Returns the iteration count; *relres = ||r||/||b||.
 * stops at ||M^-1 r|| <= max(rtol*||M^-1 b||, atol) (preconditioned norm, PETSc's KSP default) or after maxit. */
int fo_pcg_jacobi(int64_t n, const int64_t* rp, const int32_t* ci, const double* va, const double* b, double* x,
                  double rtol, double atol, int maxit, double* relres) {
  double* r = (double*)malloc(sizeof(double) * n);
  double* p = (double*)malloc(sizeof(double) * n);
  double* q = (double*)malloc(sizeof(double) * n);
  double* dinv = (double*)malloc(sizeof(double) * n);
  double rz = 0.0, rr = 0.0, bbn = 0.0;
  spmv(n, rp, ci, va, x, q);
#pragma omp parallel for schedule(static) reduction(+ : rz, rr, bbn)
  for (int64_t i = 0; i < n; ++i) {
    double d = 1.0;
    for (int64_t k = rp[i]; k < rp[i + 1]; ++k)
      if (ci[k] == i) d = va[k];
    dinv[i] = 1.0 / d;
    r[i] = b[i] - q[i];
    const double z = dinv[i] * r[i];
    p[i] = z;
    rz += r[i] * z; rr += z * z; bbn += (dinv[i] * b[i]) * (dinv[i] * b[i]);   /* preconditioned norms */
  }
  const double tol2 = fmax(rtol * rtol * bbn, atol * atol);
  int it = 0;
  while (rr > tol2 && it < maxit) {
    spmv(n, rp, ci, va, p, q);
    double pq = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : pq)
    for (int64_t i = 0; i < n; ++i) pq += p[i] * q[i];
    const double alpha = rz / pq;
    double rzn = 0.0;
    rr = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : rzn, rr)
    for (int64_t i = 0; i < n; ++i) {
      x[i] += alpha * p[i];
      const double ri = r[i] - alpha * q[i];
      r[i] = ri;
      const double zi = dinv[i] * ri;
      rzn += ri * zi;
      rr += zi * zi;
    }
    const double beta = rzn / rz;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) p[i] = dinv[i] * r[i] + beta * p[i];
    rz = rzn;
    ++it;
  }
  if (relres) *relres = bbn > 0 ? sqrt(rr / bbn) : sqrt(rr);
  free(r); free(p); free(q); free(dinv);
  return it;
}
