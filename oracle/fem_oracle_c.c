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

/* parallel zero fill with the static schedule the other loops use: first touch places the pages next to the threads that
 * will stream them (numpy's own fill would put a whole array on one NUMA node and run on one core) */
void fo_zero(double* a, int64_t n) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) a[i] = 0.0;
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

/* The same Jacobi-PCG in resumable form, for bench.py --impl reference: ONE complete solve is run as K consecutive segments (the
 * driver's "steps"), the recurrence state carried in `work` (4n doubles: r, p, q, dinv) and `state` (8 doubles: [0] initialised,
 * [1] r.z, [2] z.z, [3] ||M^-1 b||^2, [4] iterations so far).  Runs at most `seg_iters` iterations; returns 1 once converged.
 * The arithmetic (and therefore the iteration count) is fo_pcg_jacobi's. */
int fo_pcg_jacobi_segment(int64_t n, const int64_t* rp, const int32_t* ci, const double* va, const double* b, double* x, double* work,
                          double* state, double rtol, double atol, int seg_iters) {
  double *r = work, *p = work + n, *q = work + 2 * n, *dinv = work + 3 * n;
  double rz = state[1], rr = state[2], bbn = state[3];
  if (state[0] == 0.0) {
    rz = rr = bbn = 0.0;
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
      rz += r[i] * z; rr += z * z; bbn += (dinv[i] * b[i]) * (dinv[i] * b[i]);
    }
    state[0] = 1.0; state[4] = 0.0;
  }
  const double tol2 = fmax(rtol * rtol * bbn, atol * atol);
  int it = 0;
  while (rr > tol2 && it < seg_iters) {
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
  state[1] = rz; state[2] = rr; state[3] = bbn; state[4] += it;
  return rr <= tol2;
}

static inline double tet_geometry(const int32_t* v, const double* coords, double G[4][3]);
static inline double tet_geometry_fwd(const int32_t* v, const double* coords, double G[4][3]) { return tet_geometry(v, coords, G); }

/* Degree-2 (P2) heat on tetrahedra, for the bench's `p2` block: vals += k |T| sum_{c,e} (G_c . G_e) R[i][j][c][e], b += S |T| F[i].
 * cell_nodes[nc][10] = the cell's 4 vertices (sorted) then its 6 edge nodes in UFC order; R[10][10][4][4] and F[10] are the exact
 * reference tensors of oracle/fem_oracle_p2.py reference_tensors(3) (this function is checked against that module's assembly in
 * tests/test_oracle_p2.py).  The affine geometry is fo_assemble_heat's. */
void fo_assemble_heat_p2(int64_t ncells, const int32_t* cell_nodes, const double* coords, const double* R, const double* F, double k, double S,
                         const int64_t* row_ptr, const int32_t* col_idx, double* vals, double* b) {
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < ncells; ++c) {
    const int32_t* nd = cell_nodes + 10 * c;
    double G[4][3];
    const double vol = tet_geometry_fwd(nd, coords, G);
    double gg[4][4];
    for (int p = 0; p < 4; ++p)
      for (int q = 0; q < 4; ++q) gg[p][q] = G[p][0] * G[q][0] + G[p][1] * G[q][1] + G[p][2] * G[q][2];
    for (int i = 0; i < 10; ++i) {
      const int64_t base = row_ptr[nd[i]], end = row_ptr[nd[i] + 1];
      for (int j = 0; j < 10; ++j) {
        const double* r = R + ((i * 10 + j) * 16);
        double e = 0.0;
        for (int p = 0; p < 4; ++p)
          for (int q = 0; q < 4; ++q) e += gg[p][q] * r[p * 4 + q];
        const int64_t pos = find_col(col_idx, base, end, nd[j]);
#pragma omp atomic
        vals[pos] += k * vol * e;
      }
      if (b) {
#pragma omp atomic
        b[nd[i]] += S * vol * F[i];
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------------------------------------
 * BASELINE configs C3 (elasticity) and C4 (transient advection-diffusion): the cell loops and BiCGStab, so that the bench's
 * `c3` / `c4` blocks have a CPU figure and a full-size oracle to be compared with.  Same closed-form P1 element matrices as
 * oracle/fem_oracle.py (local_laplace / local_mass / local_advection / local_elasticity), checked against it in
 * tests/test_oracle_kat.py. */
static inline double tet_geometry(const int32_t* v, const double* coords, double G[4][3]) {
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
  for (int i = 0; i < 3; ++i) {
    G[1][i] = bc[i] * inv; G[2][i] = ca[i] * inv; G[3][i] = ab[i] * inv;
    G[0][i] = -(G[1][i] + G[2][i] + G[3][i]);
  }
  return fabs(det) / 6.0;
}

/* vals += kscale K + mass M + adv C(vel)   (ScalarTransportSolver.py:284-285, 292, 311; vel may be NULL when adv == 0) */
void fo_assemble_scalar(int64_t ncells, const int32_t* cells, const double* coords, double kscale, double mass, double adv,
                        const double* vel, const int64_t* row_ptr, const int32_t* col_idx, double* vals) {
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < ncells; ++c) {
    const int32_t* v = cells + 4 * c;
    double G[4][3];
    const double vol = tet_geometry(v, coords, G);
    double vg[4] = {0, 0, 0, 0};
    if (adv != 0.0 && vel)
      for (int q = 0; q < 4; ++q) vg[q] = vel[0] * G[q][0] + vel[1] * G[q][1] + vel[2] * G[q][2];
    for (int p = 0; p < 4; ++p) {
      const int64_t end = row_ptr[v[p] + 1];
      int64_t lo = row_ptr[v[p]];
      for (int q = 0; q < 4; ++q) {
        lo = find_col(col_idx, lo, end, v[q]);
        const double e = vol * (kscale * (G[p][0] * G[q][0] + G[p][1] * G[q][1] + G[p][2] * G[q][2]) + mass * (p == q ? 0.1 : 0.05) +
                                adv * 0.25 * vg[q]);
#pragma omp atomic
        vals[lo] += e;
        ++lo;
      }
    }
  }
}

/* y += (kscale K + mass M) x, cell by cell: the explicit half of Crank-Nicolson (ScalarTransportSolver.py:292-293) */
void fo_apply_scalar(int64_t ncells, const int32_t* cells, const double* coords, double kscale, double mass, const double* x, double* y) {
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < ncells; ++c) {
    const int32_t* v = cells + 4 * c;
    double G[4][3];
    const double vol = tet_geometry(v, coords, G);
    double xe[4], gx[3] = {0, 0, 0}, sx = 0.0;
    for (int q = 0; q < 4; ++q) {
      xe[q] = x[v[q]];
      sx += xe[q];
      for (int i = 0; i < 3; ++i) gx[i] += G[q][i] * xe[q];
    }
    for (int p = 0; p < 4; ++p) {
      const double e = vol * (kscale * (G[p][0] * gx[0] + G[p][1] * gx[1] + G[p][2] * gx[2]) + mass * 0.05 * (sx + xe[p]));
#pragma omp atomic
      y[v[p]] += e;
    }
  }
}

/* bc.apply(A, b): zero row, unit diagonal, b = g */
void fo_apply_dirichlet_nonsym(int64_t n, const int64_t* row_ptr, const int32_t* col_idx, double* vals, double* b, const uint8_t* flag,
                               const double* g) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < n; ++r)
    if (flag[r]) {
      for (int64_t k = row_ptr[r]; k < row_ptr[r + 1]; ++k) vals[k] = (col_idx[k] == r) ? 1.0 : 0.0;
      b[r] = g[r];
    }
}

/* right-Jacobi BiCGStab, r0_hat = r0, convergence on ||M^-1 r|| (oracle/fem_oracle.py bicgstab_jacobi) */
int fo_bicgstab_jacobi(int64_t n, const int64_t* rp, const int32_t* ci, const double* va, const double* b, double* x, double rtol,
                       double atol, int maxit, double* relres) {
  double* buf = (double*)malloc(sizeof(double) * n * 8);
  double *r = buf, *rhat = buf + n, *p = buf + 2 * n, *ph = buf + 3 * n, *v = buf + 4 * n, *sh = buf + 5 * n, *t = buf + 6 * n, *dinv = buf + 7 * n;
  spmv(n, rp, ci, va, x, v);
  double bn = 0.0, rn = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : bn, rn)
  for (int64_t i = 0; i < n; ++i) {
    double d = 1.0;
    for (int64_t k = rp[i]; k < rp[i + 1]; ++k)
      if (ci[k] == i) d = va[k];
    dinv[i] = 1.0 / d;
    r[i] = b[i] - v[i];
    rhat[i] = r[i];
    p[i] = 0.0; v[i] = 0.0;
    bn += (dinv[i] * b[i]) * (dinv[i] * b[i]);
    rn += (dinv[i] * r[i]) * (dinv[i] * r[i]);
  }
  const double tol2 = fmax(rtol * rtol * bn, atol * atol);
  double rho = 1.0, alpha = 1.0, omega = 1.0;
  int it = 0;
  while (rn > tol2 && it < maxit) {
    double rho_new = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : rho_new)
    for (int64_t i = 0; i < n; ++i) rho_new += rhat[i] * r[i];
    const double beta = (rho_new / rho) * (alpha / omega);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      p[i] = r[i] + beta * (p[i] - omega * v[i]);
      ph[i] = dinv[i] * p[i];
    }
    spmv(n, rp, ci, va, ph, v);
    double rv = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : rv)
    for (int64_t i = 0; i < n; ++i) rv += rhat[i] * v[i];
    alpha = rho_new / rv;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      r[i] -= alpha * v[i];           /* s */
      sh[i] = dinv[i] * r[i];
    }
    spmv(n, rp, ci, va, sh, t);
    double ts = 0.0, tt = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : ts, tt)
    for (int64_t i = 0; i < n; ++i) { ts += t[i] * r[i]; tt += t[i] * t[i]; }
    omega = ts / tt;
    rn = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : rn)
    for (int64_t i = 0; i < n; ++i) {
      x[i] += alpha * ph[i] + omega * sh[i];
      r[i] -= omega * t[i];
      rn += (dinv[i] * r[i]) * (dinv[i] * r[i]);
    }
    rho = rho_new;
    ++it;
  }
  if (relres) *relres = bn > 0 ? sqrt(rn / bn) : sqrt(rn);
  free(buf);
  return it;
}

/* vals += int sigma(u):grad(v), sigma = 2 mu sym(grad u) + lambda div(u) I, on the scalar CSR of the 3-component space
 * (dof = 3 v + i; LinearElasticitySolver.py:62-69, 215); b += vol/4 * f per vertex (body force, may be NULL) */
void fo_assemble_elasticity(int64_t ncells, const int32_t* cells, const double* coords, double mu, double lambda, const double* f,
                            const int64_t* row_ptr, const int32_t* col_idx, double* vals, double* b) {
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < ncells; ++c) {
    const int32_t* v = cells + 4 * c;
    double G[4][3];
    const double vol = tet_geometry(v, coords, G);
    for (int p = 0; p < 4; ++p)
      for (int i = 0; i < 3; ++i) {
        const int64_t row = 3 * (int64_t)v[p] + i, end = row_ptr[row + 1];
        int64_t lo = row_ptr[row];
        for (int q = 0; q < 4; ++q) {
          const double gg = G[p][0] * G[q][0] + G[p][1] * G[q][1] + G[p][2] * G[q][2];
          for (int j = 0; j < 3; ++j) {
            lo = find_col(col_idx, lo, end, 3 * v[q] + j);
            const double e = vol * (mu * ((i == j ? gg : 0.0) + G[p][j] * G[q][i]) + lambda * G[p][i] * G[q][j]);
#pragma omp atomic
            vals[lo] += e;
            ++lo;
          }
        }
        if (b && f) {
#pragma omp atomic
          b[row] += 0.25 * vol * f[i];
        }
      }
  }
}

/* ------------------------------------------------------------------------------------------------------------------------
 * CPU restatement of the multigrid-preconditioned CG of csrc/fsb_mg.cu (scalar problems on nested box meshes), so that the
 * `gmg` block of the bench line has its own CPU figure: same transfers (fine vertex 2C + d = coarse vertex C or midpoint of the
 * coarse edge (C, C + d)), Chebyshev smoothing of degree nu on D^-1 A over [lmax/10, lmax], damped Jacobi on the coarsest level,
 * V(nu,nu) cycle, convergence on ||D^-1 r||.  Checked against oracle/mg_oracle.py (tests/test_oracle_forms.py). */
typedef struct {
  int64_t n;
  int dims[3];
  const int64_t* rp;
  const int32_t* ci;
  const double* va;
  const uint8_t* bc;
  double lmax;
  double *dinv, *x, *b, *r, *y;
} mg_level;

static void mg_dinv(mg_level* L) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < L->n; ++i) {
    double d = 1.0;
    for (int64_t k = L->rp[i]; k < L->rp[i + 1]; ++k)
      if (L->ci[k] == i) d = L->va[k];
    L->dinv[i] = d != 0.0 ? 1.0 / d : 1.0;
  }
}

/* min(1.1 * power-iteration estimate from a pseudo-random +-1 vector, Gershgorin bound) of lambda_max(D^-1 A) */
double fo_mg_lambda_max(int64_t n, const int64_t* rp, const int32_t* ci, const double* va) {
  double* dinv = (double*)malloc(sizeof(double) * n);
  double* x = (double*)malloc(sizeof(double) * n);
  double* y = (double*)malloc(sizeof(double) * n);
  double bound = 0.0;
#pragma omp parallel for schedule(static) reduction(max : bound)
  for (int64_t i = 0; i < n; ++i) {
    double d = 1.0, s = 0.0;
    for (int64_t k = rp[i]; k < rp[i + 1]; ++k) {
      if (ci[k] == i) d = va[k];
      s += fabs(va[k]);
    }
    dinv[i] = d != 0.0 ? 1.0 / d : 1.0;
    bound = fmax(bound, s * fabs(dinv[i]));
    uint64_t h = (uint64_t)i * 0x9E3779B97F4A7C15ull;
    h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
    x[i] = (h & 1) ? 1.0 : -1.0;
  }
  double lam = 0.0;
  for (int it = 0; it < 30; ++it) {
    spmv(n, rp, ci, va, x, y);
    double xx = 0.0, rr = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : xx, rr)
    for (int64_t i = 0; i < n; ++i) { y[i] *= dinv[i]; xx += x[i] * x[i]; rr += y[i] * y[i]; }
    if (!(xx > 0.0) || !(rr > 0.0)) break;
    lam = sqrt(rr / xx);
    const double s = 1.0 / sqrt(rr);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) x[i] = s * y[i];
  }
  free(dinv); free(x); free(y);
  double est = 1.1 * lam;
  if (bound > 0.0 && (est <= 0.0 || est > bound)) est = bound;
  return est > 0.0 ? est : 2.0;
}

static void mg_jacobi(mg_level* L, int sweeps) {      /* zero start, damping 4 / (3 lmax) */
  const double w = 4.0 / (3.0 * L->lmax);
  for (int s = 0; s < sweeps; ++s) {
    if (s == 0) {
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < L->n; ++i) L->x[i] = w * L->dinv[i] * L->b[i];
    } else {
      spmv(L->n, L->rp, L->ci, L->va, L->x, L->y);
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < L->n; ++i) L->x[i] += w * L->dinv[i] * (L->b[i] - L->y[i]);
    }
  }
}

static void mg_chebyshev(mg_level* L, int deg, int zero_start) {      /* L->r is the direction vector */
  const double lmax = L->lmax, lmin = lmax / 10.0;
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
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < L->n; ++i) { L->r[i] = c2 * L->dinv[i] * L->b[i]; L->x[i] = L->r[i]; }
    } else {
      spmv(L->n, L->rp, L->ci, L->va, L->x, L->y);
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < L->n; ++i) {
        const double d = (c1 != 0.0 ? c1 * L->r[i] : 0.0) + c2 * L->dinv[i] * (L->b[i] - L->y[i]);
        L->r[i] = d;
        L->x[i] += d;
      }
    }
  }
}

static void mg_vcycle(mg_level* lv, int nlevels, int l, int nu, int coarse_sweeps) {
  mg_level* L = &lv[l];
  if (l + 1 == nlevels) { mg_jacobi(L, coarse_sweeps); return; }
  mg_chebyshev(L, nu, 1);
  spmv(L->n, L->rp, L->ci, L->va, L->x, L->y);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < L->n; ++i) L->r[i] = L->b[i] - L->y[i];
  mg_level* C = &lv[l + 1];
  const int f0 = L->dims[0], f1 = L->dims[1], f2 = L->dims[2], c0 = C->dims[0], c1 = C->dims[1];
#pragma omp parallel for schedule(static)
  for (int64_t t = 0; t < C->n; ++t) {                     /* restriction, gathered per coarse vertex */
    if (C->bc && C->bc[t]) { C->b[t] = 0.0; continue; }
    const int I = (int)(t % c0), J = (int)((t / c0) % c1), K = (int)(t / ((int64_t)c0 * c1));
    const int fi = 2 * I, fj = 2 * J, fk = 2 * K;
    double s = L->r[fi + (int64_t)f0 * (fj + (int64_t)f1 * fk)];
    for (int d = 1; d < 8; ++d) {
      const int di = d & 1, dj = (d >> 1) & 1, dk = (d >> 2) & 1;
      if ((dj && f1 == 1) || (dk && f2 == 1)) continue;
      for (int sg = -1; sg <= 1; sg += 2) {
        const int i = fi + sg * di, j = fj + sg * dj, k = fk + sg * dk;
        if (i < 0 || j < 0 || k < 0 || i >= f0 || j >= f1 || k >= f2) continue;
        s += 0.5 * L->r[i + (int64_t)f0 * (j + (int64_t)f1 * k)];
      }
    }
    C->b[t] = s;
  }
  mg_vcycle(lv, nlevels, l + 1, nu, coarse_sweeps);
#pragma omp parallel for schedule(static)
  for (int64_t t = 0; t < L->n; ++t) {                     /* prolongation */
    if (L->bc && L->bc[t]) continue;
    const int i = (int)(t % f0), j = (int)((t / f0) % f1), k = (int)(t / ((int64_t)f0 * f1));
    const int di = i & 1, dj = j & 1, dk = k & 1, I = i >> 1, J = j >> 1, K = k >> 1;
    const int64_t a = I + (int64_t)c0 * (J + (int64_t)c1 * K);
    const int64_t bb = (I + di) + (int64_t)c0 * ((J + dj) + (int64_t)c1 * (K + dk));
    L->x[t] += (di | dj | dk) ? 0.5 * (C->x[a] + C->x[bb]) : C->x[a];
  }
  mg_chebyshev(L, nu, 0);
}

/* z = V(r) on level 0 (for the parity test against the numpy restatement) and the PCG driver.  dims[l][3] = vertices per axis,
 * lmax[l] = eigenvalue estimates (fo_mg_lambda_max), bc[l] = constrained-dof flags.  x holds the start vector.
 * Returns the iteration count (fo_mg_pcg) ; *relres = ||D^-1 r|| / ||D^-1 b||. */
static mg_level* mg_setup(int nlevels, const int64_t* n, const int32_t* dims, const int64_t** rp, const int32_t** ci, const double** va,
                          const uint8_t** bc, const double* lmax) {
  mg_level* lv = (mg_level*)calloc(nlevels, sizeof(mg_level));
  for (int l = 0; l < nlevels; ++l) {
    mg_level* L = &lv[l];
    L->n = n[l]; L->rp = rp[l]; L->ci = ci[l]; L->va = va[l]; L->bc = bc ? bc[l] : NULL; L->lmax = lmax[l];
    for (int a = 0; a < 3; ++a) L->dims[a] = dims[3 * l + a];
    L->dinv = (double*)malloc(sizeof(double) * L->n);
    L->x = (double*)calloc(L->n, sizeof(double));
    L->b = (double*)calloc(L->n, sizeof(double));
    L->r = (double*)calloc(L->n, sizeof(double));
    L->y = (double*)calloc(L->n, sizeof(double));
    mg_dinv(L);
  }
  return lv;
}
static void mg_free(mg_level* lv, int nlevels) {
  for (int l = 0; l < nlevels; ++l) { free(lv[l].dinv); free(lv[l].x); free(lv[l].b); free(lv[l].r); free(lv[l].y); }
  free(lv);
}

void fo_mg_apply(int nlevels, const int64_t* n, const int32_t* dims, const int64_t** rp, const int32_t** ci, const double** va,
                 const uint8_t** bc, const double* lmax, const double* r, double* z, int nu, int coarse_sweeps) {
  mg_level* lv = mg_setup(nlevels, n, dims, rp, ci, va, bc, lmax);
  memcpy(lv[0].b, r, sizeof(double) * n[0]);
  mg_vcycle(lv, nlevels, 0, nu, coarse_sweeps);
  memcpy(z, lv[0].x, sizeof(double) * n[0]);
  mg_free(lv, nlevels);
}

int fo_mg_pcg(int nlevels, const int64_t* n, const int32_t* dims, const int64_t** rp, const int32_t** ci, const double** va,
              const uint8_t** bc, const double* lmax, const double* b, double* x, double rtol, int maxit, int nu, int coarse_sweeps,
              double* relres) {
  mg_level* lv = mg_setup(nlevels, n, dims, rp, ci, va, bc, lmax);
  mg_level* L = &lv[0];
  const int64_t n0 = L->n;
  double* r = (double*)malloc(sizeof(double) * n0);
  double* p = (double*)malloc(sizeof(double) * n0);
  double* q = (double*)malloc(sizeof(double) * n0);
  double bbn = 0.0, rr = 0.0, rz = 0.0;
  spmv(n0, L->rp, L->ci, L->va, x, q);
#pragma omp parallel for schedule(static) reduction(+ : bbn, rr)
  for (int64_t i = 0; i < n0; ++i) {
    r[i] = b[i] - q[i];
    bbn += (L->dinv[i] * b[i]) * (L->dinv[i] * b[i]);
    rr += (L->dinv[i] * r[i]) * (L->dinv[i] * r[i]);
  }
  const double tol2 = rtol * rtol * bbn;
  int it = 0;
  if (rr > tol2) {
    memcpy(L->b, r, sizeof(double) * n0);
    mg_vcycle(lv, nlevels, 0, nu, coarse_sweeps);
#pragma omp parallel for schedule(static) reduction(+ : rz)
    for (int64_t i = 0; i < n0; ++i) { p[i] = L->x[i]; rz += r[i] * L->x[i]; }
    while (it < maxit) {
      spmv(n0, L->rp, L->ci, L->va, p, q);
      double pq = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : pq)
      for (int64_t i = 0; i < n0; ++i) pq += p[i] * q[i];
      const double alpha = rz / pq;
      rr = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : rr)
      for (int64_t i = 0; i < n0; ++i) {
        x[i] += alpha * p[i];
        r[i] -= alpha * q[i];
        rr += (L->dinv[i] * r[i]) * (L->dinv[i] * r[i]);
      }
      ++it;
      if (!(rr > tol2)) break;
      memcpy(L->b, r, sizeof(double) * n0);
      mg_vcycle(lv, nlevels, 0, nu, coarse_sweeps);
      double rzn = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : rzn)
      for (int64_t i = 0; i < n0; ++i) rzn += r[i] * L->x[i];
      const double beta = rzn / rz;
      rz = rzn;
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < n0; ++i) p[i] = L->x[i] + beta * p[i];
    }
  }
  if (relres) *relres = bbn > 0.0 ? sqrt(rr / bbn) : 0.0;
  free(r); free(p); free(q);
  mg_free(lv, nlevels);
  return it;
}
