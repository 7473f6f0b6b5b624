// K7: block-CSR SpMV and the Jacobi-preconditioned CG / BiCGStab kernel chains.
//
// SpMV: persistent CTAs stream the matrix through shared memory with 1-D TMA bulk copies
// (cp.async.bulk + mbarrier, double buffered).  The non-zero stream is cut into tiles of ~tile_nnz
// blocks snapped to row boundaries, so every CTA moves the same number of bytes; one thread owns one
// scalar row of the tile, reads its values/columns from shared memory and gathers x through L1/L2.
// Dot products needed by the Krylov recurrences are fused into the producing kernel; per-CTA partials
// are combined by the last CTA to finish in a fixed order, so every reduction is bitwise reproducible.
// All Krylov scalars live on the device: the host only polls a `done` flag one batch behind.
#include "fsb_internal.cuh"
#include <algorithm>
#include <cmath>

#include "fsb_device.cuh"
#include "fsb_cgp.cuh"

// ------------------------------------------------------------------------------------ vector kernels
static constexpr int kVecThreads = 256;
static unsigned vec_grid(fsb_ctx* ctx, int64_t n) { return fsb_grid(n, kVecThreads * 4, std::min<int64_t>(kMaxPartials, (int64_t)ctx->sm_count * 8)); }
// the two classic-CG vector kernels keep 3 CTAs per SM resident (80 registers: ten 16-byte loads in flight per thread) and
// stride over the rows, so exactly one resident wave is launched: no partial last wave
static unsigned vec_grid_wave(fsb_ctx* ctx, int64_t n) { return fsb_grid(n, kVecThreads * 4, std::min<int64_t>(kMaxPartials, (int64_t)ctx->sm_count * 3)); }

__global__ void __launch_bounds__(kVecThreads)
k_dot(const double* __restrict__ x, const double* __restrict__ y, int64_t n0, int64_t n1, double* partials, double* out, unsigned* counter) {
  __shared__ double red[32];
  double s = 0.0;
  for (int64_t i = n0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n1; i += (int64_t)gridDim.x * blockDim.x) s += x[i] * y[i];
  double mine[1] = {block_sum(s, red)};
  finish_partials<1>(mine, partials, kMaxPartials, out, counter, red);
}

template <int BS>
__global__ void k_extract_dinv(int64_t n0, int64_t n1, const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx,
                               const double* __restrict__ vals, double* __restrict__ dinv, int jacobi) {
  for (int64_t row = n0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; row < n1; row += (int64_t)gridDim.x * blockDim.x) {
    double d = 1.0;
    if (jacobi) {
      const int64_t R = row / BS;
      const int i = (int)(row % BS);
      const int64_t base = row_ptr[R];
      const int len = (int)(row_ptr[R + 1] - base);
      int p = row_find(col_idx + base, 0, len, (int32_t)R);
      d = (p < len && col_idx[base + p] == R) ? 1.0 / vals[(base + p) * BS * BS + i * BS + i] : 1.0;
    }
    dinv[row] = d;
  }
}

// scalar slots in ctx->d_scalars
enum { S_PQ0 = 0, S_PQ1 = 1, S_RZ0 = 2, S_RR0 = 3, S_BB = 4, S_RZ1 = 5, S_RR1 = 6, S_FINAL_RR = 7,
       // BiCGStab
       S_RHO0 = 8, S_RRB0 = 9, S_BBB = 10, S_RHO1 = 11, S_RRB1 = 12, S_RV = 13, S_TS = 14, S_TT = 15, S_RT = 16, S_RS = 17,
       S_SPARE = 18 };
// state ints in ctx->d_state: [0] done, [1] iterations, [2] outcome (1 converged, 0 maxit, -1 breakdown)

// r = b - q ; p = z = dinv r ; sums r.z, z.z, (dinv b).(dinv b) -> out[0..3)
// (convergence is tested on the preconditioned residual ||M^-1 r||, PETSc's KSP default norm)
__global__ void __launch_bounds__(kVecThreads)
k_cg_init(int64_t n0, int64_t n1, const double* __restrict__ b, const double* __restrict__ q, const double* __restrict__ dinv,
          double* __restrict__ r, double* __restrict__ p, double* partials, double* out, unsigned* counter) {
  __shared__ double red[32];
  double s0 = 0, s1 = 0, s2 = 0;
  for (int64_t i = n0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n1; i += (int64_t)gridDim.x * blockDim.x) {
    const double bi = b[i], ri = bi - q[i], zi = dinv[i] * ri;
    r[i] = ri; p[i] = zi;
    const double zb = dinv[i] * bi;
    s0 += ri * zi; s1 += zi * zi; s2 += zb * zb;
  }
  double mine[3] = {block_sum(s0, red), block_sum(s1, red), block_sum(s2, red)};
  finish_partials<3>(mine, partials, kMaxPartials, out, counter, red);
}

// `restart`: keep the iteration count of the run being restarted (BiCGStab after a breakdown)
__global__ void k_check0(const double* __restrict__ scal, int rr_slot, int bb_slot, double rtol, double atol, int maxit, int* state, double* final_rr,
                         int restart = 0) {
  const double rr = scal[rr_slot], bb = scal[bb_slot];
  const double tol2 = fmax(rtol * rtol * bb, atol * atol);
  if (!restart) state[1] = 0;
  *final_rr = rr;
  if (rr <= tol2) { state[0] = 1; state[2] = 1; }
  else if (!(rr == rr)) { state[0] = 1; state[2] = -1; }
  else if (maxit <= state[1]) { state[0] = 1; state[2] = 0; }
  else { state[0] = 0; state[2] = 0; }
}

// Distributed CG over peer memory: which mailbox entries iteration `it` (seq = base + it, parity par) uses:
//   r.z at its start   MAIL_RZ + par      seq          (seeded after the start-up all-reduce, then posted by k_cg_update)
//   p.q                MAIL_PQ + par      seq + 1      (posted by the SpMV)
//   r.z, z.z after it  MAIL_RZ + (par^1)  seq + 1      (posted by k_cg_update)
//   halo flags raised by k_cg_pupdate to seq + 1, awaited by the next SpMV
struct CgPeer {
  PeerComm pc;              // pc.nranks <= 1: single GPU (or NCCL path): scalars come from ctx->d_scalars
  unsigned long long seq;
};

__global__ void k_mail_seed(PeerComm pc, int slot, unsigned long long seq, const double* __restrict__ scal, int s0, int s1) {
  if (threadIdx.x < pc.nranks) {      // the all-reduced start values enter the local mailbox as rank 0's contribution
    MailEntry* e = &pc.buf[pc.rank]->mail[slot][threadIdx.x];
    e->v[0] = threadIdx.x == 0 ? scal[s0] : 0.0;
    e->v[1] = threadIdx.x == 0 ? scal[s1] : 0.0;
    e->v[2] = 0.0;
    __threadfence();
    e->seq = seq;
  }
}

// alpha = rz/pq ; x += alpha p ; r -= alpha q ; z = dinv r ; sums rz', rr' -> out[0..2) (or the mailboxes)
template <bool PEER>
__global__ void __launch_bounds__(kVecThreads, 3)
k_cg_update(int64_t n0, int64_t n1, const double* __restrict__ scal, int rz_slot, int pq_slot, const double* __restrict__ p,
            const double* __restrict__ q, const double* __restrict__ dinv, double* __restrict__ x, double* __restrict__ r,
            double* partials, double* out, unsigned* counter, const int* done, CgPeer cp, int par) {
  __shared__ double red[32];
  if (*done) return;
  double alpha;
  if (PEER) {
    double a[2], b[2];
    mail_sum<2>(cp.pc, MAIL_RZ + par, cp.seq, a, red);
    mail_sum<2>(cp.pc, MAIL_PQ + par, cp.seq + 1, b, red);
    alpha = a[0] / b[0];
  } else {
    alpha = scal[rz_slot] / scal[pq_slot];
  }
  double s0 = 0, s1 = 0;
  auto one = [&](int64_t i) {
    x[i] += alpha * p[i];
    const double ri = r[i] - alpha * q[i];
    r[i] = ri;
    const double zi = dinv[i] * ri;
    s0 += ri * zi;
    s1 += zi * zi;
  };
  // 16-byte accesses over a body that starts on a 512-byte boundary (one warp request): a distributed rank's owned range starts
  // one vertex plane (an odd number of doubles) into its vectors, and a body that is only 16-byte aligned makes every
  // quarter-warp request straddle two lines (measured: the update phase 25 % slower on such a rank, profiles/cg_offset_ab_r2.txt)
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  const int64_t a_up = (n0 + 63) & ~(int64_t)63;      // 512-byte aligned body
  const int64_t a0 = a_up < n1 ? a_up : n1, npair = (n1 - a0) >> 1;
  if (tid < a0 - n0) one(n0 + tid);
  if (tid == 64 && a0 + 2 * npair < n1) one(n1 - 1);
  // two 16-byte pairs per thread and trip, all ten loads issued before the first use (bytes in flight, not occupancy,
  // is what a 7-stream kernel needs to reach the HBM rate)
  for (int64_t j = tid; j < npair; j += 2 * nth) {
    const int64_t i = a0 + 2 * j, k = i + 2 * nth;
    const bool two = j + nth < npair;
    const double2 zero = make_double2(0.0, 0.0);
    const double2 pv = *reinterpret_cast<const double2*>(p + i), qv = *reinterpret_cast<const double2*>(q + i);
    const double2 dv = *reinterpret_cast<const double2*>(dinv + i);
    double2 xv = *reinterpret_cast<double2*>(x + i), rv = *reinterpret_cast<double2*>(r + i);
    const double2 pw = two ? *reinterpret_cast<const double2*>(p + k) : zero, qw = two ? *reinterpret_cast<const double2*>(q + k) : zero;
    const double2 dw = two ? *reinterpret_cast<const double2*>(dinv + k) : zero;
    double2 xw = two ? *reinterpret_cast<double2*>(x + k) : zero, rw = two ? *reinterpret_cast<double2*>(r + k) : zero;
    xv.x += alpha * pv.x; xv.y += alpha * pv.y;
    rv.x -= alpha * qv.x; rv.y -= alpha * qv.y;
    *reinterpret_cast<double2*>(x + i) = xv;
    *reinterpret_cast<double2*>(r + i) = rv;
    const double z0 = dv.x * rv.x, z1 = dv.y * rv.y;
    s0 += rv.x * z0 + rv.y * z1;
    s1 += z0 * z0 + z1 * z1;
    if (two) {
      xw.x += alpha * pw.x; xw.y += alpha * pw.y;
      rw.x -= alpha * qw.x; rw.y -= alpha * qw.y;
      *reinterpret_cast<double2*>(x + k) = xw;
      *reinterpret_cast<double2*>(r + k) = rw;
      const double y0 = dw.x * rw.x, y1 = dw.y * rw.y;
      s0 += rw.x * y0 + rw.y * y1;
      s1 += y0 * y0 + y1 * y1;
    }
  }
  double mine[2] = {block_sum(s0, red), block_sum(s1, red)};
  if (PEER) finish_partials_mail<2>(mine, partials, kMaxPartials, counter, red, cp.pc, MAIL_RZ + (par ^ 1), cp.seq + 1);
  else finish_partials<2>(mine, partials, kMaxPartials, out, counter, red);
}

// beta = rz'/rz ; p = dinv r + beta p ; block 0 / thread 0 advances the iteration state.  Distributed: the
// boundary planes of the new p are also stored into the neighbours' ghost planes (peer memory) and the
// last CTA raises the neighbours' halo flags.
template <bool PEER>
__global__ void __launch_bounds__(kVecThreads, 3)
k_cg_pupdate(int64_t n0, int64_t n1, double* __restrict__ scal, int rz_old, int rz_new, int rr_new, int pq_slot,
             const double* __restrict__ r, const double* __restrict__ dinv, double* __restrict__ p, double rtol, double atol,
             int maxit, int* state, CgPeer cp, int par, unsigned* counter) {
  __shared__ double red[8];
  if (state[0]) return;
  constexpr bool peer = PEER;
  double rzn, rzo, rr_v, pq_v;
  if (peer) {
    double a[2], b[2], c[2];
    mail_sum<2>(cp.pc, MAIL_RZ + (par ^ 1), cp.seq + 1, a, red);
    mail_sum<2>(cp.pc, MAIL_RZ + par, cp.seq, b, red);
    mail_sum<2>(cp.pc, MAIL_PQ + par, cp.seq + 1, c, red);
    rzn = a[0]; rr_v = a[1]; rzo = b[0]; pq_v = c[0];
  } else {
    rzn = scal[rz_new]; rzo = scal[rz_old]; rr_v = scal[rr_new]; pq_v = scal[pq_slot];
  }
  const double beta = rzn / rzo;
  const int64_t plane = cp.pc.plane;
  auto push = [&](int64_t i, double v) {       // owned boundary planes -> neighbours' ghost planes
    if (cp.pc.lo_dst && i < n0 + plane) cp.pc.lo_dst[i - n0] = v;
    if (cp.pc.hi_dst && i >= n1 - plane) cp.pc.hi_dst[i - (n1 - plane)] = v;
  };
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  const int64_t a_up = (n0 + 63) & ~(int64_t)63;
  const int64_t a0 = a_up < n1 ? a_up : n1, npair = (n1 - a0) >> 1;
  if (tid < a0 - n0) { const int64_t i = n0 + tid; const double v = dinv[i] * r[i] + beta * p[i]; p[i] = v; if (peer) push(i, v); }
  if (tid == 64 && a0 + 2 * npair < n1) { const double v = dinv[n1 - 1] * r[n1 - 1] + beta * p[n1 - 1]; p[n1 - 1] = v; if (peer) push(n1 - 1, v); }
  for (int64_t j = tid; j < npair; j += 2 * nth) {
    const int64_t i = a0 + 2 * j, k = i + 2 * nth;
    const bool two = j + nth < npair;
    const double2 zero = make_double2(0.0, 0.0);
    const double2 dv = *reinterpret_cast<const double2*>(dinv + i), rv = *reinterpret_cast<const double2*>(r + i);
    double2 pv = *reinterpret_cast<double2*>(p + i);
    const double2 dw = two ? *reinterpret_cast<const double2*>(dinv + k) : zero, rw = two ? *reinterpret_cast<const double2*>(r + k) : zero;
    double2 pw = two ? *reinterpret_cast<double2*>(p + k) : zero;
    pv.x = dv.x * rv.x + beta * pv.x;
    pv.y = dv.y * rv.y + beta * pv.y;
    *reinterpret_cast<double2*>(p + i) = pv;
    if (peer) { push(i, pv.x); push(i + 1, pv.y); }
    if (two) {
      pw.x = dw.x * rw.x + beta * pw.x;
      pw.y = dw.y * rw.y + beta * pw.y;
      *reinterpret_cast<double2*>(p + k) = pw;
      if (peer) { push(k, pw.x); push(k + 1, pw.y); }
    }
  }
  if (peer) {
    __threadfence_system();           // this CTA's peer stores are visible system-wide before it counts itself done
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned t = atomicAdd(counter, 1u);
      if (t == gridDim.x - 1) {
        *counter = 0;
        __threadfence_system();
        if (cp.pc.rank > 0) st_release_sys(&cp.pc.buf[cp.pc.rank - 1]->halo_flag[1], cp.seq + 1);
        if (cp.pc.rank < cp.pc.nranks - 1) st_release_sys(&cp.pc.buf[cp.pc.rank + 1]->halo_flag[0], cp.seq + 1);
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const double rr = rr_v, bb = scal[S_BB], pq = pq_v;
    const double tol2 = fmax(rtol * rtol * bb, atol * atol);
    const int it = state[1] + 1;
    state[1] = it;
    scal[S_FINAL_RR] = rr;
    if (rr <= tol2) { state[2] = 1; __threadfence(); state[0] = 1; }
    else if (!(rr == rr) || !(pq == pq) || pq == 0.0 || rzo == 0.0) { state[2] = -1; __threadfence(); state[0] = 1; }
    else if (it >= maxit) { state[2] = 0; __threadfence(); state[0] = 1; }
  }
}

// ------------------------------------------------------------------------------------ Krylov drivers
// Work vectors live on the matrix and are reused by later solves (cudaMalloc/cudaFree synchronise the
// device; a transient run solves every step).  Every vector is zeroed at hand-out: ghost entries of a
// distributed solve are read by the SpMV gathers before the first halo exchange touches them.
static constexpr int kMaxSkew = 64 * 1024;     // bytes; 8 work vectors x the largest accepted vec_skew
struct Workspace {
  fsb_mat* A;
  int next = 0;
  int alloc(double** p, int64_t n) {
    fsb_ctx* ctx = A->ctx;
    if (next >= 8) FSB_FAIL(ctx, FSB_ERR_STATE, "Krylov workspace exhausted");
    // vec_skew: work vector k starts k * skew bytes into its block, so that the streams of a fused vector kernel (which all
    // advance at the same index) do not sit at the same offset of equally sized, equally aligned allocations
    const size_t skew = (size_t)ctx->vec_skew * (size_t)next;
    if (next >= A->work_count) {
      int rc = fsb_dmalloc(ctx, &A->work[next], (size_t)n, 512 + 8 * (size_t)kMaxSkew);
      if (rc) return rc;
      A->work_count = next + 1;
    }
    *p = A->work[next++] + skew / sizeof(double);
    FSB_CHECK_CUDA(ctx, cudaMemsetAsync(*p, 0, sizeof(double) * n, ctx->stream));
    return FSB_OK;
  }
};

struct SpmvTimer {
  // event pairs around the SpMV launches (profile mode); two halves so that one batch can be in
  // flight while the previous one is being collected
  std::vector<cudaEvent_t> ev[2];
  int used[2] = {0, 0};
  double total_ms = 0.0;
  void ensure(int pairs) {
    for (int h = 0; h < 2; ++h)
      while ((int)ev[h].size() < 2 * pairs) { cudaEvent_t e; cudaEventCreate(&e); ev[h].push_back(e); }
  }
  cudaEvent_t next(int h) { return ev[h][used[h]++]; }
  void collect(int h) {
    for (int i = 0; i + 1 < used[h]; i += 2) { float ms = 0; if (cudaEventElapsedTime(&ms, ev[h][i], ev[h][i + 1]) == cudaSuccess) total_ms += ms; }
    used[h] = 0;
  }
  ~SpmvTimer() { for (int h = 0; h < 2; ++h) for (auto e : ev[h]) cudaEventDestroy(e); }
};

// Iterations launched per host poll.  A transient run solves a similar system every step: the previous
// solve's iteration count sizes the first batch, so a 20-iteration solve does not pay for 32 launches.
static void batch_plan(const fsb_ctx* ctx, int last_iters, int* first, int* rest) {
  const int full = std::max(2, ctx->check_every & ~1);
  *first = *rest = full;
  if (last_iters > 0 && last_iters < full) {
    *first = std::min(full, (last_iters + 2) & ~1);
    *rest = std::max(2, (full / 4) & ~1);
  }
}

static int read_outcome(fsb_ctx* ctx, int rr_slot_final, int bb_slot, fsb_solve_info* info) {
  int st[4];
  double sc[32];
  FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(st, ctx->d_state, sizeof(int) * 4, cudaMemcpyDeviceToHost, ctx->stream));
  FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(sc, ctx->d_scalars, sizeof(double) * 32, cudaMemcpyDeviceToHost, ctx->stream));
  FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  info->iterations = st[1];
  info->converged = st[2];
  info->rnorm = std::sqrt(std::max(0.0, sc[rr_slot_final]));
  info->bnorm = std::sqrt(std::max(0.0, sc[bb_slot]));
  return FSB_OK;
}

// ------------------------------------------------------------------------------------ single-reduction CG
// Chronopoulos-Gear form of Jacobi-PCG: with u = M^-1 r and w = A u,
//   beta = g/g_prev ; alpha = g / (d - beta g / alpha_prev)            g = r.u, d = w.u
//   p = u + beta p ; s = w + beta s (= A p) ; x += alpha p ; r -= alpha s ; u = M^-1 r
// Both inner products of an iteration are available after ONE SpMV, so an iteration is two kernels (update,
// SpMV) with one global synchronisation point instead of three kernels with two.  On one GPU that trades 8 more
// bytes per row for one kernel boundary (a wash), so it is the distributed default: each rank waits for its peers
// once per iteration, and the reduction of (r.u, u.u) posted by the update kernel travels while the SpMV runs.
// Same iterates as classic PCG in exact arithmetic; convergence is tested on the same norm ||M^-1 r||.
enum { C1_G0 = 20, C1_Z0 = 21, C1_D0 = 22, C1_G1 = 24, C1_Z1 = 25, C1_D1 = 26, C1_A0 = 28, C1_A1 = 29 };

// start values -> parity-0 slots (and, peer path, the local mailboxes as rank 0's contribution)
__global__ void k_cg1_seed(double* __restrict__ scal, PeerComm pc, unsigned long long seq) {
  if (threadIdx.x == 0) { scal[C1_G0] = scal[S_RZ0]; scal[C1_Z0] = scal[S_RR0]; }
  if (pc.buf[pc.rank] != nullptr && (int)threadIdx.x < pc.nranks) {
    MailEntry* e = &pc.buf[pc.rank]->mail[MAIL_RZ][threadIdx.x];
    e->v[0] = threadIdx.x == 0 ? scal[S_RZ0] : 0.0;
    e->v[1] = threadIdx.x == 0 ? scal[S_RR0] : 0.0;
    e->v[2] = 0.0;
    MailEntry* f = &pc.buf[pc.rank]->mail[MAIL_PQ][threadIdx.x];
    f->v[0] = threadIdx.x == 0 ? scal[C1_D0] : 0.0;
    f->v[1] = 0.0; f->v[2] = 0.0;
    __threadfence();
    e->seq = seq;
    f->seq = seq;
  }
}

// iteration `it` (parity par): test convergence of the current iterate, then the five vector updates in one pass;
// sums r.u, u.u of the new residual -> slots of parity par^1 (or the mailboxes); peer path: the boundary planes
// of the new u go straight into the neighbours' ghost planes and the last CTA raises their halo flags
__global__ void __launch_bounds__(kVecThreads)
k_cg1_update(int64_t n0, int64_t n1, double* __restrict__ scal, int par, int it, int maxit, double rtol, double atol,
             const double* __restrict__ dinv, double* __restrict__ u, const double* __restrict__ w, double* __restrict__ p,
             double* __restrict__ s, double* __restrict__ x, double* __restrict__ r, double* partials, unsigned* counter,
             unsigned* halo_counter, int* state, CgPeer cp) {
  __shared__ double red[32];
  if (state[0]) return;
  const bool peer = cp.pc.nranks > 1;
  const int G = par ? C1_G1 : C1_G0, Gp = par ? C1_G0 : C1_G1, Ap = par ? C1_A0 : C1_A1, Ac = par ? C1_A1 : C1_A0;
  double g, zz, d;
  if (peer) {
    double a[2], b[1];
    mail_sum_pair<2, 1>(cp.pc, MAIL_RZ + par, cp.seq, a, MAIL_PQ + par, cp.seq, b);
    g = a[0]; zz = a[1]; d = b[0];
  } else {
    g = scal[G]; zz = scal[G + 1]; d = scal[G + 2];
  }
  const double bb = scal[S_BB];
  const double tol2 = fmax(rtol * rtol * bb, atol * atol);
  const double g_prev = it > 0 ? scal[Gp] : 1.0, a_prev = it > 0 ? scal[Ap] : 1.0;
  const double beta = it > 0 ? g / g_prev : 0.0;
  const double denom = it > 0 ? d - beta * g / a_prev : d;
  const double alpha = g / denom;
  const bool conv = zz <= tol2;
  const bool broken = !(zz == zz) || !(alpha == alpha) || denom == 0.0 || (it > 0 && g_prev == 0.0);
  if (conv || broken || it >= maxit) {          // the same decision in every CTA (and on every rank)
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      state[1] = it;
      state[2] = conv ? 1 : (broken ? -1 : 0);
      scal[S_FINAL_RR] = zz;
      __threadfence();
      state[0] = 1;
    }
    return;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    scal[Ac] = alpha;
    if (peer) scal[G] = g;                      // next iteration's g_prev (single GPU / NCCL: already there)
    state[1] = it + 1;
    scal[S_FINAL_RR] = zz;
  }
  const int64_t plane = cp.pc.plane;
  double s0 = 0, s1 = 0;
  auto push = [&](int64_t i, double v) {       // owned boundary planes -> neighbours' ghost planes
    if (cp.pc.lo_dst && i < n0 + plane) cp.pc.lo_dst[i - n0] = v;
    if (cp.pc.hi_dst && i >= n1 - plane) cp.pc.hi_dst[i - (n1 - plane)] = v;
  };
  // p, s, x, r, w and dinv are touched once per iteration: streaming (evict-first) accesses, so that the new u,
  // which the SpMV gathers next, is what stays in L2
  auto one = [&](int64_t i) {
    const double pi = u[i] + beta * __ldcs(p + i);
    const double si = __ldcs(w + i) + beta * __ldcs(s + i);
    __stcs(p + i, pi); __stcs(s + i, si);
    __stcs(x + i, __ldcs(x + i) + alpha * pi);
    const double ri = __ldcs(r + i) - alpha * si;
    __stcs(r + i, ri);
    const double ui = __ldcs(dinv + i) * ri;
    u[i] = ui;
    s0 += ri * ui; s1 += ui * ui;
    if (peer) push(i, ui);
  };
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  const int64_t a_up = (n0 + 63) & ~(int64_t)63;      // 512-byte aligned body
  const int64_t a0 = a_up < n1 ? a_up : n1, npair = (n1 - a0) >> 1;
  if (tid < a0 - n0) one(n0 + tid);
  if (tid == 64 && a0 + 2 * npair < n1) one(n1 - 1);
  for (int64_t j = tid; j < npair; j += nth) {
    const int64_t i = a0 + 2 * j;
    const double2 uv = *reinterpret_cast<const double2*>(u + i);
    const double2 wv = __ldcs(reinterpret_cast<const double2*>(w + i));
    double2 pv = __ldcs(reinterpret_cast<const double2*>(p + i)), sv = __ldcs(reinterpret_cast<const double2*>(s + i));
    double2 xv = __ldcs(reinterpret_cast<const double2*>(x + i)), rv = __ldcs(reinterpret_cast<const double2*>(r + i));
    const double2 dv = __ldcs(reinterpret_cast<const double2*>(dinv + i));
    pv.x = uv.x + beta * pv.x; pv.y = uv.y + beta * pv.y;
    sv.x = wv.x + beta * sv.x; sv.y = wv.y + beta * sv.y;
    xv.x += alpha * pv.x; xv.y += alpha * pv.y;
    rv.x -= alpha * sv.x; rv.y -= alpha * sv.y;
    double2 un;
    un.x = dv.x * rv.x; un.y = dv.y * rv.y;
    __stcs(reinterpret_cast<double2*>(p + i), pv);
    __stcs(reinterpret_cast<double2*>(s + i), sv);
    __stcs(reinterpret_cast<double2*>(x + i), xv);
    __stcs(reinterpret_cast<double2*>(r + i), rv);
    *reinterpret_cast<double2*>(u + i) = un;
    s0 += rv.x * un.x + rv.y * un.y;
    s1 += un.x * un.x + un.y * un.y;
    if (peer) { push(i, un.x); push(i + 1, un.y); }
  }
  if (peer) {
    __threadfence_system();           // this CTA's peer stores are visible system-wide before it counts itself done
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned t = atomicAdd(halo_counter, 1u);
      if (t == gridDim.x - 1) {
        *halo_counter = 0;
        __threadfence_system();
        if (cp.pc.rank > 0) st_release_sys(&cp.pc.buf[cp.pc.rank - 1]->halo_flag[1], cp.seq + 1);
        if (cp.pc.rank < cp.pc.nranks - 1) st_release_sys(&cp.pc.buf[cp.pc.rank + 1]->halo_flag[0], cp.seq + 1);
      }
    }
  }
  double mine[2] = {block_sum(s0, red), block_sum(s1, red)};
  if (peer) finish_partials_mail<2>(mine, partials, kMaxPartials, counter, red, cp.pc, MAIL_RZ + (par ^ 1), cp.seq + 1);
  else finish_partials<2>(mine, partials, kMaxPartials, scal + (par ? C1_G0 : C1_G1), counter, red);
}

// can this solve run as one persistent kernel?  One GPU, or z-slabs with the peer mailboxes mapped (the same answer on every rank)
// Default (cg_variant 0): when distributed — there an iteration is latency-bound (98 us of HBM work per rank at 256^3 on
// 8 GPUs) and the kernel boundaries of the chains are what it costs.  On one GPU the classic chain moves 8 bytes per row
// less and is as fast or faster (profiles/cg_ab_r2.txt: 0.843 vs 0.898 ms at 256^3, 0.118 vs 0.117 ms on one slab of it);
// cg_variant 3 forces the persistent kernel there.
static bool cg_persist_applies(fsb_ctx* ctx, fsb_mat* S) {
  if (ctx->cg_variant != 0 && ctx->cg_variant != 3) return false;
  if (!fsb_cgp_supported(S)) return false;
  if (!fsb_dist_active(ctx)) return ctx->cg_variant == 3;
  return fsb_dist_p2p_ready(ctx) && fsb_spmv_supports_p2p(S);
}

static int solve_cg1(fsb_mat* A, fsb_mat* S, fsb_vec* b, fsb_vec* x, double rtol, double atol, int32_t maxit, int32_t precond,
                     fsb_solve_info* info, cudaEvent_t e0, cudaEvent_t e1) {
  fsb_ctx* ctx = A->ctx;
  const int64_t n = A->nbrows * A->bs;
  const int64_t n0 = A->own0 * A->bs, n1 = A->own1 * A->bs;
  const bool dist = fsb_dist_active(ctx);
  const bool p2p = dist && fsb_dist_p2p_ready(ctx) && fsb_spmv_supports_p2p(S);
  const bool persist = cg_persist_applies(ctx, S);
  Workspace ws{A};
  double *r, *u, *w, *p, *s, *dinv;
  int rc;
  if ((rc = ws.alloc(&r, n)) || (rc = ws.alloc(&w, n)) || (rc = ws.alloc(&dinv, n)) || (rc = ws.alloc(&p, n)) || (rc = ws.alloc(&s, n))) return rc;
  CgPeer cp;
  memset(&cp, 0, sizeof(cp));
  cp.pc.nranks = 1;
  unsigned long long seq_base = 0;
  if (p2p) {
    if ((rc = fsb_dist_share_p(A, n))) return rc;          // collective; here the shared vector is u, the SpMV input
    u = A->p_dist;
    FSB_CHECK_CUDA(ctx, cudaMemsetAsync(u, 0, sizeof(double) * n, ctx->stream));
    if ((rc = fsb_dist_peer_comm(A, &cp.pc))) return rc;
    seq_base = fsb_dist_seq_reserve(ctx, 0);
  } else if ((rc = ws.alloc(&u, n))) {
    return rc;
  }
  if (persist && !dist) {        // single GPU: the kernel's mailbox is a local CommBuf
    if (!ctx->cg_comm) {
      FSB_CHECK_CUDA(ctx, cudaMalloc(&ctx->cg_comm, sizeof(CommBuf)));
      FSB_CHECK_CUDA(ctx, cudaMemsetAsync(ctx->cg_comm, 0, sizeof(CommBuf), ctx->stream));
    }
    cp.pc.rank = 0; cp.pc.nranks = 1;
    cp.pc.buf[0] = (CommBuf*)ctx->cg_comm;
    seq_base = ctx->cg_seq;
  }
  double* scal = ctx->d_scalars;
  int* state = ctx->d_state;
  const unsigned vg = vec_grid(ctx, n1 - n0);
  SpmvTimer timer;
  FSB_CHECK_CUDA(ctx, cudaMemsetAsync(state, 0, sizeof(int) * 8, ctx->stream));
  FSB_CHECK_CUDA(ctx, cudaMemsetAsync(ctx->d_counters, 0, sizeof(unsigned) * 16, ctx->stream));
#define DINV_LAUNCH(BS) k_extract_dinv<BS><<<fsb_grid(n1 - n0, 256, (int64_t)ctx->sm_count * 16), 256, 0, ctx->stream>>>(n0, n1, A->row_ptr, A->col_idx, A->vals, dinv, precond == 1)
  if (A->bs == 1) DINV_LAUNCH(1); else if (A->bs == 2) DINV_LAUNCH(2); else DINV_LAUNCH(3);
#undef DINV_LAUNCH
  FSB_LAUNCH_CHECK(ctx);
  // r0 = b - A x0 ; u0 = M^-1 r0 ; w0 = A u0
  if (dist && (rc = fsb_dist_halo_raw(ctx, x->d, n))) return rc;
  if ((rc = fsb_launch_spmv(S, x->d, w, nullptr, 0, nullptr, nullptr))) return rc;
  k_cg_init<<<vg, kVecThreads, 0, ctx->stream>>>(n0, n1, b->d, w, dinv, r, u, ctx->d_partials, scal + S_RZ0, ctx->d_counters + 1);
  FSB_LAUNCH_CHECK(ctx);
  if (dist && (rc = fsb_dist_allreduce_sum_dev(ctx, scal + S_RZ0, 3))) return rc;
  k_check0<<<1, 1, 0, ctx->stream>>>(scal, S_RR0, S_BB, rtol, atol, maxit, state, scal + S_FINAL_RR);
  FSB_LAUNCH_CHECK(ctx);
  if (dist && (rc = fsb_dist_halo_raw(ctx, u, n))) return rc;
  if ((rc = fsb_launch_spmv(S, u, w, u, 0, scal + C1_D0, state))) return rc;
  if (dist && (rc = fsb_dist_allreduce_sum_dev(ctx, scal + C1_D0, 1))) return rc;
  k_cg1_seed<<<1, 32, 0, ctx->stream>>>(scal, cp.pc, seq_base);
  FSB_LAUNCH_CHECK(ctx);

  if (persist) {
    // the whole iteration loop in one cooperative launch (fsb_cgp.cu); k_check0 has already decided a zero-iteration solve
    CgpVectors cv{dinv, u, w, p, s, x->d, r};
    double phase_ms[4] = {0, 0, 0, 0};
    if ((rc = fsb_cgp_run(A, S, cv, rtol, atol, maxit, S_BB, S_FINAL_RR, cp.pc, seq_base, phase_ms))) return rc;
    FSB_CHECK_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
    rc = read_outcome(ctx, S_FINAL_RR, S_BB, info);
    if (rc) return rc;
    A->last_iters = info->iterations;
    if (p2p) fsb_dist_seq_reserve(ctx, (unsigned long long)info->iterations + 2);   // identical on every rank
    else ctx->cg_seq += (unsigned long long)info->iterations + 2;
    info->spmv_ms = phase_ms[2];
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    info->solve_ms = ms;
    if (info->converged < 0) FSB_FAIL(ctx, FSB_ERR_BREAKDOWN, "CG breakdown (non-finite or zero recurrence scalar)");
    return FSB_OK;
  }

  int batch_first, batch_rest;
  batch_plan(ctx, A->last_iters, &batch_first, &batch_rest);
  cudaEvent_t polled[2];
  FSB_CHECK_CUDA(ctx, cudaEventCreateWithFlags(&polled[0], cudaEventDisableTiming));
  FSB_CHECK_CUDA(ctx, cudaEventCreateWithFlags(&polled[1], cudaEventDisableTiming));
  struct PollGuard { cudaEvent_t* e; ~PollGuard() { cudaEventDestroy(e[0]); cudaEventDestroy(e[1]); } } pguard{polled};
  if (ctx->profile) timer.ensure(std::max(batch_first, batch_rest));
  int launched = 0;
  for (int nb = 0;; ++nb) {
    const int slot = nb & 1;
    const int batch = nb == 0 ? batch_first : batch_rest;
    const bool more = launched <= maxit;          // iteration `maxit` is the launch that records the maxit outcome
    if (more) {
      for (int k = 0; k < batch; ++k) {
        const int it = launched + k, par = it & 1;
        cp.seq = seq_base + (unsigned long long)it;
        k_cg1_update<<<vg, kVecThreads, 0, ctx->stream>>>(n0, n1, scal, par, it, maxit, rtol, atol, dinv, u, w, p, s, x->d, r, ctx->d_partials,
                                                          ctx->d_counters + 2, ctx->d_counters + 4, state, cp);
        FSB_LAUNCH_CHECK(ctx);
        if (dist && !p2p && (rc = fsb_dist_halo_raw(ctx, u, n))) return rc;
        if (ctx->profile) cudaEventRecord(timer.next(slot), ctx->stream);
        if (p2p) {
          fsb_spmv_dist dd{cp.pc, cp.seq + 1, MAIL_PQ + (par ^ 1), cp.seq + 1};
          if ((rc = fsb_launch_spmv(S, u, w, u, 0, nullptr, state, &dd))) return rc;
        } else if ((rc = fsb_launch_spmv(S, u, w, u, 0, scal + (par ? C1_D0 : C1_D1), state))) {
          return rc;
        }
        if (ctx->profile) cudaEventRecord(timer.next(slot), ctx->stream);
        // one all-reduce per iteration: (r.u, u.u, w.u) sit next to each other
        if (dist && !p2p && (rc = fsb_dist_allreduce_sum_dev(ctx, scal + (par ? C1_G0 : C1_G1), 3))) return rc;
      }
      launched += batch;
    }
    FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(ctx->h_state + 8 * slot, state, sizeof(int) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    FSB_CHECK_CUDA(ctx, cudaEventRecord(polled[slot], ctx->stream));
    if (nb > 0) {
      FSB_CHECK_CUDA(ctx, cudaEventSynchronize(polled[slot ^ 1]));
      if (ctx->profile) timer.collect(slot ^ 1);
      if (ctx->h_state[8 * (slot ^ 1)]) break;
    }
    if (!more) {
      FSB_CHECK_CUDA(ctx, cudaEventSynchronize(polled[slot]));
      break;
    }
  }
  FSB_CHECK_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
  rc = read_outcome(ctx, S_FINAL_RR, S_BB, info);
  if (rc) return rc;
  A->last_iters = info->iterations;
  if (p2p) fsb_dist_seq_reserve(ctx, (unsigned long long)info->iterations + 2);   // identical on every rank
  if (ctx->profile) { timer.collect(0); timer.collect(1); info->spmv_ms = timer.total_ms; }
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  info->solve_ms = ms;
  if (info->converged < 0) FSB_FAIL(ctx, FSB_ERR_BREAKDOWN, "CG breakdown (non-finite or zero recurrence scalar)");
  return FSB_OK;
}

extern "C" int fsb_solve_cg(fsb_mat* A, fsb_vec* b, fsb_vec* x, double rtol, double atol, int32_t maxit, int32_t precond,
                            fsb_solve_info* info) {
  if (!A || !b || !x || !info) return FSB_ERR_ARG;
  fsb_ctx* ctx = A->ctx;
  const int64_t n = A->nbrows * A->bs;
  if (b->n != n || x->n != n) FSB_FAIL(ctx, FSB_ERR_ARG, "vector sizes do not match the matrix");
  memset(info, 0, sizeof(*info));
  const int64_t n0 = A->own0 * A->bs, n1 = A->own1 * A->bs;
  const bool dist = fsb_dist_active(ctx);
  cudaEvent_t e0, e1;
  FSB_CHECK_CUDA(ctx, cudaEventCreate(&e0));
  FSB_CHECK_CUDA(ctx, cudaEventCreate(&e1));
  struct EvGuard { cudaEvent_t a, b; ~EvGuard() { cudaEventDestroy(a); cudaEventDestroy(b); } } guard{e0, e1};
  FSB_CHECK_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
  int rc;
  fsb_mat* S = A;             // the SpMV operand: A itself, or its copy without the exactly-zero blocks
  if (ctx->drop_zeros && (rc = fsb_mat_squeeze(A, &S))) return rc;
  info->operand_nnzb = S->nnzb;
  // cg_variant: 0 = classic on one GPU, single-reduction when distributed; 1 = classic; 2 = single-reduction
  if (ctx->cg_variant == 2 || ctx->cg_variant == 3 || (ctx->cg_variant == 0 && (dist || cg_persist_applies(ctx, S))))
    return solve_cg1(A, S, b, x, rtol, atol, maxit, precond, info, e0, e1);
  // peer-memory path: mailboxes mapped on every rank, staged SpMV kernel; the decision is the same on every rank
  const bool p2p = dist && fsb_dist_p2p_ready(ctx) && fsb_spmv_supports_p2p(S);
  Workspace ws{A};
  double *r, *p, *q, *dinv;
  if ((rc = ws.alloc(&r, n)) || (rc = ws.alloc(&q, n)) || (rc = ws.alloc(&dinv, n))) return rc;
  CgPeer cp;
  memset(&cp, 0, sizeof(cp));
  cp.pc.nranks = 1;
  unsigned long long seq_base = 0;
  if (p2p) {
    if ((rc = fsb_dist_share_p(A, n))) return rc;          // collective; the neighbours write p's ghost planes directly
    p = A->p_dist;
    FSB_CHECK_CUDA(ctx, cudaMemsetAsync(p, 0, sizeof(double) * n, ctx->stream));
    if ((rc = fsb_dist_peer_comm(A, &cp.pc))) return rc;
    seq_base = fsb_dist_seq_reserve(ctx, 0);
  } else if ((rc = ws.alloc(&p, n))) {
    return rc;
  }
  double* scal = ctx->d_scalars;
  int* state = ctx->d_state;
  const unsigned vg = vec_grid(ctx, n1 - n0), vgw = vec_grid_wave(ctx, n1 - n0);
  SpmvTimer timer;

  FSB_CHECK_CUDA(ctx, cudaMemsetAsync(state, 0, sizeof(int) * 8, ctx->stream));
  // last-CTA counters: a kernel that observes `done` half-way may leave one partially counted
  FSB_CHECK_CUDA(ctx, cudaMemsetAsync(ctx->d_counters, 0, sizeof(unsigned) * 16, ctx->stream));
#define DINV_LAUNCH(BS) k_extract_dinv<BS><<<fsb_grid(n1 - n0, 256, (int64_t)ctx->sm_count * 16), 256, 0, ctx->stream>>>(n0, n1, A->row_ptr, A->col_idx, A->vals, dinv, precond == 1)
  if (A->bs == 1) DINV_LAUNCH(1); else if (A->bs == 2) DINV_LAUNCH(2); else DINV_LAUNCH(3);
#undef DINV_LAUNCH
  FSB_LAUNCH_CHECK(ctx);
  // r0 = b - A x0
  if (dist && (rc = fsb_dist_halo_raw(ctx, x->d, n))) return rc;
  if ((rc = fsb_launch_spmv(S, x->d, q, nullptr, 0, nullptr, nullptr))) return rc;
  k_cg_init<<<vg, kVecThreads, 0, ctx->stream>>>(n0, n1, b->d, q, dinv, r, p, ctx->d_partials, scal + S_RZ0, ctx->d_counters + 1);
  FSB_LAUNCH_CHECK(ctx);
  if (dist && (rc = fsb_dist_allreduce_sum_dev(ctx, scal + S_RZ0, 3))) return rc;
  k_check0<<<1, 1, 0, ctx->stream>>>(scal, S_RR0, S_BB, rtol, atol, maxit, state, scal + S_FINAL_RR);
  FSB_LAUNCH_CHECK(ctx);
  if (p2p) {
    k_mail_seed<<<1, 32, 0, ctx->stream>>>(cp.pc, MAIL_RZ + 0, seq_base, scal, S_RZ0, S_RR0);
    FSB_LAUNCH_CHECK(ctx);
  }

  // iteration batches; the host polls the state one batch behind the launches
  int batch_first, batch_rest;
  batch_plan(ctx, A->last_iters, &batch_first, &batch_rest);
  cudaEvent_t polled[2];
  FSB_CHECK_CUDA(ctx, cudaEventCreateWithFlags(&polled[0], cudaEventDisableTiming));
  FSB_CHECK_CUDA(ctx, cudaEventCreateWithFlags(&polled[1], cudaEventDisableTiming));
  struct PollGuard { cudaEvent_t* e; ~PollGuard() { cudaEventDestroy(e[0]); cudaEventDestroy(e[1]); } } pguard{polled};
  if (ctx->profile) timer.ensure(std::max(batch_first, batch_rest));
  int launched = 0;
  for (int nb = 0;; ++nb) {
    const int slot = nb & 1;
    const int batch = nb == 0 ? batch_first : batch_rest;
    const bool more = launched < maxit;
    if (more) {
      for (int k = 0; k < batch; ++k) {
        const int it = launched + k, par = it & 1;
        const int pq = par ? S_PQ1 : S_PQ0, rz = par ? S_RZ1 : S_RZ0, rzn = par ? S_RZ0 : S_RZ1, rrn = par ? S_RR0 : S_RR1;
        cp.seq = seq_base + (unsigned long long)it;
        // ghost planes of p: NCCL send/recv, or (peer path) already written by the neighbours' k_cg_pupdate
        if (dist && (!p2p || it == 0) && (rc = fsb_dist_halo_raw(ctx, p, n))) return rc;
        if (ctx->profile) cudaEventRecord(timer.next(slot), ctx->stream);
        if (p2p) {
          fsb_spmv_dist dd{cp.pc, it > 0 ? cp.seq : 0ull, MAIL_PQ + par, cp.seq + 1};
          if ((rc = fsb_launch_spmv(S, p, q, p, 0, nullptr, state, &dd))) return rc;
        } else if ((rc = fsb_launch_spmv(S, p, q, p, 0, scal + pq, state))) {
          return rc;
        }
        if (ctx->profile) cudaEventRecord(timer.next(slot), ctx->stream);
        if (dist && !p2p && (rc = fsb_dist_allreduce_sum_dev(ctx, scal + pq, 1))) return rc;
        if (p2p) k_cg_update<true><<<vgw, kVecThreads, 0, ctx->stream>>>(n0, n1, scal, rz, pq, p, q, dinv, x->d, r, ctx->d_partials, scal + rzn, ctx->d_counters + 2, state, cp, par);
        else k_cg_update<false><<<vgw, kVecThreads, 0, ctx->stream>>>(n0, n1, scal, rz, pq, p, q, dinv, x->d, r, ctx->d_partials, scal + rzn, ctx->d_counters + 2, state, cp, par);
        FSB_LAUNCH_CHECK(ctx);
        if (dist && !p2p && (rc = fsb_dist_allreduce_sum_dev(ctx, scal + rzn, 2))) return rc;
        if (p2p) k_cg_pupdate<true><<<vgw, kVecThreads, 0, ctx->stream>>>(n0, n1, scal, rz, rzn, rrn, pq, r, dinv, p, rtol, atol, maxit, state, cp, par, ctx->d_counters + 4);
        else k_cg_pupdate<false><<<vgw, kVecThreads, 0, ctx->stream>>>(n0, n1, scal, rz, rzn, rrn, pq, r, dinv, p, rtol, atol, maxit, state, cp, par, ctx->d_counters + 4);
        FSB_LAUNCH_CHECK(ctx);
      }
      launched += batch;
    }
    FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(ctx->h_state + 8 * slot, state, sizeof(int) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    FSB_CHECK_CUDA(ctx, cudaEventRecord(polled[slot], ctx->stream));
    if (nb > 0) {   // look at the snapshot taken after the previous batch while this one runs
      FSB_CHECK_CUDA(ctx, cudaEventSynchronize(polled[slot ^ 1]));
      if (ctx->profile) timer.collect(slot ^ 1);
      if (ctx->h_state[8 * (slot ^ 1)]) break;
    }
    if (!more) {    // nothing new was launched: the device has enforced maxit by now
      FSB_CHECK_CUDA(ctx, cudaEventSynchronize(polled[slot]));
      break;
    }
  }
  FSB_CHECK_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
  rc = read_outcome(ctx, S_FINAL_RR, S_BB, info);
  if (rc) return rc;
  A->last_iters = info->iterations;
  if (p2p) fsb_dist_seq_reserve(ctx, (unsigned long long)info->iterations + 2);   // identical on every rank
  if (ctx->profile) { timer.collect(0); timer.collect(1); info->spmv_ms = timer.total_ms; }
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  info->solve_ms = ms;
  if (info->converged < 0) FSB_FAIL(ctx, FSB_ERR_BREAKDOWN, "CG breakdown (non-finite or zero recurrence scalar)");
  return FSB_OK;
}

extern "C" int fsb_dot(fsb_vec* x, fsb_vec* y, double* result) {
  if (!x || !y || !result || x->n != y->n) return FSB_ERR_ARG;
  fsb_ctx* ctx = x->ctx;
  int64_t n0 = 0, n1 = x->n;
  fsb_dist_owned_range(ctx, x->n, &n0, &n1);
  k_dot<<<vec_grid(ctx, n1 - n0), kVecThreads, 0, ctx->stream>>>(x->d, y->d, n0, n1, ctx->d_partials, ctx->d_scalars + S_SPARE, ctx->d_counters + 3);
  FSB_LAUNCH_CHECK(ctx);
  if (fsb_dist_active(ctx)) {
    int rc = fsb_dist_allreduce_sum_dev(ctx, ctx->d_scalars + S_SPARE, 1);
    if (rc) return rc;
  }
  FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(result, ctx->d_scalars + S_SPARE, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FSB_OK;
}

// ------------------------------------------------------------------------------------ BiCGStab
// Right-Jacobi BiCGStab with two synchronisation-free fusions (4 kernels per iteration instead of 7):
//   v = A ph            (+ rhat.v)                                   SpMV
//   s = r - alpha v ; sh = dinv s                 (+ rhat.s)         k_bcg_s
//   t = A sh            (+ t.s, t.t, rhat.t)                         SpMV
//   omega = ts/tt ; rho' = rhat.s - omega rhat.t  (= rhat.r_new, so no reduction separates the two updates)
//   x += alpha ph + omega sh ; r = s - omega t ; p = r + beta (p - omega v) ; ph = dinv p   (+ |dinv r|^2)   k_bcg_xp
// The last CTA of k_bcg_xp holds the reduced norm and advances the iteration state itself (single GPU);
// distributed runs all-reduce first and use the one-thread k_bcg_check.

// r = b - q ; rhat = p = r ; ph = dinv r ; sums rho=rhat.r, |dinv r|^2, |dinv b|^2 -> out[0..3)
__global__ void __launch_bounds__(kVecThreads)
k_bcg_init(int64_t n0, int64_t n1, const double* __restrict__ b, const double* __restrict__ q, const double* __restrict__ dinv,
           double* __restrict__ r, double* __restrict__ rhat, double* __restrict__ p, double* __restrict__ ph,
           double* partials, double* out, unsigned* counter) {
  __shared__ double red[32];
  double s0 = 0, s1 = 0, s2 = 0;
  for (int64_t i = n0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n1; i += (int64_t)gridDim.x * blockDim.x) {
    const double bi = b[i], ri = bi - q[i];
    r[i] = ri; rhat[i] = ri; p[i] = ri;
    const double zi = dinv[i] * ri, zb = dinv[i] * bi;
    ph[i] = zi;
    s0 += ri * ri; s1 += zi * zi; s2 += zb * zb;
  }
  double mine[3] = {block_sum(s0, red), block_sum(s1, red), block_sum(s2, red)};
  finish_partials<3>(mine, partials, kMaxPartials, out, counter, red);
}

// alpha = rho/(rhat.v) ; s = r - alpha v (stored in r) ; sh = dinv s ; sum rhat.s -> out[0]
__global__ void __launch_bounds__(kVecThreads)
k_bcg_s(int64_t n0, int64_t n1, const double* __restrict__ scal, int rho_cur, const double* __restrict__ v,
        const double* __restrict__ rhat, const double* __restrict__ dinv, double* __restrict__ r, double* __restrict__ sh,
        double* partials, double* out, unsigned* counter, const int* done) {
  __shared__ double red[32];
  if (*done) return;
  const double alpha = scal[rho_cur] / scal[S_RV];
  double s0 = 0;
  for (int64_t i = n0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n1; i += (int64_t)gridDim.x * blockDim.x) {
    const double si = r[i] - alpha * v[i];
    r[i] = si;
    sh[i] = dinv[i] * si;
    s0 += rhat[i] * si;
  }
  double mine[1] = {block_sum(s0, red)};
  finish_partials<1>(mine, partials, kMaxPartials, out, counter, red);
}

// the recurrence scalars of one iteration; `ok` is false when any of them is zero or not finite, and then the
// iteration leaves x untouched (a breakdown is reported and the host restarts from x)
struct BcgScalars { double alpha, omega, rho_new, beta; bool ok; };
__device__ __forceinline__ BcgScalars bcg_scalars(const double* __restrict__ scal, int rho_cur) {
  BcgScalars c;
  const double rho = scal[rho_cur];
  c.alpha = rho / scal[S_RV];
  c.omega = scal[S_TS] / scal[S_TT];
  c.rho_new = scal[S_RS] - c.omega * scal[S_RT];
  c.beta = (c.rho_new / rho) * (c.alpha / c.omega);
  c.ok = isfinite(c.alpha) && isfinite(c.omega) && isfinite(c.beta) && isfinite(c.rho_new) && c.omega != 0.0;
  return c;
}

__device__ __forceinline__ void bcg_advance(double* __restrict__ scal, int rr_new, double rho_new, bool ok, double rtol, double atol, int maxit,
                                            int* state) {
  const double rr = scal[rr_new], bb = scal[S_BBB];
  const double tol2 = fmax(rtol * rtol * bb, atol * atol);
  const int it = state[1] + 1;
  state[1] = it;
  if (!ok) { state[2] = -1; state[0] = 1; return; }          // x was not updated: S_FINAL_RR keeps the last valid norm
  scal[S_FINAL_RR] = rr;
  if (rr <= tol2) { state[2] = 1; state[0] = 1; }
  else if (!(rr == rr) || rho_new == 0.0) { state[2] = -1; state[0] = 1; }
  else if (it >= maxit) { state[2] = 0; state[0] = 1; }
}

// the two updates of one iteration in one pass; sum |dinv r|^2 -> out[0]; rho' -> scal[rho_next]
__global__ void __launch_bounds__(kVecThreads)
k_bcg_xp(int64_t n0, int64_t n1, double* __restrict__ scal, int rho_cur, int rho_next, int rr_new, const double* __restrict__ sh,
         const double* __restrict__ t, const double* __restrict__ v, const double* __restrict__ dinv, double* __restrict__ x,
         double* __restrict__ r, double* __restrict__ p, double* __restrict__ ph, double* partials, unsigned* counter,
         double rtol, double atol, int maxit, int* state, int advance) {
  __shared__ double red[32];
  if (state[0]) return;
  const BcgScalars c = bcg_scalars(scal, rho_cur);
  const double alpha = c.alpha, omega = c.omega, rho_new = c.rho_new, beta = c.beta;
  if (blockIdx.x == 0 && threadIdx.x == 0) scal[rho_next] = rho_new;
  double s0 = 0;
  for (int64_t i = n0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c.ok && i < n1; i += (int64_t)gridDim.x * blockDim.x) {
    const double di = dinv[i], shi = sh[i];
    x[i] += alpha * ph[i] + omega * shi;
    const double ri = r[i] - omega * t[i];
    r[i] = ri;
    const double zi = di * ri;
    s0 += zi * zi;
    const double pi = ri + beta * (p[i] - omega * v[i]);
    p[i] = pi;
    ph[i] = di * pi;
  }
  double mine[1] = {block_sum(s0, red)};
  const bool last = finish_partials_last<1>(mine, partials, kMaxPartials, scal + rr_new, counter, red);
  if (advance && last && threadIdx.x == 0) bcg_advance(scal, rr_new, rho_new, c.ok, rtol, atol, maxit, state);
}

__global__ void k_bcg_check(double* __restrict__ scal, int rho_cur, int rho_new, int rr_new, double rtol, double atol, int maxit, int* state) {
  if (state[0]) return;
  const BcgScalars c = bcg_scalars(scal, rho_cur);
  bcg_advance(scal, rr_new, scal[rho_new], c.ok, rtol, atol, maxit, state);
}

extern "C" int fsb_solve_bicgstab(fsb_mat* A, fsb_vec* b, fsb_vec* x, double rtol, double atol, int32_t maxit,
                                  int32_t precond, fsb_solve_info* info) {
  if (!A || !b || !x || !info) return FSB_ERR_ARG;
  fsb_ctx* ctx = A->ctx;
  const int64_t n = A->nbrows * A->bs;
  if (b->n != n || x->n != n) FSB_FAIL(ctx, FSB_ERR_ARG, "vector sizes do not match the matrix");
  memset(info, 0, sizeof(*info));
  const int64_t n0 = A->own0 * A->bs, n1 = A->own1 * A->bs;
  const bool dist = fsb_dist_active(ctx);
  cudaEvent_t e0, e1;
  FSB_CHECK_CUDA(ctx, cudaEventCreate(&e0));
  FSB_CHECK_CUDA(ctx, cudaEventCreate(&e1));
  struct EvGuard { cudaEvent_t a, b; ~EvGuard() { cudaEventDestroy(a); cudaEventDestroy(b); } } guard{e0, e1};
  FSB_CHECK_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
  int rc;
  fsb_mat* S = A;
  if (ctx->drop_zeros && (rc = fsb_mat_squeeze(A, &S))) return rc;
  info->operand_nnzb = S->nnzb;
  SpmvTimer timer;
  Workspace ws{A};
  double *r, *rhat, *p, *ph, *v, *sh, *t, *dinv;
  if ((rc = ws.alloc(&r, n)) || (rc = ws.alloc(&rhat, n)) || (rc = ws.alloc(&p, n)) || (rc = ws.alloc(&ph, n)) ||
      (rc = ws.alloc(&v, n)) || (rc = ws.alloc(&sh, n)) || (rc = ws.alloc(&t, n)) || (rc = ws.alloc(&dinv, n)))
    return rc;
  double* scal = ctx->d_scalars;
  int* state = ctx->d_state;
  const unsigned vg = vec_grid(ctx, n1 - n0);
  FSB_CHECK_CUDA(ctx, cudaMemsetAsync(state, 0, sizeof(int) * 8, ctx->stream));
  // last-CTA counters: a kernel that observes `done` half-way may leave one partially counted
  FSB_CHECK_CUDA(ctx, cudaMemsetAsync(ctx->d_counters, 0, sizeof(unsigned) * 16, ctx->stream));
#define DINV_LAUNCH(BS) k_extract_dinv<BS><<<fsb_grid(n1 - n0, 256, (int64_t)ctx->sm_count * 16), 256, 0, ctx->stream>>>(n0, n1, A->row_ptr, A->col_idx, A->vals, dinv, precond == 1)
  if (A->bs == 1) DINV_LAUNCH(1); else if (A->bs == 2) DINV_LAUNCH(2); else DINV_LAUNCH(3);
#undef DINV_LAUNCH
  FSB_LAUNCH_CHECK(ctx);
  // (re)start from x: r = b - A x, rhat = p = r.  A breakdown (a recurrence scalar hits zero or stops being finite:
  // with rho carried by recurrence that happens once the residual is down at rounding level, or on tiny systems
  // that are solved exactly half-way through an iteration) leaves x at its last valid value, so the standard
  // remedy applies: restart the recurrences from the true residual, a few times at most.
  int launched = 0, restarts = 0;
  bool finished = false;
  while (!finished) {
    if (dist && (rc = fsb_dist_halo_raw(ctx, x->d, n))) return rc;
    if ((rc = fsb_launch_spmv(S, x->d, v, nullptr, 0, nullptr, nullptr))) return rc;
    k_bcg_init<<<vg, kVecThreads, 0, ctx->stream>>>(n0, n1, b->d, v, dinv, r, rhat, p, ph, ctx->d_partials, scal + S_RHO0, ctx->d_counters + 1);
    FSB_LAUNCH_CHECK(ctx);
    if (dist && (rc = fsb_dist_allreduce_sum_dev(ctx, scal + S_RHO0, 3))) return rc;
    k_check0<<<1, 1, 0, ctx->stream>>>(scal, S_RRB0, S_BBB, rtol, atol, maxit, state, scal + S_FINAL_RR, restarts > 0);
    FSB_LAUNCH_CHECK(ctx);

    int first, rest;
    batch_plan(ctx, A->last_iters, &first, &rest);
    if (ctx->profile) timer.ensure(2 * std::max(first, rest));      // two SpMVs per iteration, collected after every batch's sync
    const int base = launched;
    bool stopped = false;
    while (!stopped) {
      const int batch = launched == base ? first : rest;
      for (int k = 0; k < batch; ++k) {
        const int par = (launched - base + k) & 1;
        const int rho = par ? S_RHO1 : S_RHO0, rhon = par ? S_RHO0 : S_RHO1, rrn = par ? S_RRB0 : S_RRB1;
        if (dist && (rc = fsb_dist_halo_raw(ctx, ph, n))) return rc;
        if (ctx->profile) cudaEventRecord(timer.next(0), ctx->stream);
        if ((rc = fsb_launch_spmv(S, ph, v, rhat, 0, scal + S_RV, state))) return rc;
        if (ctx->profile) cudaEventRecord(timer.next(0), ctx->stream);
        if (dist && (rc = fsb_dist_allreduce_sum_dev(ctx, scal + S_RV, 1))) return rc;
        k_bcg_s<<<vg, kVecThreads, 0, ctx->stream>>>(n0, n1, scal, rho, v, rhat, dinv, r, sh, ctx->d_partials, scal + S_RS, ctx->d_counters + 5, state);
        FSB_LAUNCH_CHECK(ctx);
        if (dist && (rc = fsb_dist_halo_raw(ctx, sh, n))) return rc;
        if (ctx->profile) cudaEventRecord(timer.next(0), ctx->stream);
        if ((rc = fsb_launch_spmv(S, sh, t, r, 1, scal + S_TS, state, nullptr, rhat))) return rc;
        if (ctx->profile) cudaEventRecord(timer.next(0), ctx->stream);
        if (dist && (rc = fsb_dist_allreduce_sum_dev(ctx, scal + S_TS, 4))) return rc;        // t.s, t.t, rhat.t, rhat.s
        k_bcg_xp<<<vg, kVecThreads, 0, ctx->stream>>>(n0, n1, scal, rho, rhon, rrn, sh, t, v, dinv, x->d, r, p, ph, ctx->d_partials,
                                                      ctx->d_counters + 2, rtol, atol, maxit, state, dist ? 0 : 1);
        FSB_LAUNCH_CHECK(ctx);
        if (dist) {
          if ((rc = fsb_dist_allreduce_sum_dev(ctx, scal + rrn, 1))) return rc;
          k_bcg_check<<<1, 1, 0, ctx->stream>>>(scal, rho, rhon, rrn, rtol, atol, maxit, state);
          FSB_LAUNCH_CHECK(ctx);
        }
      }
      launched += batch;
      FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(ctx->h_state, state, sizeof(int) * 4, cudaMemcpyDeviceToHost, ctx->stream));
      FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      if (ctx->profile) timer.collect(0);
      if (ctx->h_state[0] || launched >= maxit) stopped = true;
    }
    // outcome -1 with iterations to spare: restart (state[1] keeps counting across restarts)
    if (ctx->h_state[0] && ctx->h_state[2] == -1 && restarts < 4 && ctx->h_state[1] < maxit) {
      ++restarts;
      static const bool trace = getenv("FSB_SOLVE_TRACE") != nullptr;
      if (trace) fprintf(stderr, "libfsb: BiCGStab restart %d after %d iterations\n", restarts, ctx->h_state[1]);
      FSB_CHECK_CUDA(ctx, cudaMemsetAsync(ctx->d_counters, 0, sizeof(unsigned) * 16, ctx->stream));
      continue;
    }
    finished = true;
  }
  FSB_CHECK_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
  rc = read_outcome(ctx, S_FINAL_RR, S_BBB, info);
  if (rc) return rc;
  A->last_iters = info->iterations;
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  info->solve_ms = ms;
  if (ctx->profile) info->spmv_ms = timer.total_ms;      // includes the launches of a batch's tail that early-exit on `done`
  // restarts exhausted with a valid iterate: the recurrences stall at rounding level; report "not converged" with the
  // last valid residual norm instead of an error (x is the best iterate).  Non-finite data still raises.
  if (info->converged < 0 && restarts >= 4 && std::isfinite(info->rnorm) && info->rnorm > 0.0) info->converged = 0;
  if (info->converged < 0) FSB_FAIL(ctx, FSB_ERR_BREAKDOWN, "BiCGStab breakdown (zero or non-finite recurrence scalar)");
  return FSB_OK;
}
