// K7c: the whole Jacobi-PCG iteration loop as ONE persistent kernel per solve (single GPU and z-slab distributed).
//
// What it replaces: two (single-reduction CG) or three (classic CG) dependent launches per iteration, each ending in a
// grid-wide "last CTA" reduction tail and — distributed — each starting with every CTA polling the peers' mailboxes.
// At 8 GPUs a 256^3 iteration has 98 us of HBM work and cost 194 us that way.
//
// Layout of the kernel (cooperative launch: all CTAs are co-resident for the whole solve):
//   * W worker CTAs = the SpMV kernel's own shape (consumer warps + one TMA producer warp, fsb_spmv_core.cuh) and ONE
//     service CTA that moves no data: it folds the per-CTA partial sums, posts them to every rank's mailbox, and raises
//     the neighbours' halo flags.  Workers never run a reduction tail;
//   * Chronopoulos-Gear recurrences (fsb_solve.cu: same arithmetic as k_cg1_update + the SpMV with w.u fused), so an
//     iteration has one exposed reduction.  Per iteration:
//       top   every CTA reads (r.u, u.u, w.u) of the current iterate from its own rank's mailbox (ranks added in rank
//             order: bitwise the same alpha, beta and stop decision in every CTA of every rank);
//       U     consumers update p, s, x, r, u over a fixed row slice; boundary planes of the new u go straight into the
//             neighbours' ghost planes (peer stores over NVLink); the producer warp is already prefetching the first
//             SpMV tiles of this iteration (the matrix does not depend on u);
//       B1    grid barrier on a monotone arrival counter (u is complete on this rank); the service CTA then raises the
//             neighbours' halo flags and posts (r.u, u.u) — that exchange overlaps the SpMV;
//       S     w = A u over the CTA's tiles, interior tiles first; only the tiles that gather from a ghost plane wait
//             for the neighbour's flag (by then it has long been raised);
//       B2    arrive; the service CTA posts w.u; the wait for it is the `top` of the next iteration.
//   * every spin loop has a watchdog (trap after `cg_timeout_s`), so a lost peer ends in an error, not a hung GPU.
// Results: identical recurrences to the two-kernel single-reduction path; reductions are summed in CTA order, then
// rank order, so runs are bitwise reproducible.
#include "fsb_spmv_core.cuh"
#include <algorithm>
#include <cmath>

#include "fsb_cgp.cuh"

struct CgpArgs {
  SpmvArgs sp;                 // matrix + tiling; x = u (gathered), y = w, w = u (fused w.u)
  int64_t n0, n1;              // owned scalar rows
  const double* dinv;
  double *u, *w, *p, *s, *x, *r;
  double* scal;                // ctx scalars: [bb_slot] in, [final_slot] out
  int bb_slot, final_slot;
  int* state;                  // [0] done, [1] iterations, [2] outcome
  double rtol, atol;
  int maxit;
  unsigned long long* arrive;  // monotone arrival counter (zero at launch)
  double* part_u;              // [2][kMaxPartials]: (r.u, u.u) partials of phase U
  double* part_s;              // [kMaxPartials]: w.u partials of phase S
  PeerComm pc;                 // nranks == 1: buf[0] is a local CommBuf
  unsigned long long seq_base;
  unsigned long long timeout_ns;
  int debug;                   // timing experiments only (results are wrong): 1 skip the peer stores, 2 skip their system fence
  int umode;                   // update-phase row partition: 0 one contiguous slice per CTA, 1 grid-stride
  unsigned long long* phase_ns;  // [4] accumulated by worker 0: U, B1 wait, S, top wait (profile mode) or null
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// streaming 16-byte load that does not allocate a line in L1: the kernel runs with ~200 KB of the SM's 256 KB configured as
// shared memory and the update phase's seven streams compete for what is left of L1 (measured: 178.5 -> 169.1 us per update
// phase on the 2-GPU slab of 256^3 against ld.global.cs, profiles/cg_offset_ab_r2.txt)
__device__ __forceinline__ double2 ld_stream2(const double* p) {
  double2 v;
  asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ void bar_consumers(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

// spin until *p >= want (SYS: a value written by a peer GPU).  A peer that never arrives ends the kernel with a trap
// (reported by the next CUDA call on the host) instead of hanging the device.
template <bool SYS>
__device__ __forceinline__ void spin_until(const unsigned long long* p, unsigned long long want, unsigned long long timeout_ns) {
  unsigned spins = 0;
  unsigned long long t0 = 0;
  while ((SYS ? ld_acquire_sys(p) : ld_acquire_gpu(p)) < want) {
    if ((++spins & 0x3fffu) == 0) {
      const unsigned long long now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > timeout_ns) __trap();
    }
  }
}

// mbarrier wait with the same watchdog (a pipeline bug must not hang the device either)
__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity, unsigned long long timeout_ns) {
  uint32_t ok;
  unsigned spins = 0;
  unsigned long long t0 = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (!ok && (++spins & 0x3ffu) == 0) {
      const unsigned long long now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > timeout_ns) __trap();
    }
  } while (!ok);
}

// sum over the consumer threads (named barrier 1), result valid in thread 0
__device__ __forceinline__ double consumer_sum(double v, double* red, int nthreads) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  bar_consumers(nthreads);
  if (lane == 0) red[w] = v;
  bar_consumers(nthreads);
  v = ((int)threadIdx.x < (nthreads >> 5)) ? red[threadIdx.x] : 0.0;
  if (w == 0) v = warp_sum(v);
  return v;
}

template <int BS, int ROWS, int LPR, int NST, int MINB>
__global__ void __launch_bounds__(SpmvCfg<BS, ROWS, LPR>::THREADS, MINB) k_cg_persist(CgpArgs a) {
  using Cfg = SpmvCfg<BS, ROWS, LPR>;
  constexpr int CONSUMERS = Cfg::CONSUMERS;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[NST], empty[NST];
  __shared__ int64_t s_info[NST][4];
  __shared__ double red[32];
  __shared__ double s_mail[2][2][kMaxRanks][2];     // [iteration parity][slot: RZ, PQ][rank][value]
  const SpmvStage<BS, ROWS, LPR> st(smem, a.sp.cap);
  const int tid = threadIdx.x;
  const int W = (int)gridDim.x - 1;                 // worker CTAs; the last CTA is the service CTA
  const int c = blockIdx.x;
  const bool service = c == W;
  const bool peer = a.pc.nranks > 1;
  CommBuf* const mybuf = a.pc.buf[a.pc.rank];

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NST; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], CONSUMERS / 32);
    }
    mbar_fence_init();
  }
  __syncthreads();

  // tiles that gather from a ghost plane of u: the first tiles_lo (rows of the first owned plane) and the last tiles_hi
  // (rows of the last owned plane); the producer schedules them after the interior tiles
  const int64_t ntiles = a.sp.ntiles;
  int64_t nlo = 0, nhi = 0;
  if (peer && !service && tid == CONSUMERS) {
    const int64_t prow = a.pc.plane / BS;
    if (a.pc.rank > 0) {                                   // tiles with tile_row[t] < own0 + prow
      int64_t lo = 0, hi = ntiles;
      const int64_t key = a.sp.own0 + prow;
      while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (a.sp.tile_row[mid] < key) lo = mid + 1; else hi = mid; }
      nlo = lo;
    }
    if (a.pc.rank < a.pc.nranks - 1) {                     // tiles with tile_row[t+1] > own1 - prow
      int64_t lo = 0, hi = ntiles;
      const int64_t key = a.sp.own1 - prow + 1;
      while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (a.sp.tile_row[mid + 1] < key) lo = mid + 1; else hi = mid; }
      nhi = ntiles - lo;
    }
    if (nlo + nhi >= ntiles) { nlo = ntiles; nhi = 0; }    // thin slab: every tile waits
  }
  const int64_t nint = ntiles - nlo - nhi;

  const double bb = a.scal[a.bb_slot];
  const double tol2 = fmax(a.rtol * a.rtol * bb, a.atol * a.atol);
  double g_prev = 1.0, a_prev = 1.0;
  uint32_t tcount = 0;                  // tiles this CTA has put through the pipeline (same value in producer and consumers)
  const uint64_t policy = a.sp.l2_hint ? l2_evict_first_policy() : 0ull;
  unsigned long long t_mark = 0;
  const bool timing = a.phase_ns != nullptr && c == 0 && tid == 0;
  if (timing) t_mark = globaltimer_ns();

  if (service && tid >= 32) return;     // the service CTA works with one warp (its block-wide barriers below are warp barriers)
  for (int it = 0;; ++it) {
    const int par = it & 1;
    const unsigned long long seq = a.seq_base + (unsigned long long)it;
    // ---------------- top: the three inner products of the current iterate, all ranks, rank order
    if (tid < 2 * kMaxRanks) {
      const int which = tid / kMaxRanks, rk = tid % kMaxRanks;
      if (rk < a.pc.nranks) {
        const MailEntry* e = &mybuf->mail[(which ? MAIL_PQ : MAIL_RZ) + par][rk];
        spin_until<true>(&e->seq, seq, a.timeout_ns);
        s_mail[par][which][rk][0] = *reinterpret_cast<const volatile double*>(&e->v[0]);
        s_mail[par][which][rk][1] = *reinterpret_cast<const volatile double*>(&e->v[1]);
      }
    }
    if (service) __syncwarp(); else __syncthreads();
    if (timing) { const unsigned long long t = globaltimer_ns(); a.phase_ns[3] += t - t_mark; t_mark = t; }
    double g = 0.0, zz = 0.0, d = 0.0;
    for (int rk = 0; rk < a.pc.nranks; ++rk) { g += s_mail[par][0][rk][0]; zz += s_mail[par][0][rk][1]; d += s_mail[par][1][rk][0]; }
    const double beta = it > 0 ? g / g_prev : 0.0;
    const double denom = it > 0 ? d - beta * g / a_prev : d;
    const double alpha = g / denom;
    const bool conv = zz <= tol2;
    const bool broken = !(zz == zz) || !(alpha == alpha) || denom == 0.0 || (it > 0 && g_prev == 0.0);
    if (conv || broken || it >= a.maxit) {          // the same decision in every CTA of every rank
      if (c == 0 && tid == 0) {
        a.state[1] = it;
        a.state[2] = conv ? 1 : (broken ? -1 : 0);
        a.scal[a.final_slot] = zz;
        __threadfence();
        a.state[0] = 1;
      }
      break;
    }
    g_prev = g; a_prev = alpha;

    if (service) {
      // ---------------- service CTA (its first warp; the others have left): B1 -> halo flags + (r.u, u.u); B2 -> w.u.
      // One warp, shuffles only: nothing on this path waits for a block-wide barrier
      if (tid == 0) spin_until<false>(a.arrive, (unsigned long long)(2 * it + 1) * (unsigned long long)W, a.timeout_ns);
      __syncwarp();
      if (peer && tid == 0) {         // every worker's peer stores happen-before its arrival: the new planes are in place
        if (a.pc.rank > 0) st_release_sys(&a.pc.buf[a.pc.rank - 1]->halo_flag[1], seq + 1);
        if (a.pc.rank < a.pc.nranks - 1) st_release_sys(&a.pc.buf[a.pc.rank + 1]->halo_flag[0], seq + 1);
      }
      double t0 = 0.0, t1 = 0.0;
      for (int i = tid; i < W; i += 32) { t0 += __ldcg(a.part_u + i); t1 += __ldcg(a.part_u + kMaxPartials + i); }
      t0 = warp_sum(t0); t1 = warp_sum(t1);
      if (tid < a.pc.nranks) {
        MailEntry* e = &a.pc.buf[tid]->mail[MAIL_RZ + (par ^ 1)][a.pc.rank];
        e->v[0] = t0; e->v[1] = t1;
        st_release_sys(&e->seq, seq + 1);
      }
      if (tid == 0) spin_until<false>(a.arrive, (unsigned long long)(2 * it + 2) * (unsigned long long)W, a.timeout_ns);
      __syncwarp();
      double t2 = 0.0;
      for (int i = tid; i < W; i += 32) t2 += __ldcg(a.part_s + i);
      t2 = warp_sum(t2);
      if (tid < a.pc.nranks) {
        MailEntry* e = &a.pc.buf[tid]->mail[MAIL_PQ + (par ^ 1)][a.pc.rank];
        e->v[0] = t2; e->v[1] = 0.0;
        st_release_sys(&e->seq, seq + 1);
      }
      continue;
    }

    if (tid >= CONSUMERS) {
      // ---------------- producer warp: this iteration's tiles (prefetch starts while the consumers are in phase U)
      bool halo_ok = !peer;
      for (int64_t j = c; j < ntiles; j += W, ++tcount) {
        const int s = tcount % NST;
        if (tid == CONSUMERS) {
          if (tcount >= NST) mbar_wait_wd(&empty[s], ((tcount / NST) - 1) & 1, a.timeout_ns);
          int64_t tile = j;
          if (nlo < ntiles) {           // interior tiles first, then the ghost-reading ones
            if (j < nint) tile = nlo + j;
            else { const int64_t b = j - nint; tile = b < nlo ? b : ntiles - nhi + (b - nlo); }
          }
          if (!halo_ok && (j >= nint || nlo >= ntiles)) {
            if (a.pc.rank > 0) spin_until<true>(&mybuf->halo_flag[0], seq + 1, a.timeout_ns);
            if (a.pc.rank < a.pc.nranks - 1) spin_until<true>(&mybuf->halo_flag[1], seq + 1, a.timeout_ns);
            halo_ok = true;
          }
          spmv_issue_tile<BS, ROWS, LPR>(a.sp, st, tile, s, s_info[s], &full[s], policy);
        }
        __syncwarp();
      }
      continue;
    }

    // ---------------- consumers, phase U: p = u + beta p ; s = w + beta s ; x += alpha p ; r -= alpha s ; u = dinv r
    {
      const int64_t n0 = a.n0, n1 = a.n1;
      const int64_t plane = a.pc.plane;
      double s0 = 0.0, s1 = 0.0;
      bool pushed = false;
      auto push = [&](int64_t i, double v) {       // owned boundary planes -> neighbours' ghost planes
        if (a.debug & 1) return;
        if (a.pc.lo_dst && i < n0 + plane) { a.pc.lo_dst[i - n0] = v; pushed = true; }
        if (a.pc.hi_dst && i >= n1 - plane) { a.pc.hi_dst[i - (n1 - plane)] = v; pushed = true; }
      };
      auto one = [&](int64_t i) {
        const double pi = a.u[i] + beta * __ldcs(a.p + i);
        const double si = __ldcs(a.w + i) + beta * __ldcs(a.s + i);
        __stcs(a.p + i, pi); __stcs(a.s + i, si);
        __stcs(a.x + i, __ldcs(a.x + i) + alpha * pi);
        const double ri = __ldcs(a.r + i) - alpha * si;
        __stcs(a.r + i, ri);
        const double ui = __ldcs(a.dinv + i) * ri;
        a.u[i] = ui;
        s0 += ri * ui; s1 += ui * ui;
        if (peer) push(i, ui);
      };
      // the body starts on a 512-byte boundary = one warp request of 16-byte accesses.  A rank with a lower ghost plane owns rows
      // from an odd offset on; measured on the 2-GPU slab of 256^3 (profiles/cg_offset_ab_r2.txt): 16-byte aligned body 218 us,
      // 128-byte 194-199 us, 512-byte 179 us = the time of the rank whose range starts at 0
      const int64_t amask = 63;
      const int64_t a_up = (n0 + amask) & ~amask;
      const int64_t a0 = a_up < n1 ? a_up : n1, npair = (n1 - a0) >> 1;
      if (c == 0) for (int64_t i = n0 + tid; i < a0; i += CONSUMERS) one(i);
      if (c == 1 % W && tid == 32 && a0 + 2 * npair < n1) one(n1 - 1);
      const int64_t jq = npair / W, jr = npair % W;
      const int64_t j0 = a.umode ? (int64_t)c * CONSUMERS : jq * c + (c < jr ? c : jr);
      const int64_t j1 = a.umode ? npair : j0 + jq + (c < jr ? 1 : 0);
      const int64_t jstep = a.umode ? (int64_t)W * CONSUMERS : CONSUMERS;
      for (int64_t j = j0 + tid; j < j1; j += jstep) {
        const int64_t i = a0 + 2 * j;
        const double2 uv = *reinterpret_cast<const double2*>(a.u + i);
        const double2 wv = ld_stream2(a.w + i), dv = ld_stream2(a.dinv + i);
        double2 pv = ld_stream2(a.p + i), sv = ld_stream2(a.s + i), xv = ld_stream2(a.x + i), rv = ld_stream2(a.r + i);
        pv.x = uv.x + beta * pv.x; pv.y = uv.y + beta * pv.y;
        sv.x = wv.x + beta * sv.x; sv.y = wv.y + beta * sv.y;
        xv.x += alpha * pv.x; xv.y += alpha * pv.y;
        rv.x -= alpha * sv.x; rv.y -= alpha * sv.y;
        double2 un;
        un.x = dv.x * rv.x; un.y = dv.y * rv.y;
        __stcs(reinterpret_cast<double2*>(a.p + i), pv);
        __stcs(reinterpret_cast<double2*>(a.s + i), sv);
        __stcs(reinterpret_cast<double2*>(a.x + i), xv);
        __stcs(reinterpret_cast<double2*>(a.r + i), rv);
        *reinterpret_cast<double2*>(a.u + i) = un;
        s0 += rv.x * un.x + rv.y * un.y;
        s1 += un.x * un.x + un.y * un.y;
        if (peer) { push(i, un.x); push(i + 1, un.y); }
      }
      if (pushed && !(a.debug & 2)) __threadfence_system();
      s0 = consumer_sum(s0, red, CONSUMERS);
      s1 = consumer_sum(s1, red, CONSUMERS);
      // ---------------- B1: u is complete on this rank once every worker has arrived
      if (tid == 0) {
        a.part_u[c] = s0; a.part_u[kMaxPartials + c] = s1;
        __threadfence();
        atomicAdd(a.arrive, 1ull);
        if (timing) { const unsigned long long t = globaltimer_ns(); a.phase_ns[0] += t - t_mark; t_mark = t; }
        spin_until<false>(a.arrive, (unsigned long long)(2 * it + 1) * (unsigned long long)W, a.timeout_ns);
        if (timing) { const unsigned long long t = globaltimer_ns(); a.phase_ns[1] += t - t_mark; t_mark = t; }
      }
      bar_consumers(CONSUMERS);
    }
    // ---------------- consumers, phase S: w = A u (+ w.u)
    {
      double d0 = 0.0, d1 = 0.0, d2 = 0.0;
      for (int64_t j = c; j < ntiles; j += W, ++tcount) {
        const int s = tcount % NST;
        mbar_wait_wd(&full[s], (tcount / NST) & 1, a.timeout_ns);
        spmv_consume_tile<BS, ROWS, LPR, true>(a.sp, st, s, s_info[s], d0, d1, d2);
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&empty[s]);
      }
      d0 = consumer_sum(d0, red, CONSUMERS);
      if (tid == 0) {         // B2: arrive only; the wait is the mailbox poll at the top of the next iteration
        a.part_s[c] = d0;
        __threadfence();
        atomicAdd(a.arrive, 1ull);
        if (timing) { const unsigned long long t = globaltimer_ns(); a.phase_ns[2] += t - t_mark; t_mark = t; }
      }
    }
  }
}

// ------------------------------------------------------------------------------------ host side
bool fsb_cgp_supported(fsb_mat* S) {
  SpmvPlan pl;
  if (fsb_spmv_plan(S, &pl) != FSB_OK) return false;
  return (pl.bs == 1 && pl.rows == 256 && pl.lpr == 2 && pl.nst == 2) || (pl.bs == 1 && pl.rows == 256 && pl.lpr == 1 && pl.nst == 2) ||
         (pl.bs == 3 && pl.rows == 192 && pl.lpr == 2 && pl.nst == 3);
}

int fsb_cgp_run(fsb_mat* A, fsb_mat* S, const CgpVectors& v, double rtol, double atol, int maxit, int bb_slot, int final_slot,
                const PeerComm& pc, unsigned long long seq_base, double* phase_ms) {
  fsb_ctx* ctx = A->ctx;
  SpmvPlan pl;
  int rc = fsb_spmv_plan(S, &pl);
  if (rc) FSB_FAIL(ctx, FSB_ERR_STATE, "persistent CG needs the staged SpMV kernel");
  CgpArgs a;
  fsb_spmv_fill_args(S, &a.sp);
  a.sp.x = v.u; a.sp.y = v.w; a.sp.w = v.u;
  a.n0 = A->own0 * A->bs; a.n1 = A->own1 * A->bs;
  a.dinv = v.dinv; a.u = v.u; a.w = v.w; a.p = v.p; a.s = v.s; a.x = v.x; a.r = v.r;
  a.scal = ctx->d_scalars; a.bb_slot = bb_slot; a.final_slot = final_slot; a.state = ctx->d_state;
  a.rtol = rtol; a.atol = atol; a.maxit = maxit;
  a.arrive = reinterpret_cast<unsigned long long*>(ctx->d_counters + 8);
  a.part_u = ctx->d_partials; a.part_s = ctx->d_partials + 2 * kMaxPartials;
  a.pc = pc; a.seq_base = seq_base;
  a.timeout_ns = (unsigned long long)std::max(1, ctx->cg_timeout_s) * 1000000000ull;
  a.phase_ns = nullptr;
  a.umode = ctx->cg_umode;
  a.debug = ctx->cg_debug;
  if (ctx->profile) {
    a.phase_ns = reinterpret_cast<unsigned long long*>(ctx->d_scalars + 48);
    FSB_CHECK_CUDA(ctx, cudaMemsetAsync(a.phase_ns, 0, 4 * sizeof(unsigned long long), ctx->stream));
  }
  FSB_CHECK_CUDA(ctx, cudaMemsetAsync(a.arrive, 0, sizeof(unsigned long long), ctx->stream));

  const void* fn = nullptr;
  int threads = 0;
#define FSB_CGP_CASE(BS, ROWS, LPR, NST, MINB, COND)                                                 \
  if (!fn && pl.bs == BS && pl.rows == ROWS && pl.lpr == LPR && pl.nst == NST && (COND)) {           \
    fn = (const void*)k_cg_persist<BS, ROWS, LPR, NST, MINB>;                                        \
    threads = SpmvCfg<BS, ROWS, LPR>::THREADS;                                                       \
  }
  // The one-lane (short-row, squeezed operand) configuration exists at two register budgets.  4 CTAs per SM (56 registers) bring its
  // SpMV phase to the stand-alone kernel's rate (6 390 against 4 980 GB/s on the 17 M-row operand) at the price of a slower update
  // phase and a dearer grid barrier (591 instead of 295 arrivals): measured (tools/cg_ab.py, profiles/cg_ab_r2.txt) 0.627 against
  // 0.656 ms per iteration at 17 M rows but 90.1 against 80.6 us at 2.1 M rows (one rank of an 8-GPU 256^3 run).  Large slabs take 4.
  const int64_t nrows = (A->own1 - A->own0) * A->bs;
  const int want4 = ctx->cg_minb ? ctx->cg_minb == 4 : nrows >= 6000000;
  FSB_CGP_CASE(1, 256, 2, 2, 2, true)
  FSB_CGP_CASE(1, 256, 1, 2, 4, want4)
  FSB_CGP_CASE(1, 256, 1, 2, 2, true)
  FSB_CGP_CASE(3, 192, 2, 3, 1, true)
#undef FSB_CGP_CASE
  if (!fn) FSB_FAIL(ctx, FSB_ERR_STATE, "persistent CG: no kernel for this SpMV configuration");
  FSB_CHECK_CUDA(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
  int per_sm = 0;
  FSB_CHECK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, pl.smem));
  if (per_sm < 1) FSB_FAIL(ctx, FSB_ERR_STATE, "persistent CG kernel does not fit an SM");
  per_sm = std::min(per_sm, pl.per_sm);
  const int64_t resident = (int64_t)per_sm * ctx->sm_count;
  const int workers = (int)std::max<int64_t>(1, std::min<int64_t>({resident - 1, S->ntiles, (int64_t)kMaxPartials}));
  void* params[] = {&a};
  FSB_CHECK_CUDA(ctx, cudaLaunchCooperativeKernel(fn, dim3((unsigned)workers + 1), dim3((unsigned)threads), params, pl.smem, ctx->stream));
  ctx->launches++;
  if (phase_ms && ctx->profile) {
    unsigned long long ns[4];
    FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(ns, a.phase_ns, sizeof(ns), cudaMemcpyDeviceToHost, ctx->stream));
    FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < 4; ++k) phase_ms[k] = (double)ns[k] * 1e-6;
    static const bool trace = getenv("FSB_SOLVE_TRACE") != nullptr;
    if (trace)
      fprintf(stderr, "libfsb: rank %d persistent CG, %d workers x %d threads: update %.3f ms, barrier wait %.3f ms, SpMV %.3f ms, scalar wait %.3f ms (worker 0)\n",
              pc.rank, workers, threads, phase_ms[0], phase_ms[1], phase_ms[2], phase_ms[3]);
  }
  return FSB_OK;
}
