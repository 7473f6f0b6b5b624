// K7a: block-CSR SpMV  y = A x  (+ fused dot products) for the Krylov kernel chains.
//
// Design (HBM-bound: 12 B per non-zero + 24 B per row of algorithmic traffic):
//   * the non-zero stream is cut into tiles of ~T blocks snapped to row boundaries (equal bytes per tile;
//     per-tile first row and first non-zero are precomputed once per matrix);
//   * persistent CTAs, one producer warp + consumer warps.  The producer streams each tile's values,
//     column indices and row_ptr slice into shared memory with 1-D TMA bulk copies (cp.async.bulk ->
//     SASS UBLKCP) signalling a `full` mbarrier; consumers release the stage through an `empty`
//     mbarrier, so tile issue (two dependent index loads) is off the consumers' critical path and NST-1
//     tiles stay in flight per CTA;
//   * LPR lanes share one scalar row; each lane gathers up to UNR x-entries at once through L1/L2
//     (x is read once from HBM: neighbouring rows reuse it from L2), partial sums are combined by
//     warp shuffles, rows are written once;
//   * fused reductions (y.w, y.y) leave per-CTA partials that the last CTA combines in a fixed order.
// Modes kept for A/B measurements: 2 = first version (one thread per row, no specialisation),
// 1 = plain row-per-thread kernel without staging (also the fallback for untileable matrices).
#include "fsb_spmv_core.cuh"
#include <algorithm>
#include <cmath>

// ------------------------------------------------------------------------------------ mode 2: first version
template <int BS, int THREADS>
__global__ void __launch_bounds__(THREADS) k_spmv_tma(SpmvArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ double red[32];
  if (a.done && *a.done) return;
  constexpr int VB = 8 * BS * BS;   // bytes of values per block
  const size_t stage_bytes = (size_t)a.cap * (VB + 4);
  auto vals_s = [&](int s) { return reinterpret_cast<const double*>(smem + s * stage_bytes); };
  auto cols_s = [&](int s) { return reinterpret_cast<const int32_t*>(smem + s * stage_bytes + (size_t)a.cap * VB); };

  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();

  auto issue = [&](int64_t tile, int s) {
    const int64_t r0 = a.tile_row[tile], r1 = a.tile_row[tile + 1];
    if (r1 <= r0) { mbar_arrive(&bar[s]); return; }
    const int64_t k0 = a.row_ptr[r0], k1 = a.row_ptr[r1];
    const int64_t al0 = k0 & ~3ll;
    const uint32_t cnt = (uint32_t)(((k1 - al0) + 3) & ~3ll);
    mbar_expect_tx(&bar[s], cnt * (VB + 4));
    bulk_g2s((void*)vals_s(s), a.vals + al0 * BS * BS, cnt * VB, &bar[s]);
    bulk_g2s((void*)cols_s(s), a.col_idx + al0, cnt * 4, &bar[s]);
  };

  double d0 = 0.0, d1 = 0.0, d2 = 0.0;
  int64_t tile = blockIdx.x;
  if (threadIdx.x == 0 && tile < a.ntiles) issue(tile, 0);
  for (int it = 0; tile < a.ntiles; tile += gridDim.x, ++it) {
    const int s = it & 1;
    const int64_t next = tile + gridDim.x;
    if (threadIdx.x == 0 && next < a.ntiles) issue(next, s ^ 1);
    const int64_t r0 = a.tile_row[tile], r1 = a.tile_row[tile + 1];
    mbar_wait(&bar[s], (it >> 1) & 1);
    if (r1 > r0) {
      const int64_t al0 = a.row_ptr[r0] & ~3ll;
      const double* __restrict__ vs = vals_s(s);
      const int32_t* __restrict__ cs = cols_s(s);
      const int nscalar = (int)(r1 - r0) * BS;
      for (int lr = threadIdx.x; lr < nscalar; lr += THREADS) {
        const int64_t R = r0 + lr / BS;
        const int i = lr % BS;
        const int ks = (int)(a.row_ptr[R] - al0), ke = (int)(a.row_ptr[R + 1] - al0);
        double acc = 0.0;
#pragma unroll 4
        for (int k = ks; k < ke; ++k) {
          const int64_t c = cs[k];
#pragma unroll
          for (int j = 0; j < BS; ++j) acc += vs[(k * BS + i) * BS + j] * __ldg(a.x + c * BS + j);
        }
        const int64_t row = R * BS + i;
        a.y[row] = acc;
        if (a.w) d0 += acc * a.w[row];
        if (a.want_yy) d1 += acc * acc;
        if (a.w2) d2 += acc * a.w2[row];
      }
    }
    __syncthreads();   // stage s is free for the prefetch issued at the top of the next iteration
  }
  if (a.out) {
    if (a.w2) {
      double mine[3] = {block_sum(d0, red), block_sum(d1, red), block_sum(d2, red)};
      finish_partials<3>(mine, a.partials, kMaxPartials, a.out, a.counter, red);
    } else {
      double mine[2] = {block_sum(d0, red), block_sum(d1, red)};
      finish_partials<2>(mine, a.partials, kMaxPartials, a.out, a.counter, red);
    }
  }
}

// ------------------------------------------------------------------------------------ mode 0: producer/consumer
// FLAT: the two-phase tile consumer (spmv_consume_tile_flat) — a template parameter, not a run-time branch: carrying both consumers
// in one kernel cost the 256 x 2 configuration 16 registers and with them its second CTA per SM (measured: 0.88 -> 0.66 of peak)
template <int BS, int ROWS, int LPR, int NST, bool FLAT = false>
__global__ void __launch_bounds__(SpmvCfg<BS, ROWS, LPR>::THREADS) k_spmv_ws(SpmvArgs a) {
  using Cfg = SpmvCfg<BS, ROWS, LPR>;
  constexpr int CONSUMERS = Cfg::CONSUMERS;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[NST], empty[NST];
  __shared__ int64_t s_info[NST][4];    // per stage: r0, r1, aligned first nnz, aligned first row (or -1: row_ptr not staged)
  __shared__ double red[32];
  if (a.done && *a.done) return;
  const SpmvStage<BS, ROWS, LPR> st(smem, a.cap);

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < NST; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], CONSUMERS / 32);
    }
    mbar_fence_init();
    if (a.halo_seq) {     // ghost planes of x are written by the neighbours' p-update kernels (peer stores)
      const CommBuf* mine = a.pc.buf[a.pc.rank];
      if (a.pc.rank > 0) while (ld_acquire_sys(&mine->halo_flag[0]) < a.halo_seq) { }
      if (a.pc.rank < a.pc.nranks - 1) while (ld_acquire_sys(&mine->halo_flag[1]) < a.halo_seq) { }
    }
  }
  __syncthreads();

  double d0 = 0.0, d1 = 0.0, d2 = 0.0;
  if (threadIdx.x >= CONSUMERS) {
    // ===== producer warp: lane 0 walks the tiles and issues the bulk copies (the warp stays converged) =====
    const uint64_t policy = a.l2_hint ? l2_evict_first_policy() : 0ull;
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++it) {
      const int s = it % NST;
      if (threadIdx.x == CONSUMERS) {
        if (it >= NST) mbar_wait(&empty[s], ((it / NST) - 1) & 1);      // consumers have drained this stage
        spmv_issue_tile<BS, ROWS, LPR>(a, st, tile, s, s_info[s], &full[s], policy);
      }
      __syncwarp();
    }
  } else {
    // ===== consumer warps =====
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++it) {
      const int s = it % NST;
      mbar_wait(&full[s], (it / NST) & 1);
      if (FLAT) spmv_consume_tile_flat<BS, ROWS, LPR>(a, st, s, s_info[s], d0, d1, d2);
      else spmv_consume_tile<BS, ROWS, LPR, false>(a, st, s, s_info[s], d0, d1, d2);
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[s]);      // this warp is done with stage s
    }
  }
  if (a.mail_slot >= 0) {
    double mine[2];
    mine[0] = block_sum(d0, red);
    mine[1] = block_sum(d1, red);
    finish_partials_mail<2>(mine, a.partials, kMaxPartials, a.counter, red, a.pc, a.mail_slot, a.mail_seq);
  } else if (a.out) {
    if (a.w2) {
      double mine[3] = {block_sum(d0, red), block_sum(d1, red), block_sum(d2, red)};
      finish_partials<3>(mine, a.partials, kMaxPartials, a.out, a.counter, red);
    } else {
      double mine[2] = {block_sum(d0, red), block_sum(d1, red)};
      finish_partials<2>(mine, a.partials, kMaxPartials, a.out, a.counter, red);
    }
  }
}

// ------------------------------------------------------------------------------------ mode 1: plain
template <int BS>
__global__ void __launch_bounds__(256) k_spmv_plain(SpmvArgs a) {
  __shared__ double red[32];
  if (a.done && *a.done) return;
  double d0 = 0.0, d1 = 0.0, d2 = 0.0;
  const int64_t n0 = a.own0 * BS, n1 = a.own1 * BS;
  for (int64_t row = n0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; row < n1; row += (int64_t)gridDim.x * blockDim.x) {
    const int64_t R = row / BS;
    const int i = (int)(row % BS);
    double acc = 0.0;
    for (int64_t k = a.row_ptr[R]; k < a.row_ptr[R + 1]; ++k) {
      const int64_t c = a.col_idx[k];
#pragma unroll
      for (int j = 0; j < BS; ++j) acc += a.vals[(k * BS + i) * BS + j] * __ldg(a.x + c * BS + j);
    }
    a.y[row] = acc;
    if (a.w) d0 += acc * a.w[row];
    if (a.want_yy) d1 += acc * acc;
    if (a.w2) d2 += acc * a.w2[row];
  }
  if (a.out) {
    if (a.w2) {
      double mine[3] = {block_sum(d0, red), block_sum(d1, red), block_sum(d2, red)};
      finish_partials<3>(mine, a.partials, kMaxPartials, a.out, a.counter, red);
    } else {
      double mine[2] = {block_sum(d0, red), block_sum(d1, red)};
      finish_partials<2>(mine, a.partials, kMaxPartials, a.out, a.counter, red);
    }
  }
}

// ------------------------------------------------------------------------------------ tiling
// tile t covers block rows [tile_row[t], tile_row[t+1]): the first row whose row_ptr reaches the t-th
// multiple of tile_nnz past the owned range's first non-zero; tile_k[t] = row_ptr[tile_row[t]].
__global__ void k_tile_rows(const int64_t* __restrict__ row_ptr, int64_t own0, int64_t own1, int64_t tile_nnz,
                            int64_t ntiles, int64_t* __restrict__ tile_row, int64_t* __restrict__ tile_k) {
  const int64_t base = row_ptr[own0];
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t <= ntiles; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t lo = own0, hi = own1;          // first row r in [own0, own1] with row_ptr[r] >= target
    if (t == ntiles) lo = own1;
    else {
      const int64_t target = base + t * tile_nnz;
      while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (row_ptr[mid] < target) lo = mid + 1; else hi = mid;
      }
    }
    tile_row[t] = lo;
    tile_k[t] = row_ptr[lo];
  }
}

// Scalar rows per tile: 256 (192 for 3x3 blocks) or half of that.  Option value 0 = measured best per
// block size on B200 (profiles/spmv_sweep_r1.txt): CSR 256 rows x 2 lanes x 2 stages (two CTAs per SM beat
// a deeper pipeline with one); 3x3 blocks 192 rows x 2 lanes x 3 stages (one CTA per SM either way, so
// the third stage is free and keeps two 73 KB tiles in flight).
// Short rows (<= ~8 entries: a squeezed P1 operand, 2-D P1 matrices) are bound by per-row cost, not bytes,
// with 2 lanes x 8 slots per row: they get one lane per row (tools/spmv_short_rows.py on the squeezed C2
// operand: 256 rows x 1 lane x 2 stages 6 265 GB/s, x 2 lanes 3 062 GB/s; deeper pipelines lose the second CTA).
static bool spmv_short_rows(const fsb_mat* A) { return A->bs == 1 && A->avg_row > 0.0 && A->avg_row <= 8.5; }
// Long rows (degree-2 spaces: ~28 blocks per row on average, 10 to ~90 per row).  Measured on P2 operands at 48^3 and 64^3
// (tools/spmv_long_rows.py, profiles/spmv_long_rows_r2.txt, last table = after the flat consumer became a template parameter):
//   * 3x3 blocks: a tile of 192 scalar rows does not fit a stage (the plain kernel ran: 3 340 GB/s).  Half the rows, 4 lanes per
//     row, 3 stages and the two-phase ("flat") tile consumer — products over the tile's non-zeros, then row sums, so the gathers
//     are balanced whatever the row lengths: 4 680 GB/s (rows-per-lane form of the same tiling: 4 150);
//   * scalar CSR: 128 rows x 2 lanes x 2 stages, rows form: 4 250 GB/s (flat: 4 195, a tie; 256 rows x 2 lanes: 2 670 with two CTAs
//     per SM).  These operands gather 28 x-entries per row from two distant index ranges (vertex nodes, edge nodes): at 32 B per
//     gathered sector that is 0.9 KB of L2 -> SM traffic per row against 0.36 KB from HBM, so the L2 rate, not HBM, bounds them.
static bool spmv_long_rows(const fsb_mat* A) { return A->avg_row > 20.0; }
static int spmv_rows(fsb_ctx* ctx, const fsb_mat* A) {
  const int big = A->bs == 3 ? 192 : 256;
  const int opt = ctx->spmv_rows ? ctx->spmv_rows : (spmv_long_rows(A) ? 128 : 256);
  return opt == 128 ? big / 2 : (opt == 512 && A->bs == 1 ? 512 : big);
}
static int spmv_lpr(fsb_ctx* ctx, const fsb_mat* A) {
  return ctx->spmv_lpr ? ctx->spmv_lpr : (spmv_short_rows(A) ? 1 : (spmv_long_rows(A) && A->bs == 3 ? 4 : 2));
}
static bool spmv_flat(fsb_ctx* ctx, const fsb_mat* A) { return ctx->spmv_flat == 1 || (ctx->spmv_flat == 0 && spmv_long_rows(A) && A->bs == 3); }
static int spmv_stages(fsb_ctx* ctx, const fsb_mat* A) { return ctx->spmv_stages ? ctx->spmv_stages : (A->bs == 3 ? 3 : 2); }
static constexpr size_t kSmemBudget = 200 * 1024;

int fsb_mat_setup_tiles(fsb_mat* A) {
  fsb_ctx* ctx = A->ctx;
  fsb_dfree(A->ctx, A->tile_row);
  A->tile_row = nullptr;
  A->ntiles = 0; A->tile_nnz = 0; A->tile_cap = 0; A->stage_bytes = 0;
  A->avg_row = 0.0;
  A->tile_rows = spmv_rows(ctx, A);
  const int64_t nrows = A->own1 - A->own0;
  if (nrows <= 0) return FSB_OK;
  int64_t k01[2];
  FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(&k01[0], A->row_ptr + A->own0, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(&k01[1], A->row_ptr + A->own1, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const int64_t nnz = k01[1] - k01[0];
  if (nnz <= 0) return FSB_OK;
  const int bs = A->bs;
  const double avg = (double)nnz / (double)nrows;
  A->avg_row = avg;
  A->tile_rows = spmv_rows(ctx, A);
  const int rows_target = A->tile_rows / bs;
  int64_t T = (int64_t)std::ceil(avg * rows_target);
  T = (T + 15) & ~15ll;
  int64_t cap = T + A->max_row_len + 8;
  cap = (cap + 3) & ~3ll;
  const size_t stage = (size_t)cap * (8 * bs * bs + 4) + (size_t)(2 * rows_target + 12) * 8;
  if (2 * stage > kSmemBudget) return FSB_OK;   // not tileable (very long rows): plain kernel is used
  A->stage_bytes = stage;
  A->tile_nnz = (int)T;
  A->tile_cap = (int)cap;
  A->ntiles = (nnz + T - 1) / T;
  int rc = fsb_dmalloc(ctx, &A->tile_row, 2 * ((size_t)A->ntiles + 1));
  if (rc) return rc;
  k_tile_rows<<<fsb_grid(A->ntiles + 1, 256, 4096), 256, 0, ctx->stream>>>(A->row_ptr, A->own0, A->own1, T, A->ntiles, A->tile_row,
                                                                           A->tile_row + A->ntiles + 1);
  FSB_LAUNCH_CHECK(ctx);
  return FSB_OK;
}

// ------------------------------------------------------------------------------------ launch
bool fsb_spmv_supports_p2p(fsb_mat* A) { return A->ctx->spmv_mode == 0 && A->ntiles > 0; }

// the staged kernel's configuration for A, and the matrix part of its arguments (shared with the persistent CG kernel)
int fsb_spmv_plan(fsb_mat* A, SpmvPlan* plan) {
  fsb_ctx* ctx = A->ctx;
  if (A->own1 <= A->own0) return FSB_ERR_STATE;
  if (ctx->spmv_mode != 1 && A->tile_rows != spmv_rows(ctx, A)) {
    int rc = fsb_mat_setup_tiles(A);      // the tile size option changed since the matrix was set up
    if (rc) return rc;
  }
  if (ctx->spmv_mode != 0 || A->ntiles <= 0) return FSB_ERR_STATE;
  plan->bs = A->bs;
  plan->rows = A->tile_rows;
  plan->lpr = spmv_lpr(ctx, A);
  plan->nst = std::max(2, std::min(spmv_stages(ctx, A), (int)((224 * 1024) / A->stage_bytes)));
  plan->smem = (size_t)plan->nst * A->stage_bytes;
  const int threads = plan->rows * plan->lpr + 32;
  plan->per_sm = (int)std::max<size_t>(1, std::min<size_t>(2048 / threads, (227 * 1024) / (plan->smem + 1024)));
  return FSB_OK;
}

void fsb_spmv_fill_args(fsb_mat* A, SpmvArgs* a) {
  fsb_ctx* ctx = A->ctx;
  a->row_ptr = A->row_ptr; a->col_idx = A->col_idx; a->vals = A->vals;
  a->tile_row = A->tile_row; a->tile_k = A->tile_row ? A->tile_row + A->ntiles + 1 : nullptr;
  a->ntiles = A->ntiles; a->own0 = A->own0; a->own1 = A->own1; a->cap = A->tile_cap;
  a->x = nullptr; a->y = nullptr; a->w = nullptr; a->want_yy = 0; a->w2 = nullptr; a->l2_hint = ctx->spmv_hint;
  a->flat = spmv_flat(ctx, A);
  a->partials = ctx->d_partials; a->out = nullptr; a->counter = ctx->d_counters + 0; a->done = nullptr;
  memset(&a->pc, 0, sizeof(a->pc));
  a->halo_seq = 0; a->mail_slot = -1; a->mail_seq = 0;
}

int fsb_launch_spmv(fsb_mat* A, const double* x, double* y, const double* w, int want_yy, double* out, const int* done,
                    const fsb_spmv_dist* dd, const double* w2) {
  fsb_ctx* ctx = A->ctx;
  if (A->own1 <= A->own0) return FSB_OK;
  if (ctx->spmv_mode != 1 && A->tile_rows != spmv_rows(ctx, A)) {
    int rc = fsb_mat_setup_tiles(A);      // the tile size option changed since the matrix was set up
    if (rc) return rc;
  }
  SpmvArgs a;
  fsb_spmv_fill_args(A, &a);
  a.x = x; a.y = y; a.w = w; a.want_yy = want_yy; a.w2 = w2; a.out = out; a.done = done;
  if (dd) {
    if (!fsb_spmv_supports_p2p(A)) FSB_FAIL(ctx, FSB_ERR_STATE, "peer-memory SpMV needs the staged kernel (spmv_mode 0)");
    a.pc = dd->pc; a.halo_seq = dd->halo_seq; a.mail_slot = dd->mail_slot; a.mail_seq = dd->mail_seq;
  }
  const bool tiled = ctx->spmv_mode != 1 && A->ntiles > 0;
  if (tiled && ctx->spmv_mode == 0) {
    const int lpr = spmv_lpr(ctx, A);
    const int rows = A->tile_rows;
    const int nst = std::max(2, std::min(spmv_stages(ctx, A), (int)((224 * 1024) / A->stage_bytes)));
    const size_t smem = (size_t)nst * A->stage_bytes;
    bool launched = false;
#define FSB_SPMV_CASE(BS, ROWS, LPR, NST, FLAT)                                                                        \
  if (!launched && A->bs == BS && rows == ROWS && lpr == LPR && nst == NST && (a.flat != 0) == FLAT) {                 \
    using Cfg = SpmvCfg<BS, ROWS, LPR>;                                                                                \
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(2048 / Cfg::THREADS, (227 * 1024) / (smem + 1024)));  \
    const unsigned grid = (unsigned)std::min<int64_t>(A->ntiles, (int64_t)ctx->sm_count * per_sm);                     \
    static bool attr_set = false;                                                                                      \
    if (!attr_set) {                                                                                                   \
      FSB_CHECK_CUDA(ctx, cudaFuncSetAttribute(k_spmv_ws<BS, ROWS, LPR, NST, FLAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024)); \
      attr_set = true;                                                                                                 \
    }                                                                                                                  \
    k_spmv_ws<BS, ROWS, LPR, NST, FLAT><<<grid, Cfg::THREADS, smem, ctx->stream>>>(a);                                 \
    launched = true;                                                                                                   \
  }
#define FSB_SPMV_NST(BS, ROWS, LPR) FSB_SPMV_CASE(BS, ROWS, LPR, 2, false) FSB_SPMV_CASE(BS, ROWS, LPR, 3, false) FSB_SPMV_CASE(BS, ROWS, LPR, 4, false)
#define FSB_SPMV_NST_FLAT(BS, ROWS, LPR) FSB_SPMV_CASE(BS, ROWS, LPR, 2, true) FSB_SPMV_CASE(BS, ROWS, LPR, 3, true)
    // the flat (two-phase) consumer exists for the long-row configurations only; asked for elsewhere, the rows form runs
    if (a.flat && !((A->bs == 3 && rows == 96 && (lpr == 2 || lpr == 4)) || (A->bs == 1 && rows == 128 && (lpr == 2 || lpr == 4)) || (A->bs == 1 && rows == 256 && lpr == 2)) )
      a.flat = 0;
    if (a.flat && nst > 3) a.flat = 0;
    FSB_SPMV_NST(1, 512, 1) FSB_SPMV_NST(1, 256, 1) FSB_SPMV_NST(1, 256, 2) FSB_SPMV_NST(1, 128, 1) FSB_SPMV_NST(1, 128, 2) FSB_SPMV_NST(1, 128, 4)
    FSB_SPMV_NST(2, 256, 2) FSB_SPMV_NST(2, 128, 2)
    FSB_SPMV_NST(3, 192, 2) FSB_SPMV_NST(3, 192, 4) FSB_SPMV_NST(3, 96, 2) FSB_SPMV_NST(3, 96, 4) FSB_SPMV_NST(3, 96, 8)
    FSB_SPMV_NST_FLAT(3, 96, 2) FSB_SPMV_NST_FLAT(3, 96, 4) FSB_SPMV_NST_FLAT(1, 128, 2) FSB_SPMV_NST_FLAT(1, 128, 4) FSB_SPMV_NST_FLAT(1, 256, 2)
#undef FSB_SPMV_NST_FLAT
#undef FSB_SPMV_NST
#undef FSB_SPMV_CASE
    if (!launched) FSB_FAIL(ctx, FSB_ERR_ARG, "unsupported spmv_rows/spmv_lpr/spmv_stages combination");
  } else if (tiled) {
    const size_t smem = 2 * (size_t)A->tile_cap * (8 * A->bs * A->bs + 4);
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (220 * 1024) / (smem + 1024)));
    const unsigned grid = (unsigned)std::min<int64_t>(A->ntiles, (int64_t)ctx->sm_count * per_sm);
    static bool attr_set[4] = {false, false, false, false};
#define FSB_SPMV1_LAUNCH(BS, TH)                                                                                      \
  do {                                                                                                                \
    if (!attr_set[BS]) {                                                                                              \
      FSB_CHECK_CUDA(ctx, cudaFuncSetAttribute(k_spmv_tma<BS, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); \
      attr_set[BS] = true;                                                                                            \
    }                                                                                                                 \
    k_spmv_tma<BS, TH><<<grid, TH, smem, ctx->stream>>>(a);                                                           \
  } while (0)
    if (A->bs == 1) FSB_SPMV1_LAUNCH(1, 256);
    else if (A->bs == 2) FSB_SPMV1_LAUNCH(2, 256);
    else FSB_SPMV1_LAUNCH(3, 192);
#undef FSB_SPMV1_LAUNCH
  } else {
    const unsigned grid = fsb_grid((A->own1 - A->own0) * A->bs, 256, (int64_t)ctx->sm_count * 8);
    if (A->bs == 1) k_spmv_plain<1><<<grid, 256, 0, ctx->stream>>>(a);
    else if (A->bs == 2) k_spmv_plain<2><<<grid, 256, 0, ctx->stream>>>(a);
    else k_spmv_plain<3><<<grid, 256, 0, ctx->stream>>>(a);
  }
  FSB_LAUNCH_CHECK(ctx);
  return FSB_OK;
}

extern "C" int fsb_spmv(fsb_mat* A, fsb_vec* x, fsb_vec* y) {
  if (!A || !x || !y) return FSB_ERR_ARG;
  const int64_t n = A->nbrows * A->bs;
  if (x->n != n || y->n != n || x == y) FSB_FAIL(A->ctx, FSB_ERR_ARG, "vector sizes do not match the matrix");
  if (fsb_dist_active(A->ctx)) {
    int rc = fsb_dist_halo_raw(A->ctx, x->d, x->n);
    if (rc) return rc;
  }
  return fsb_launch_spmv(A, x->d, y->d, nullptr, 0, nullptr, nullptr, nullptr);
}
