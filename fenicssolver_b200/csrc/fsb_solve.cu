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

// ------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
// 1-D TMA bulk copy global -> shared, completion signalled on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ------------------------------------------------------------------------------------ reductions
// Combine per-CTA partials: the last CTA to arrive sums partials[k*stride + 0..nblocks) for k < NV in a
// fixed order and stores the NV results to out[0..NV).  `counter` must be 0 on entry and is reset.
template <int NV>
__device__ __forceinline__ void finish_partials(const double (&mine)[NV], double* __restrict__ partials, int stride,
                                                double* __restrict__ out, unsigned* counter, double* sm) {
  __shared__ bool is_last;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) partials[k * stride + blockIdx.x] = mine[k];
    __threadfence();
    unsigned t = atomicAdd(counter, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double s = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) s += __ldcg(partials + k * stride + i);
    s = block_sum(s, sm);
    if (threadIdx.x == 0) out[k] = s;
  }
  if (threadIdx.x == 0) *counter = 0;
}

// ------------------------------------------------------------------------------------ SpMV
struct SpmvArgs {
  const int64_t* row_ptr;
  const int32_t* col_idx;
  const double* vals;
  const int64_t* tile_row;
  int64_t ntiles;
  int64_t own0, own1;       // owned block rows
  int cap;                  // stage capacity in blocks
  const double* x;
  double* y;
  const double* w;          // optional: d0 = sum y.w
  int want_yy;              // d1 = sum y.y
  double* partials;
  double* out;              // out[0]=d0, out[1]=d1
  unsigned* counter;
  const int* done;          // optional early-exit flag
};

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

  double d0 = 0.0, d1 = 0.0;
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
      }
    }
    __syncthreads();   // stage s is free for the prefetch issued at the top of the next iteration
  }
  if (a.out) {
    double mine[2];
    mine[0] = block_sum(d0, red);
    mine[1] = block_sum(d1, red);
    finish_partials<2>(mine, a.partials, kMaxPartials, a.out, a.counter, red);
  }
}

// v2 of the staged kernel.  LPR lanes share one scalar row (shorter dependent gather chain, more warps
// per SM); the tile's slice of row_ptr arrives by TMA with the values/columns and the tile bounds are
// published through shared memory by the issuing thread, so the compute phase has no global-memory
// dependency other than the x gathers (issued UNR at a time); NST stages keep NST-1 tiles in flight per
// CTA, which is what covers the loaded HBM latency (ncu: long-scoreboard + barrier stalls, r1 profile).
template <int BS, int ROWS, int LPR>
struct SpmvV2 {
  static constexpr int THREADS = ROWS * LPR;           // ROWS scalar rows per pass
  static constexpr int RCAP = 2 * (ROWS / BS) + 8;     // block rows whose row_ptr slice fits the stage
  static constexpr int UNR = BS == 1 ? 8 : (BS == 2 ? 4 : 3);   // gathers in flight per lane (x BS)
};

template <int BS, int ROWS, int LPR, int NST>
__global__ void __launch_bounds__(SpmvV2<BS, ROWS, LPR>::THREADS) k_spmv_tma2(SpmvArgs a) {
  using Cfg = SpmvV2<BS, ROWS, LPR>;
  constexpr int THREADS = Cfg::THREADS, RCAP = Cfg::RCAP, UNR = Cfg::UNR;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar[NST];
  __shared__ int64_t s_info[NST][4];    // per stage: r0, r1, aligned first nnz, aligned first row (or -1: row_ptr not staged)
  __shared__ double red[32];
  if (a.done && *a.done) return;
  constexpr int VB = 8 * BS * BS;
  const size_t rp_bytes = (size_t)(RCAP + 4) * 8;
  const size_t stage_bytes = (size_t)a.cap * (VB + 4) + rp_bytes;
  auto vals_s = [&](int s) { return reinterpret_cast<const double*>(smem + s * stage_bytes); };
  auto cols_s = [&](int s) { return reinterpret_cast<const int32_t*>(smem + s * stage_bytes + (size_t)a.cap * VB); };
  auto rptr_s = [&](int s) { return reinterpret_cast<const int64_t*>(smem + s * stage_bytes + (size_t)a.cap * (VB + 4)); };

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < NST; ++s) mbar_init(&bar[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  auto issue = [&](int64_t tile, int s) {
    const int64_t r0 = a.tile_row[tile], r1 = a.tile_row[tile + 1];
    s_info[s][0] = r0; s_info[s][1] = r1;
    if (r1 <= r0) { mbar_arrive(&bar[s]); return; }
    const int64_t k0 = a.row_ptr[r0], k1 = a.row_ptr[r1];
    const int64_t al0 = k0 & ~3ll;
    const uint32_t cnt = (uint32_t)(((k1 - al0) + 3) & ~3ll);
    const int64_t ra0 = r0 & ~1ll;
    const bool stage_rp = (r1 - ra0 + 1) <= RCAP;
    const uint32_t nrp = stage_rp ? (uint32_t)(((r1 - ra0 + 1) + 1) & ~1ll) : 0u;
    s_info[s][2] = al0; s_info[s][3] = stage_rp ? ra0 : -1;
    mbar_expect_tx(&bar[s], cnt * (VB + 4) + nrp * 8);
    bulk_g2s((void*)vals_s(s), a.vals + al0 * BS * BS, cnt * VB, &bar[s]);
    bulk_g2s((void*)cols_s(s), a.col_idx + al0, cnt * 4, &bar[s]);
    if (stage_rp) bulk_g2s((void*)rptr_s(s), a.row_ptr + ra0, nrp * 8, &bar[s]);
  };

  const int sub = threadIdx.x % LPR;
  double d0 = 0.0, d1 = 0.0;
  int64_t tile = blockIdx.x;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int d = 0; d < NST - 1; ++d)
      if (tile + (int64_t)d * gridDim.x < a.ntiles) issue(tile + (int64_t)d * gridDim.x, d);
  }
  for (int it = 0; tile < a.ntiles; tile += gridDim.x, ++it) {
    const int s = it % NST;
    const int64_t ahead = tile + (int64_t)(NST - 1) * gridDim.x;
    // stage (it+NST-1)%NST was consumed in iteration it-1, which ended with __syncthreads
    if (threadIdx.x == 0 && ahead < a.ntiles) issue(ahead, (it + NST - 1) % NST);
    mbar_wait(&bar[s], (it / NST) & 1);
    const int64_t r0 = s_info[s][0], r1 = s_info[s][1];
    if (r1 > r0) {
      const int64_t al0 = s_info[s][2], ra0 = s_info[s][3];
      const double* __restrict__ vs = vals_s(s);
      const int32_t* __restrict__ cs = cols_s(s);
      const int64_t* __restrict__ rp = rptr_s(s);
      const int nscalar = (int)(r1 - r0) * BS;
      for (int base = 0; base < nscalar; base += THREADS / LPR) {       // warp-uniform trip count
        const int lr = base + threadIdx.x / LPR;
        const bool live = lr < nscalar;
        const int64_t R = r0 + (live ? lr / BS : 0);
        const int i = live ? lr % BS : 0;
        int ks, ke;
        if (ra0 >= 0) { ks = (int)(rp[R - ra0] - al0); ke = (int)(rp[R + 1 - ra0] - al0); }
        else { ks = (int)(a.row_ptr[R] - al0); ke = (int)(a.row_ptr[R + 1] - al0); }
        if (!live) ke = ks;
        double acc = 0.0;
        for (int k = ks + sub; k < ke; k += LPR * UNR) {
          double v[UNR][BS], xg[UNR][BS];
#pragma unroll
          for (int u = 0; u < UNR; ++u) {
            const int kk = k + u * LPR;
            const bool ok = kk < ke;
            const int64_t c = ok ? cs[kk] : 0;
#pragma unroll
            for (int j = 0; j < BS; ++j) {
              v[u][j] = ok ? vs[(kk * BS + i) * BS + j] : 0.0;
              xg[u][j] = ok ? __ldg(a.x + c * BS + j) : 0.0;
            }
          }
#pragma unroll
          for (int u = 0; u < UNR; ++u)
#pragma unroll
            for (int j = 0; j < BS; ++j) acc += v[u][j] * xg[u][j];
        }
#pragma unroll
        for (int o = LPR >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (live && sub == 0) {
          const int64_t row = R * BS + i;
          a.y[row] = acc;
          if (a.w) d0 += acc * a.w[row];
          if (a.want_yy) d1 += acc * acc;
        }
      }
    }
    __syncthreads();
  }
  if (a.out) {
    double mine[2];
    mine[0] = block_sum(d0, red);
    mine[1] = block_sum(d1, red);
    finish_partials<2>(mine, a.partials, kMaxPartials, a.out, a.counter, red);
  }
}

// plain fallback: one thread per scalar row straight from global memory
template <int BS>
__global__ void __launch_bounds__(256) k_spmv_plain(SpmvArgs a) {
  __shared__ double red[32];
  if (a.done && *a.done) return;
  double d0 = 0.0, d1 = 0.0;
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
  }
  if (a.out) {
    double mine[2];
    mine[0] = block_sum(d0, red);
    mine[1] = block_sum(d1, red);
    finish_partials<2>(mine, a.partials, kMaxPartials, a.out, a.counter, red);
  }
}

__global__ void k_tile_rows(const int64_t* __restrict__ row_ptr, int64_t own0, int64_t own1, int64_t tile_nnz,
                            int64_t ntiles, int64_t* __restrict__ tile_row) {
  const int64_t base = row_ptr[own0];
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t <= ntiles; t += (int64_t)gridDim.x * blockDim.x) {
    if (t == ntiles) { tile_row[t] = own1; continue; }
    const int64_t target = base + t * tile_nnz;
    int64_t lo = own0, hi = own1;          // first row r in [own0, own1] with row_ptr[r] >= target
    while (lo < hi) {
      int64_t mid = (lo + hi) >> 1;
      if (row_ptr[mid] < target) lo = mid + 1; else hi = mid;
    }
    tile_row[t] = lo;
  }
}

// scalar rows per tile: 256 (192 for 3x3 blocks) or half of that (ctx option spmv_rows)
static int spmv_rows(fsb_ctx* ctx, int bs) { const int big = bs == 3 ? 192 : 256; return ctx->spmv_rows == 128 ? big / 2 : big; }
static constexpr size_t kSmemBudget = 200 * 1024;

int fsb_mat_setup_tiles(fsb_mat* A) {
  fsb_ctx* ctx = A->ctx;
  cudaFree(A->tile_row);
  A->tile_row = nullptr;
  A->ntiles = 0; A->tile_nnz = 0; A->tile_cap = 0;
  const int64_t nrows = A->own1 - A->own0;
  if (nrows <= 0) return FSB_OK;
  int64_t k01[2];
  FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(&k01[0], A->row_ptr + A->own0, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(&k01[1], A->row_ptr + A->own1, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const int64_t nnz = k01[1] - k01[0];
  if (nnz <= 0) return FSB_OK;
  const int bs = A->bs;
  const int rows_target = spmv_rows(ctx, bs) / bs;
  A->tile_rows = spmv_rows(ctx, bs);
  const double avg = (double)nnz / (double)nrows;
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
  int rc = fsb_dmalloc(ctx, &A->tile_row, (size_t)A->ntiles + 1);
  if (rc) return rc;
  k_tile_rows<<<fsb_grid(A->ntiles + 1, 256, 4096), 256, 0, ctx->stream>>>(A->row_ptr, A->own0, A->own1, T, A->ntiles, A->tile_row);
  FSB_LAUNCH_CHECK(ctx);
  return FSB_OK;
}

// launches y = A x (+ fused dots) on the ctx stream
static int launch_spmv(fsb_mat* A, const double* x, double* y, const double* w, int want_yy, double* out,
                       const int* done) {
  fsb_ctx* ctx = A->ctx;
  SpmvArgs a;
  a.row_ptr = A->row_ptr; a.col_idx = A->col_idx; a.vals = A->vals; a.tile_row = A->tile_row;
  a.ntiles = A->ntiles; a.own0 = A->own0; a.own1 = A->own1; a.cap = A->tile_cap;
  a.x = x; a.y = y; a.w = w; a.want_yy = want_yy;
  a.partials = ctx->d_partials; a.out = out; a.counter = ctx->d_counters + 0; a.done = done;
  if (A->own1 <= A->own0) return FSB_OK;
  if (ctx->spmv_mode != 1 && A->tile_rows != spmv_rows(ctx, A->bs)) {
    int rc = fsb_mat_setup_tiles(A);      // the tile size option changed since the matrix was set up
    if (rc) return rc;
    a.tile_row = A->tile_row; a.ntiles = A->ntiles; a.cap = A->tile_cap;
  }
  const bool tiled = ctx->spmv_mode != 1 && A->ntiles > 0;
  if (tiled && ctx->spmv_mode == 0) {
    // v2: ROWS x LPR threads, NST stages (clamped to what fits in shared memory)
    const int lpr = ctx->spmv_lpr;
    const int rows = A->tile_rows;
    int nst = std::max(2, std::min(ctx->spmv_stages, (int)((224 * 1024) / A->stage_bytes)));
    const size_t smem = (size_t)nst * A->stage_bytes;
    bool launched = false;
#define FSB_SPMV2_CASE(BS, ROWS, LPR, NST)                                                                             \
  if (!launched && A->bs == BS && rows == ROWS && lpr == LPR && nst == NST) {                                          \
    using Cfg = SpmvV2<BS, ROWS, LPR>;                                                                                 \
    int per_sm = (int)std::max<size_t>(1, std::min<size_t>(2048 / Cfg::THREADS, (227 * 1024) / (smem + 1024)));        \
    const unsigned grid = (unsigned)std::min<int64_t>(A->ntiles, (int64_t)ctx->sm_count * per_sm);                     \
    static bool attr_set = false;                                                                                      \
    if (!attr_set) {                                                                                                   \
      FSB_CHECK_CUDA(ctx, cudaFuncSetAttribute(k_spmv_tma2<BS, ROWS, LPR, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024)); \
      attr_set = true;                                                                                                 \
    }                                                                                                                  \
    k_spmv_tma2<BS, ROWS, LPR, NST><<<grid, Cfg::THREADS, smem, ctx->stream>>>(a);                                     \
    launched = true;                                                                                                   \
  }
#define FSB_SPMV2_NST(BS, ROWS, LPR) FSB_SPMV2_CASE(BS, ROWS, LPR, 2) FSB_SPMV2_CASE(BS, ROWS, LPR, 3) FSB_SPMV2_CASE(BS, ROWS, LPR, 4)
    FSB_SPMV2_NST(1, 256, 2) FSB_SPMV2_NST(1, 256, 4) FSB_SPMV2_NST(1, 128, 2) FSB_SPMV2_NST(1, 128, 4)
    FSB_SPMV2_NST(2, 256, 2) FSB_SPMV2_NST(2, 128, 2)
    FSB_SPMV2_NST(3, 192, 2) FSB_SPMV2_NST(3, 192, 4) FSB_SPMV2_NST(3, 96, 2) FSB_SPMV2_NST(3, 96, 4)
#undef FSB_SPMV2_NST
#undef FSB_SPMV2_CASE
    if (!launched) FSB_FAIL(ctx, FSB_ERR_ARG, "unsupported spmv_rows/spmv_lpr/spmv_stages combination");
  } else if (tiled) {
    const size_t smem = 2 * (size_t)A->tile_cap * (8 * A->bs * A->bs + 4);
    int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (220 * 1024) / (smem + 1024)));
    const unsigned grid = (unsigned)std::min<int64_t>(A->ntiles, (int64_t)ctx->sm_count * per_sm);
    static bool attr_set[4] = {false, false, false, false};
#define FSB_SPMV_LAUNCH(BS, TH)                                                                                       \
  do {                                                                                                                \
    if (!attr_set[BS]) {                                                                                              \
      FSB_CHECK_CUDA(ctx, cudaFuncSetAttribute(k_spmv_tma<BS, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); \
      attr_set[BS] = true;                                                                                            \
    }                                                                                                                 \
    k_spmv_tma<BS, TH><<<grid, TH, smem, ctx->stream>>>(a);                                                           \
  } while (0)
    if (A->bs == 1) FSB_SPMV_LAUNCH(1, 256);
    else if (A->bs == 2) FSB_SPMV_LAUNCH(2, 256);
    else FSB_SPMV_LAUNCH(3, 192);
#undef FSB_SPMV_LAUNCH
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
  return launch_spmv(A, x->d, y->d, nullptr, 0, nullptr, nullptr);
}

// ------------------------------------------------------------------------------------ vector kernels
static constexpr int kVecThreads = 256;
static unsigned vec_grid(fsb_ctx* ctx, int64_t n) { return fsb_grid(n, kVecThreads * 4, std::min<int64_t>(kMaxPartials, (int64_t)ctx->sm_count * 8)); }

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
       S_RHO0 = 8, S_RRB0 = 9, S_BBB = 10, S_RHO1 = 11, S_RRB1 = 12, S_RV = 13, S_TS = 14, S_TT = 15, S_SPARE = 16 };
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

__global__ void k_check0(const double* __restrict__ scal, int rr_slot, int bb_slot, double rtol, double atol, int maxit, int* state, double* final_rr) {
  const double rr = scal[rr_slot], bb = scal[bb_slot];
  const double tol2 = fmax(rtol * rtol * bb, atol * atol);
  state[1] = 0;
  *final_rr = rr;
  if (rr <= tol2) { state[0] = 1; state[2] = 1; }
  else if (!(rr == rr)) { state[0] = 1; state[2] = -1; }
  else if (maxit <= 0) { state[0] = 1; state[2] = 0; }
  else { state[0] = 0; state[2] = 0; }
}

// alpha = rz/pq ; x += alpha p ; r -= alpha q ; z = dinv r ; sums rz', rr' -> out[0..2)
__global__ void __launch_bounds__(kVecThreads)
k_cg_update(int64_t n0, int64_t n1, const double* __restrict__ scal, int rz_slot, int pq_slot, const double* __restrict__ p,
            const double* __restrict__ q, const double* __restrict__ dinv, double* __restrict__ x, double* __restrict__ r,
            double* partials, double* out, unsigned* counter, const int* done) {
  __shared__ double red[32];
  if (*done) return;
  const double alpha = scal[rz_slot] / scal[pq_slot];
  double s0 = 0, s1 = 0;
  for (int64_t i = n0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n1; i += (int64_t)gridDim.x * blockDim.x) {
    x[i] += alpha * p[i];
    const double ri = r[i] - alpha * q[i];
    r[i] = ri;
    const double zi = dinv[i] * ri;
    s0 += ri * zi;
    s1 += zi * zi;
  }
  double mine[2] = {block_sum(s0, red), block_sum(s1, red)};
  finish_partials<2>(mine, partials, kMaxPartials, out, counter, red);
}

// beta = rz'/rz ; p = dinv r + beta p ; block 0 / thread 0 advances the iteration state
__global__ void __launch_bounds__(kVecThreads)
k_cg_pupdate(int64_t n0, int64_t n1, double* __restrict__ scal, int rz_old, int rz_new, int rr_new, int pq_slot,
             const double* __restrict__ r, const double* __restrict__ dinv, double* __restrict__ p, double rtol, double atol,
             int maxit, int* state) {
  if (state[0]) return;
  const double rzn = scal[rz_new], rzo = scal[rz_old];
  const double beta = rzn / rzo;
  for (int64_t i = n0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n1; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = dinv[i] * r[i] + beta * p[i];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const double rr = scal[rr_new], bb = scal[S_BB], pq = scal[pq_slot];
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
struct Workspace {
  fsb_ctx* ctx;
  std::vector<double*> bufs;
  int alloc(double** p, int64_t n) {
    int rc = fsb_dmalloc(ctx, p, (size_t)n);
    if (!rc) { bufs.push_back(*p); cudaMemsetAsync(*p, 0, sizeof(double) * n, ctx->stream); }
    return rc;
  }
  ~Workspace() { for (double* b : bufs) cudaFree(b); }
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

extern "C" int fsb_solve_cg(fsb_mat* A, fsb_vec* b, fsb_vec* x, double rtol, double atol, int32_t maxit, int32_t precond,
                            fsb_solve_info* info) {
  if (!A || !b || !x || !info) return FSB_ERR_ARG;
  fsb_ctx* ctx = A->ctx;
  const int64_t n = A->nbrows * A->bs;
  if (b->n != n || x->n != n) FSB_FAIL(ctx, FSB_ERR_ARG, "vector sizes do not match the matrix");
  memset(info, 0, sizeof(*info));
  const int64_t n0 = A->own0 * A->bs, n1 = A->own1 * A->bs;
  const bool dist = fsb_dist_active(ctx);
  Workspace ws{ctx};
  double *r, *p, *q, *dinv;
  int rc;
  if ((rc = ws.alloc(&r, n)) || (rc = ws.alloc(&p, n)) || (rc = ws.alloc(&q, n)) || (rc = ws.alloc(&dinv, n))) return rc;
  double* scal = ctx->d_scalars;
  int* state = ctx->d_state;
  const unsigned vg = vec_grid(ctx, n1 - n0);
  cudaEvent_t e0, e1;
  FSB_CHECK_CUDA(ctx, cudaEventCreate(&e0));
  FSB_CHECK_CUDA(ctx, cudaEventCreate(&e1));
  struct EvGuard { cudaEvent_t a, b; ~EvGuard() { cudaEventDestroy(a); cudaEventDestroy(b); } } guard{e0, e1};
  SpmvTimer timer;
  FSB_CHECK_CUDA(ctx, cudaEventRecord(e0, ctx->stream));

  FSB_CHECK_CUDA(ctx, cudaMemsetAsync(state, 0, sizeof(int) * 8, ctx->stream));
#define DINV_LAUNCH(BS) k_extract_dinv<BS><<<fsb_grid(n1 - n0, 256, (int64_t)ctx->sm_count * 16), 256, 0, ctx->stream>>>(n0, n1, A->row_ptr, A->col_idx, A->vals, dinv, precond == 1)
  if (A->bs == 1) DINV_LAUNCH(1); else if (A->bs == 2) DINV_LAUNCH(2); else DINV_LAUNCH(3);
#undef DINV_LAUNCH
  FSB_LAUNCH_CHECK(ctx);
  // r0 = b - A x0
  if (dist && (rc = fsb_dist_halo_raw(ctx, x->d, n))) return rc;
  if ((rc = launch_spmv(A, x->d, q, nullptr, 0, nullptr, nullptr))) return rc;
  k_cg_init<<<vg, kVecThreads, 0, ctx->stream>>>(n0, n1, b->d, q, dinv, r, p, ctx->d_partials, scal + S_RZ0, ctx->d_counters + 1);
  FSB_LAUNCH_CHECK(ctx);
  if (dist && (rc = fsb_dist_allreduce_sum_dev(ctx, scal + S_RZ0, 3))) return rc;
  k_check0<<<1, 1, 0, ctx->stream>>>(scal, S_RR0, S_BB, rtol, atol, maxit, state, scal + S_FINAL_RR);
  FSB_LAUNCH_CHECK(ctx);

  // iteration batches; the host polls the state one batch behind the launches
  const int batch = std::max(2, ctx->check_every & ~1);
  cudaEvent_t polled[2];
  FSB_CHECK_CUDA(ctx, cudaEventCreateWithFlags(&polled[0], cudaEventDisableTiming));
  FSB_CHECK_CUDA(ctx, cudaEventCreateWithFlags(&polled[1], cudaEventDisableTiming));
  struct PollGuard { cudaEvent_t* e; ~PollGuard() { cudaEventDestroy(e[0]); cudaEventDestroy(e[1]); } } pguard{polled};
  if (ctx->profile) timer.ensure(batch);
  int launched = 0;
  for (int nb = 0;; ++nb) {
    const int slot = nb & 1;
    const bool more = launched < maxit;
    if (more) {
      for (int k = 0; k < batch; ++k) {
        const int par = (launched + k) & 1;
        const int pq = par ? S_PQ1 : S_PQ0, rz = par ? S_RZ1 : S_RZ0, rzn = par ? S_RZ0 : S_RZ1, rrn = par ? S_RR0 : S_RR1;
        if (dist && (rc = fsb_dist_halo_raw(ctx, p, n))) return rc;
        if (ctx->profile) cudaEventRecord(timer.next(slot), ctx->stream);
        if ((rc = launch_spmv(A, p, q, p, 0, scal + pq, state))) return rc;
        if (ctx->profile) cudaEventRecord(timer.next(slot), ctx->stream);
        if (dist && (rc = fsb_dist_allreduce_sum_dev(ctx, scal + pq, 1))) return rc;
        k_cg_update<<<vg, kVecThreads, 0, ctx->stream>>>(n0, n1, scal, rz, pq, p, q, dinv, x->d, r, ctx->d_partials, scal + rzn, ctx->d_counters + 2, state);
        FSB_LAUNCH_CHECK(ctx);
        if (dist && (rc = fsb_dist_allreduce_sum_dev(ctx, scal + rzn, 2))) return rc;
        k_cg_pupdate<<<vg, kVecThreads, 0, ctx->stream>>>(n0, n1, scal, rz, rzn, rrn, pq, r, dinv, p, rtol, atol, maxit, state);
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
// r = b - q ; rhat = r ; sums rho=rhat.r, |dinv r|^2, |dinv b|^2 -> out[0..3)
__global__ void __launch_bounds__(kVecThreads)
k_bcg_init(int64_t n0, int64_t n1, const double* __restrict__ b, const double* __restrict__ q, const double* __restrict__ dinv,
           double* __restrict__ r, double* __restrict__ rhat, double* partials, double* out, unsigned* counter) {
  __shared__ double red[32];
  double s0 = 0, s1 = 0, s2 = 0;
  for (int64_t i = n0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n1; i += (int64_t)gridDim.x * blockDim.x) {
    const double bi = b[i], ri = bi - q[i];
    r[i] = ri; rhat[i] = ri;
    const double zi = dinv[i] * ri, zb = dinv[i] * bi;
    s0 += ri * ri; s1 += zi * zi; s2 += zb * zb;
  }
  double mine[3] = {block_sum(s0, red), block_sum(s1, red), block_sum(s2, red)};
  finish_partials<3>(mine, partials, kMaxPartials, out, counter, red);
}

// p = r + beta (p - omega v), beta = (rho/rho_prev)(alpha/omega) ; ph = dinv p      (first: p = r)
__global__ void __launch_bounds__(kVecThreads)
k_bcg_p(int64_t n0, int64_t n1, const double* __restrict__ scal, int rho_cur, int rho_prev, int first,
        const double* __restrict__ r, const double* __restrict__ v, const double* __restrict__ dinv, double* __restrict__ p,
        double* __restrict__ ph, const int* done) {
  if (*done) return;
  double beta = 0.0, omega = 0.0;
  if (!first) {
    const double alpha = scal[rho_prev] / scal[S_RV];
    omega = scal[S_TS] / scal[S_TT];
    beta = (scal[rho_cur] / scal[rho_prev]) * (alpha / omega);
  }
  for (int64_t i = n0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n1; i += (int64_t)gridDim.x * blockDim.x) {
    const double pi = first ? r[i] : r[i] + beta * (p[i] - omega * v[i]);
    p[i] = pi;
    ph[i] = dinv[i] * pi;
  }
}

// alpha = rho/(rhat.v) ; s = r - alpha v (stored in r) ; sh = dinv s
__global__ void __launch_bounds__(kVecThreads)
k_bcg_s(int64_t n0, int64_t n1, const double* __restrict__ scal, int rho_cur, const double* __restrict__ v,
        const double* __restrict__ dinv, double* __restrict__ r, double* __restrict__ sh, const int* done) {
  if (*done) return;
  const double alpha = scal[rho_cur] / scal[S_RV];
  for (int64_t i = n0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n1; i += (int64_t)gridDim.x * blockDim.x) {
    const double si = r[i] - alpha * v[i];
    r[i] = si;
    sh[i] = dinv[i] * si;
  }
}

// omega = (t.s)/(t.t) ; x += alpha ph + omega sh ; r = s - omega t ; sums rho'=rhat.r, rr -> out[0..2)
__global__ void __launch_bounds__(kVecThreads)
k_bcg_x(int64_t n0, int64_t n1, const double* __restrict__ scal, int rho_cur, const double* __restrict__ ph,
        const double* __restrict__ sh, const double* __restrict__ t, const double* __restrict__ rhat,
        const double* __restrict__ dinv, double* __restrict__ x, double* __restrict__ r, double* partials, double* out, unsigned* counter, const int* done) {
  __shared__ double red[32];
  if (*done) return;
  const double alpha = scal[rho_cur] / scal[S_RV];
  const double omega = scal[S_TS] / scal[S_TT];
  double s0 = 0, s1 = 0;
  for (int64_t i = n0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n1; i += (int64_t)gridDim.x * blockDim.x) {
    x[i] += alpha * ph[i] + omega * sh[i];
    const double ri = r[i] - omega * t[i];
    r[i] = ri;
    const double zi = dinv[i] * ri;
    s0 += rhat[i] * ri;
    s1 += zi * zi;
  }
  double mine[2] = {block_sum(s0, red), block_sum(s1, red)};
  finish_partials<2>(mine, partials, kMaxPartials, out, counter, red);
}

__global__ void k_bcg_check(double* __restrict__ scal, int rho_new, int rr_new, double rtol, double atol, int maxit, int* state) {
  if (state[0]) return;
  const double rr = scal[rr_new], bb = scal[S_BBB], rho = scal[rho_new], tt = scal[S_TT], rv = scal[S_RV];
  const double tol2 = fmax(rtol * rtol * bb, atol * atol);
  const int it = state[1] + 1;
  state[1] = it;
  scal[S_FINAL_RR] = rr;
  if (rr <= tol2) { state[2] = 1; state[0] = 1; }
  else if (!(rr == rr) || rho == 0.0 || tt == 0.0 || rv == 0.0 || !(rho == rho)) { state[2] = -1; state[0] = 1; }
  else if (it >= maxit) { state[2] = 0; state[0] = 1; }
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
  Workspace ws{ctx};
  double *r, *rhat, *p, *ph, *v, *sh, *t, *dinv;
  int rc;
  if ((rc = ws.alloc(&r, n)) || (rc = ws.alloc(&rhat, n)) || (rc = ws.alloc(&p, n)) || (rc = ws.alloc(&ph, n)) ||
      (rc = ws.alloc(&v, n)) || (rc = ws.alloc(&sh, n)) || (rc = ws.alloc(&t, n)) || (rc = ws.alloc(&dinv, n)))
    return rc;
  double* scal = ctx->d_scalars;
  int* state = ctx->d_state;
  const unsigned vg = vec_grid(ctx, n1 - n0);
  cudaEvent_t e0, e1;
  FSB_CHECK_CUDA(ctx, cudaEventCreate(&e0));
  FSB_CHECK_CUDA(ctx, cudaEventCreate(&e1));
  struct EvGuard { cudaEvent_t a, b; ~EvGuard() { cudaEventDestroy(a); cudaEventDestroy(b); } } guard{e0, e1};
  FSB_CHECK_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
  FSB_CHECK_CUDA(ctx, cudaMemsetAsync(state, 0, sizeof(int) * 8, ctx->stream));
#define DINV_LAUNCH(BS) k_extract_dinv<BS><<<fsb_grid(n1 - n0, 256, (int64_t)ctx->sm_count * 16), 256, 0, ctx->stream>>>(n0, n1, A->row_ptr, A->col_idx, A->vals, dinv, precond == 1)
  if (A->bs == 1) DINV_LAUNCH(1); else if (A->bs == 2) DINV_LAUNCH(2); else DINV_LAUNCH(3);
#undef DINV_LAUNCH
  FSB_LAUNCH_CHECK(ctx);
  if (dist && (rc = fsb_dist_halo_raw(ctx, x->d, n))) return rc;
  if ((rc = launch_spmv(A, x->d, v, nullptr, 0, nullptr, nullptr))) return rc;
  k_bcg_init<<<vg, kVecThreads, 0, ctx->stream>>>(n0, n1, b->d, v, dinv, r, rhat, ctx->d_partials, scal + S_RHO0, ctx->d_counters + 1);
  FSB_LAUNCH_CHECK(ctx);
  if (dist && (rc = fsb_dist_allreduce_sum_dev(ctx, scal + S_RHO0, 3))) return rc;
  k_check0<<<1, 1, 0, ctx->stream>>>(scal, S_RRB0, S_BBB, rtol, atol, maxit, state, scal + S_FINAL_RR);
  FSB_LAUNCH_CHECK(ctx);

  const int batch = std::max(2, ctx->check_every & ~1);
  int launched = 0;
  bool finished = false;
  while (!finished) {
    for (int k = 0; k < batch; ++k) {
      const int par = (launched + k) & 1;
      const int rho = par ? S_RHO1 : S_RHO0, rhon = par ? S_RHO0 : S_RHO1, rrn = par ? S_RRB0 : S_RRB1;
      k_bcg_p<<<vg, kVecThreads, 0, ctx->stream>>>(n0, n1, scal, rho, rhon, (launched + k) == 0, r, v, dinv, p, ph, state);
      FSB_LAUNCH_CHECK(ctx);
      if (dist && (rc = fsb_dist_halo_raw(ctx, ph, n))) return rc;
      if ((rc = launch_spmv(A, ph, v, rhat, 0, scal + S_RV, state))) return rc;
      if (dist && (rc = fsb_dist_allreduce_sum_dev(ctx, scal + S_RV, 1))) return rc;
      k_bcg_s<<<vg, kVecThreads, 0, ctx->stream>>>(n0, n1, scal, rho, v, dinv, r, sh, state);
      FSB_LAUNCH_CHECK(ctx);
      if (dist && (rc = fsb_dist_halo_raw(ctx, sh, n))) return rc;
      if ((rc = launch_spmv(A, sh, t, r, 1, scal + S_TS, state))) return rc;
      if (dist && (rc = fsb_dist_allreduce_sum_dev(ctx, scal + S_TS, 2))) return rc;
      k_bcg_x<<<vg, kVecThreads, 0, ctx->stream>>>(n0, n1, scal, rho, ph, sh, t, rhat, dinv, x->d, r, ctx->d_partials, scal + rhon, ctx->d_counters + 2, state);
      FSB_LAUNCH_CHECK(ctx);
      if (dist && (rc = fsb_dist_allreduce_sum_dev(ctx, scal + rhon, 2))) return rc;
      k_bcg_check<<<1, 1, 0, ctx->stream>>>(scal, rhon, rrn, rtol, atol, maxit, state);
      FSB_LAUNCH_CHECK(ctx);
    }
    launched += batch;
    FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(ctx->h_state, state, sizeof(int) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->h_state[0] || launched >= maxit) finished = true;
  }
  FSB_CHECK_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
  rc = read_outcome(ctx, S_FINAL_RR, S_BBB, info);
  if (rc) return rc;
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  info->solve_ms = ms;
  if (info->converged < 0) FSB_FAIL(ctx, FSB_ERR_BREAKDOWN, "BiCGStab breakdown (zero or non-finite recurrence scalar)");
  return FSB_OK;
}
