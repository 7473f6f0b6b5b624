// The tile pipeline of the block-CSR SpMV, shared by the stand-alone kernel (fsb_spmv.cu: k_spmv_ws) and the
// persistent CG kernel (fsb_cgp.cu: k_cg_persist), which runs the same producer/consumer pipeline once per
// iteration without leaving the SMs.
//   * spmv_issue_tile: one lane of the producer warp starts the 1-D TMA bulk copies (values, columns, row_ptr slice)
//     of a tile into a pipeline stage;
//   * spmv_consume_tile: the consumer warps compute the rows of a staged tile (LPR lanes per scalar row, UNR gathers
//     in flight per lane) and accumulate the fused dot products.
#pragma once
#include "fsb_device.cuh"

struct SpmvArgs {
  const int64_t* row_ptr;
  const int32_t* col_idx;
  const double* vals;
  const int64_t* tile_row;  // [ntiles+1] first block row per tile
  const int64_t* tile_k;    // [ntiles+1] row_ptr[tile_row[t]]
  int64_t ntiles;
  int64_t own0, own1;       // owned block rows
  int cap;                  // stage capacity in blocks
  const double* x;
  double* y;
  const double* w;          // optional: d0 = sum y.w
  int want_yy;              // d1 = sum y.y
  const double* w2;         // optional: d2 = sum y.w2 (then out has 3 entries)
  int l2_hint;              // 1: matrix stream marked evict-first in L2, y written with streaming stores
  int flat;                 // 1: two-phase tiles (products over the tile's non-zeros, then row sums): long, uneven rows
  double* partials;
  double* out;              // out[0]=d0, out[1]=d1
  unsigned* counter;
  const int* done;          // optional early-exit flag
  // distributed CG over peer memory (pc.nranks <= 1: off)
  PeerComm pc;
  unsigned long long halo_seq;   // != 0: wait until the neighbours have delivered their planes of x (flag >= halo_seq)
  int mail_slot;                 // >= 0: post (d0, d1) to every rank's mailbox with mail_seq instead of `out`
  unsigned long long mail_seq;
};

template <int BS, int ROWS, int LPR>
struct SpmvCfg {
  static constexpr int CONSUMERS = ROWS * LPR;         // ROWS scalar rows per pass
  static constexpr int THREADS = CONSUMERS + 32;       // + one producer warp
  static constexpr int RCAP = 2 * (ROWS / BS) + 8;     // block rows whose row_ptr slice fits the stage
  static constexpr int UNR = BS == 1 ? 8 : 4;          // blocks in flight per lane (x BS gathers each)
  static constexpr int VB = 8 * BS * BS;               // bytes of values per block
};

// shared-memory layout of one pipeline stage: [cap blocks of values | cap columns | row_ptr slice]
template <int BS, int ROWS, int LPR>
struct SpmvStage {
  using Cfg = SpmvCfg<BS, ROWS, LPR>;
  unsigned char* base;
  size_t stage_bytes;
  int cap;
  __device__ __forceinline__ SpmvStage(unsigned char* smem, int cap_) : base(smem), cap(cap_) {
    stage_bytes = (size_t)cap_ * (Cfg::VB + 4) + (size_t)(Cfg::RCAP + 4) * 8;
  }
  __device__ __forceinline__ const double* vals(int s) const { return reinterpret_cast<const double*>(base + s * stage_bytes); }
  __device__ __forceinline__ const int32_t* cols(int s) const { return reinterpret_cast<const int32_t*>(base + s * stage_bytes + (size_t)cap * Cfg::VB); }
  __device__ __forceinline__ const int64_t* rptr(int s) const { return reinterpret_cast<const int64_t*>(base + s * stage_bytes + (size_t)cap * (Cfg::VB + 4)); }
};

// producer lane: describe `tile` in info[0..4) = {r0, r1, aligned first nnz, aligned first row or -1} and start its copies;
// an empty tile only arrives on the barrier
template <int BS, int ROWS, int LPR>
__device__ __forceinline__ void spmv_issue_tile(const SpmvArgs& a, const SpmvStage<BS, ROWS, LPR>& st, int64_t tile, int s, int64_t* info,
                                                uint64_t* full, uint64_t policy) {
  using Cfg = SpmvCfg<BS, ROWS, LPR>;
  constexpr int VB = Cfg::VB, RCAP = Cfg::RCAP;
  const int64_t r0 = a.tile_row[tile], r1 = a.tile_row[tile + 1];
  const int64_t k0 = a.tile_k[tile], k1 = a.tile_k[tile + 1];
  info[0] = r0; info[1] = r1;
  if (r1 <= r0) {
    mbar_arrive(full);
    return;
  }
  const int64_t al0 = k0 & ~3ll;
  const uint32_t cnt = (uint32_t)(((k1 - al0) + 3) & ~3ll);
  const int64_t ra0 = r0 & ~1ll;
  const bool stage_rp = (r1 - ra0 + 1) <= RCAP;
  const uint32_t nrp = stage_rp ? (uint32_t)(((r1 - ra0 + 1) + 1) & ~1ll) : 0u;
  info[2] = al0; info[3] = stage_rp ? ra0 : -1;
  mbar_expect_tx(full, cnt * (VB + 4) + nrp * 8);
  if (a.l2_hint) {
    bulk_g2s_hint((void*)st.vals(s), a.vals + al0 * BS * BS, cnt * VB, full, policy);
    bulk_g2s_hint((void*)st.cols(s), a.col_idx + al0, cnt * 4, full, policy);
    if (stage_rp) bulk_g2s_hint((void*)st.rptr(s), a.row_ptr + ra0, nrp * 8, full, policy);
  } else {
    bulk_g2s((void*)st.vals(s), a.vals + al0 * BS * BS, cnt * VB, full);
    bulk_g2s((void*)st.cols(s), a.col_idx + al0, cnt * 4, full);
    if (stage_rp) bulk_g2s((void*)st.rptr(s), a.row_ptr + ra0, nrp * 8, full);
  }
}

// x may be rewritten by other CTAs while a persistent kernel runs (COHERENT): then the gathers are ordinary
// global loads (ordered by the grid barrier's acquire) instead of the non-coherent read-only path
template <bool COHERENT>
__device__ __forceinline__ double spmv_ld(const double* p) {
  if (COHERENT) return *p;
  return __ldg(p);
}

// consumer threads (threadIdx.x < CONSUMERS): the rows of the tile staged in `s`
template <int BS, int ROWS, int LPR, bool COHERENT>
__device__ __forceinline__ void spmv_consume_tile(const SpmvArgs& a, const SpmvStage<BS, ROWS, LPR>& st, int s, const int64_t* info,
                                                  double& d0, double& d1, double& d2) {
  using Cfg = SpmvCfg<BS, ROWS, LPR>;
  constexpr int UNR = Cfg::UNR;
  const int sub = threadIdx.x % LPR;
  const int64_t r0 = info[0], r1 = info[1];
  if (r1 <= r0) return;
  const int64_t al0 = info[2], ra0 = info[3];
  const double* __restrict__ vs = st.vals(s);
  const int32_t* __restrict__ cs = st.cols(s);
  const int64_t* __restrict__ rp = st.rptr(s);
  const int nscalar = (int)(r1 - r0) * BS;
  for (int base = 0; base < nscalar; base += ROWS) {       // warp-uniform trip count
    const int lr = base + threadIdx.x / LPR;
    const bool live = lr < nscalar;
    const int64_t R = r0 + (live ? lr / BS : 0);
    const int i = live ? lr % BS : 0;
    const int64_t row = R * BS + i;
    int ks, ke;
    if (ra0 >= 0) { ks = (int)(rp[R - ra0] - al0); ke = (int)(rp[R + 1 - ra0] - al0); }
    else { ks = (int)(a.row_ptr[R] - al0); ke = (int)(a.row_ptr[R + 1] - al0); }
    if (!live) ke = ks;
    const double wv = (a.w && live && sub == 0) ? spmv_ld<COHERENT>(a.w + row) : 0.0;   // issued with the gathers
    const double wv2 = (a.w2 && live && sub == 0) ? spmv_ld<COHERENT>(a.w2 + row) : 0.0;
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
          xg[u][j] = ok ? spmv_ld<COHERENT>(a.x + c * BS + j) : 0.0;
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
      if (a.l2_hint) __stcs(a.y + row, acc); else a.y[row] = acc;
      d0 += acc * wv;
      if (a.want_yy) d1 += acc * acc;
      d2 += acc * wv2;
    }
  }
}

// Two-phase form of the same tile for long, uneven rows (degree-2 spaces: 28 blocks per row on average, 10 to 90 per row).
// With LPR lanes per row the gathers of a tile are as unbalanced as its rows: a warp waits for its longest row while the lanes
// of short rows idle.  Here phase A walks the tile's NON-ZEROS, not its rows: every consumer thread takes blocks k, k + CONSUMERS,
// ... of the staged tile, gathers x, and overwrites the block's first BS values in shared memory with its BS products-sums —
// perfectly balanced, UNR gathers in flight per thread.  After a consumer barrier phase B adds up each row's entries from shared
// memory (LPR lanes per row, shuffle-combined) and writes y and the fused dots.
template <int BS, int ROWS, int LPR>
__device__ __forceinline__ void spmv_consume_tile_flat(const SpmvArgs& a, const SpmvStage<BS, ROWS, LPR>& st, int s, const int64_t* info,
                                                       double& d0, double& d1, double& d2) {
  using Cfg = SpmvCfg<BS, ROWS, LPR>;
  constexpr int UNR = BS == 1 ? 8 : 2, CONSUMERS = Cfg::CONSUMERS;      // blocks in flight per thread in phase A
  const int64_t r0 = info[0], r1 = info[1];
  if (r1 <= r0) return;                                  // uniform over the consumers: nobody reaches the barrier below
  const int64_t al0 = info[2], ra0 = info[3];
  double* __restrict__ vs = const_cast<double*>(st.vals(s));
  const int32_t* __restrict__ cs = st.cols(s);
  const int64_t* __restrict__ rp = st.rptr(s);
  int kt0, kt1;
  if (ra0 >= 0) { kt0 = (int)(rp[r0 - ra0] - al0); kt1 = (int)(rp[r1 - ra0] - al0); }
  else { kt0 = (int)(a.row_ptr[r0] - al0); kt1 = (int)(a.row_ptr[r1] - al0); }
  // ---- phase A: one product (BS == 1) or one block-times-vector (BS products-sums) per non-zero block, in place
  for (int k = kt0 + (int)threadIdx.x; k < kt1; k += CONSUMERS * UNR) {
    double v[UNR][BS * BS], xg[UNR][BS];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int kk = k + u * CONSUMERS;
      const bool ok = kk < kt1;
      const int64_t c = ok ? cs[kk] : 0;
#pragma unroll
      for (int j = 0; j < BS; ++j) xg[u][j] = ok ? __ldg(a.x + c * BS + j) : 0.0;
#pragma unroll
      for (int e = 0; e < BS * BS; ++e) v[u][e] = ok ? vs[kk * BS * BS + e] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int kk = k + u * CONSUMERS;
      if (kk < kt1) {
#pragma unroll
        for (int i = 0; i < BS; ++i) {
          double t = 0.0;
#pragma unroll
          for (int j = 0; j < BS; ++j) t += v[u][i * BS + j] * xg[u][j];
          vs[kk * BS * BS + i] = t;
        }
      }
    }
  }
  asm volatile("bar.sync 1, %0;" ::"r"(CONSUMERS) : "memory");
  // ---- phase B: row sums from shared memory
  const int sub = threadIdx.x % LPR;
  const int nscalar = (int)(r1 - r0) * BS;
  for (int base = 0; base < nscalar; base += ROWS) {
    const int lr = base + threadIdx.x / LPR;
    const bool live = lr < nscalar;
    const int64_t R = r0 + (live ? lr / BS : 0);
    const int i = live ? lr % BS : 0;
    const int64_t row = R * BS + i;
    int ks, ke;
    if (ra0 >= 0) { ks = (int)(rp[R - ra0] - al0); ke = (int)(rp[R + 1 - ra0] - al0); }
    else { ks = (int)(a.row_ptr[R] - al0); ke = (int)(a.row_ptr[R + 1] - al0); }
    if (!live) ke = ks;
    const double wv = (a.w && live && sub == 0) ? __ldg(a.w + row) : 0.0;
    const double wv2 = (a.w2 && live && sub == 0) ? __ldg(a.w2 + row) : 0.0;
    double acc0 = 0.0, acc1 = 0.0;
    int k = ks + sub;
    for (; k + LPR < ke; k += 2 * LPR) { acc0 += vs[k * BS * BS + i]; acc1 += vs[(k + LPR) * BS * BS + i]; }
    if (k < ke) acc0 += vs[k * BS * BS + i];
    double acc = acc0 + acc1;
#pragma unroll
    for (int o = LPR >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (live && sub == 0) {
      if (a.l2_hint) __stcs(a.y + row, acc); else a.y[row] = acc;
      d0 += acc * wv;
      if (a.want_yy) d1 += acc * acc;
      d2 += acc * wv2;
    }
  }
  // the stage's values were rewritten through the generic proxy; the next bulk copy into it goes through the async proxy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// SpMV configuration chosen for a matrix (fsb_spmv.cu)
struct SpmvPlan {
  int bs, rows, lpr, nst;
  size_t smem;       // dynamic shared memory of the pipeline
  int per_sm;        // resident CTAs per SM
};
int fsb_spmv_plan(fsb_mat* A, SpmvPlan* plan);      // FSB_OK and plan filled when the staged kernel applies (spmv_mode 0, tiled)
void fsb_spmv_fill_args(fsb_mat* A, SpmvArgs* a);   // matrix + tiling fields; vectors, dots and peer fields cleared
