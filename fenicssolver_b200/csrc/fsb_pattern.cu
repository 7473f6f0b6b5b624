// K2: DofMap + sparsity pattern on the device (dolfin SparsityPatternBuilder equivalent).
//
// vertex->cell adjacency (count, scan, fill) then one warp per row: the candidate columns are the
// vertices of the incident cells; duplicates are removed and the survivors ranked inside the warp, so
// row_ptr/col_idx come out sorted without any global sort.  A per-cell position map (uint8 offset of
// each local (a,b) entry inside row a) is emitted for the numeric phase.
#include "fsb_internal.cuh"
#include <algorithm>

// ------------------------------------------------------------------------------------ scan
static constexpr int kScanThreads = 512;
static constexpr int kScanItems = 8;   // per thread
static constexpr int kScanTile = kScanThreads * kScanItems;

__global__ void k_scan_block_sums(const int32_t* __restrict__ in, int64_t n, int64_t* __restrict__ bsum) {
  __shared__ long long smi[32];
  int64_t base = (int64_t)blockIdx.x * kScanTile;
  long long s = 0;
  for (int k = 0; k < kScanItems; ++k) {
    int64_t i = base + (int64_t)k * kScanThreads + threadIdx.x;
    if (i < n) s += in[i];
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) smi[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    long long t = threadIdx.x < (kScanThreads >> 5) ? smi[threadIdx.x] : 0;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) bsum[blockIdx.x] = t;
  }
}

// single block: exclusive scan of the block sums in place
__global__ void k_scan_spine(int64_t* __restrict__ bsum, int64_t nb) {
  __shared__ long long smi[32];
  __shared__ long long carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < nb; base += blockDim.x) {
    int64_t i = base + threadIdx.x;
    long long v = i < nb ? bsum[i] : 0, incl = v;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) {
      long long t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) smi[w] = incl;
    __syncthreads();
    if (w == 0) {
      long long t = lane < (int)(blockDim.x >> 5) ? smi[lane] : 0, ti = t;
      for (int o = 1; o < 32; o <<= 1) {
        long long u = __shfl_up_sync(0xffffffffu, ti, o);
        if (lane >= o) ti += u;
      }
      smi[lane] = ti - t;   // exclusive warp offsets
    }
    __syncthreads();
    long long carry = carry_s;
    if (i < nb) bsum[i] = carry + smi[w] + incl - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_s = carry + smi[w] + incl;
    __syncthreads();
  }
}

__global__ void k_scan_apply(const int32_t* __restrict__ in, int64_t n, const int64_t* __restrict__ bsum,
                             int64_t* __restrict__ out) {
  // items are laid out thread-major inside the tile so that each thread scans kScanItems consecutive values
  __shared__ long long smi[32];
  int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  long long v[kScanItems], s = 0;
  for (int k = 0; k < kScanItems; ++k) {
    int64_t i = base + k;
    v[k] = i < n ? in[i] : 0;
    s += v[k];
  }
  long long incl = s;
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int o = 1; o < 32; o <<= 1) {
    long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) smi[w] = incl;
  __syncthreads();
  if (w == 0) {
    long long t = lane < (kScanThreads >> 5) ? smi[lane] : 0, ti = t;
    for (int o = 1; o < 32; o <<= 1) {
      long long u = __shfl_up_sync(0xffffffffu, ti, o);
      if (lane >= o) ti += u;
    }
    smi[lane] = ti - t;
  }
  __syncthreads();
  long long run = bsum[blockIdx.x] + smi[w] + incl - s;
  for (int k = 0; k < kScanItems; ++k) {
    int64_t i = base + k;
    if (i < n) out[i] = run;
    run += v[k];
    if (i == n - 1) out[n] = run;
  }
}

int fsb_exclusive_scan(fsb_ctx* ctx, const int32_t* in, int64_t* out, int64_t n) {
  if (n <= 0) {
    int64_t z = 0;
    FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(out, &z, sizeof(z), cudaMemcpyHostToDevice, ctx->stream));
    FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FSB_OK;
  }
  int64_t nb = (n + kScanTile - 1) / kScanTile;
  int64_t* bsum = nullptr;
  int rc = fsb_dmalloc(ctx, &bsum, (size_t)nb);
  if (rc) return rc;
  // block sums use a strided read, the apply pass a thread-major read; both cover the same tile
  k_scan_block_sums<<<(unsigned)nb, kScanThreads, 0, ctx->stream>>>(in, n, bsum);
  FSB_LAUNCH_CHECK(ctx);
  k_scan_spine<<<1, 1024, 0, ctx->stream>>>(bsum, nb);
  FSB_LAUNCH_CHECK(ctx);
  k_scan_apply<<<(unsigned)nb, kScanThreads, 0, ctx->stream>>>(in, n, bsum, out);
  FSB_LAUNCH_CHECK(ctx);
  FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  fsb_dfree(ctx, bsum);
  return FSB_OK;
}

// ------------------------------------------------------------------------------------ vertex -> cell adjacency
__global__ void k_v2c_count(const int32_t* __restrict__ cells, int64_t nent, int32_t* __restrict__ deg) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nent; i += (int64_t)gridDim.x * blockDim.x)
    atomicAdd(deg + cells[i], 1);
}
__global__ void k_v2c_fill(const int32_t* __restrict__ cells, int64_t ncells, int nl, const int64_t* __restrict__ vptr,
                           int32_t* __restrict__ cursor, int32_t* __restrict__ v2c) {
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncells; c += (int64_t)gridDim.x * blockDim.x)
    for (int a = 0; a < nl; ++a) {
      int v = cells[c * nl + a];
      int slot = atomicAdd(cursor + v, 1);
      v2c[vptr[v] + slot] = (int32_t)c;
    }
}

// ascending cell order inside every vertex's list: the atomic cursor of k_v2c_fill leaves them in arrival order, and the row-gather
// assembly (fsb_assemble_rows.cu) sums a row's contributions in list order — sorted lists make it bitwise reproducible
__global__ void k_v2c_sort(int64_t nverts, const int64_t* __restrict__ vptr, int32_t* __restrict__ v2c) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nverts; v += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p0 = vptr[v], p1 = vptr[v + 1];
    for (int64_t i = p0 + 1; i < p1; ++i) {
      const int32_t c = v2c[i];
      int64_t j = i;
      while (j > p0 && v2c[j - 1] > c) { v2c[j] = v2c[j - 1]; --j; }
      v2c[j] = c;
    }
  }
}

int fsb_mesh_sort_adjacency(fsb_mesh* mesh) {
  fsb_ctx* ctx = mesh->ctx;
  if (!mesh->v2c || mesh->v2c_sorted) return FSB_OK;
  k_v2c_sort<<<fsb_grid(mesh->nverts, 256, (int64_t)ctx->sm_count * 16), 256, 0, ctx->stream>>>(mesh->nverts, mesh->v2c_ptr, mesh->v2c);
  FSB_LAUNCH_CHECK(ctx);
  mesh->v2c_sorted = true;
  return FSB_OK;
}

// ------------------------------------------------------------------------------------ rows
// One warp per row.  Candidates L[i] = cells[v2c[p0 + i/nl]][i%nl], i < m = deg*nl.  They are sorted in
// shared memory when m <= kRowCap, otherwise scanned from global memory (L1).
static constexpr int kRowCap = 512;      // power of two
static constexpr int kRowWarps = 8;

// The counting pass (FILL = false) also parks the sorted columns of every row with at most kRowKeep entries in `keep`
// (stride kRowKeep), so the fill pass only has to copy them; rows longer than that are sorted again by the fill pass
// (FILL = true, which skips the short rows when `keep` is given).
static constexpr int kRowKeep = 32;

template <bool FILL>
__global__ void __launch_bounds__(kRowWarps * 32)
k_rows(const int32_t* __restrict__ cells, int nl, const int64_t* __restrict__ vptr, const int32_t* __restrict__ v2c,
       int64_t nrows, int32_t* __restrict__ row_len, const int64_t* __restrict__ row_ptr, int32_t* __restrict__ col_idx,
       int32_t* __restrict__ keep) {
  __shared__ int32_t s_cand[kRowWarps][kRowCap];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int64_t r = (int64_t)blockIdx.x * kRowWarps + w; r < nrows; r += (int64_t)gridDim.x * kRowWarps) {
    if (FILL && keep && row_ptr[r + 1] - row_ptr[r] <= kRowKeep) continue;      // copied from `keep` by k_rows_copy
    const int64_t p0 = vptr[r];
    const int m = (int)(vptr[r + 1] - p0) * nl;
    if (!FILL && keep && m <= 256) {
      // fast path of the counting pass: every neighbour appears ~6 times among the candidates, so insert them into a
      // 64-slot shared-memory hash set (32-bit CAS), and if at most kRowKeep distinct columns come out, sort those 32
      // values inside the warp with shuffles: no 128-element shared-memory sort
      int32_t* tab = s_cand[w];
      tab[lane] = -1; tab[lane + 32] = -1;
      __syncwarp();
      bool ovf = false;
      for (int i = lane; i < m; i += 32) {
        const int32_t v = cells[(int64_t)v2c[p0 + i / nl] * nl + i % nl];
        unsigned h = ((unsigned)v * 2654435761u) >> 26;
        for (int probes = 0;; ++probes) {
          const int32_t old = atomicCAS(&tab[h], -1, v);
          if (old == -1 || old == v) break;
          if (probes >= 64) { ovf = true; break; }
          h = (h + 1) & 63;
        }
      }
      __syncwarp();
      ovf = __any_sync(0xffffffffu, ovf);
      const int32_t a0 = tab[lane], a1 = tab[lane + 32];
      const unsigned m0 = __ballot_sync(0xffffffffu, a0 != -1), m1 = __ballot_sync(0xffffffffu, a1 != -1);
      const int count = __popc(m0) + __popc(m1);
      __syncwarp();
      if (!ovf && count <= kRowKeep) {
        const unsigned lt = (1u << lane) - 1;
        if (a0 != -1) tab[64 + __popc(m0 & lt)] = a0;
        if (a1 != -1) tab[64 + __popc(m0) + __popc(m1 & lt)] = a1;
        __syncwarp();
        int32_t v = lane < count ? tab[64 + lane] : 0x7fffffff;
#pragma unroll
        for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
          for (int j = k >> 1; j > 0; j >>= 1) {
            const int32_t o = __shfl_xor_sync(0xffffffffu, v, j);
            const bool up = (lane & k) == 0, lower = (lane & j) == 0;
            v = (lower == up) ? min(v, o) : max(v, o);
          }
        if (lane < count) keep[r * kRowKeep + lane] = v;
        if (lane == 0) row_len[r] = count;
        __syncwarp();
        continue;
      }
    }
    const bool cached = m <= kRowCap;
    if (cached) {
      // common case: bitonic-sort the candidates in shared memory (padded with INT_MAX to a power of
      // two), then the first element of each run of equal values is a column; ranks come from ballots
      int P = 32;
      while (P < m) P <<= 1;
      for (int i = lane; i < P; i += 32)
        s_cand[w][i] = i < m ? cells[(int64_t)v2c[p0 + i / nl] * nl + i % nl] : 0x7fffffff;
      __syncwarp();
      for (int k = 2; k <= P; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
          for (int t = lane; t < (P >> 1); t += 32) {
            const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1)), hi = lo | j;     // t-th compare-exchange pair
            const int32_t a = s_cand[w][lo], b = s_cand[w][hi];
            const bool up = (lo & k) == 0;
            if ((a > b) == up) { s_cand[w][lo] = b; s_cand[w][hi] = a; }
          }
          __syncwarp();
        }
      int count = 0;
      for (int i0 = 0; i0 < m; i0 += 32) {
        const int i = i0 + lane;
        const int32_t v = i < m ? s_cand[w][i] : 0x7fffffff;
        const bool first = i < m && (i == 0 || s_cand[w][i - 1] != v);
        const unsigned mask = __ballot_sync(0xffffffffu, first);
        const int slot = count + __popc(mask & ((1u << lane) - 1));
        if (FILL && first) col_idx[row_ptr[r] + slot] = v;
        if (!FILL && keep && first && slot < kRowKeep) keep[r * kRowKeep + slot] = v;
        count += __popc(mask);
      }
      if (!FILL && lane == 0) row_len[r] = count;
      __syncwarp();
      continue;
    }
    // rare case (valence above the shared-memory cache): quadratic scan straight from global memory
    auto cand = [&](int i) -> int32_t { return cells[(int64_t)v2c[p0 + i / nl] * nl + i % nl]; };
    // pass 1: first occurrences
    int nfirst = 0;
    for (int i0 = 0; i0 < m; i0 += 32) {
      int i = i0 + lane;
      bool first = i < m;
      int32_t v = first ? cand(i) : 0;
      for (int j = 0; j < i0 + 32 && j < m; ++j) {
        int32_t u = cand(j);
        if (j < i && u == v) first = false;
      }
      nfirst += __popc(__ballot_sync(0xffffffffu, first));
      if (FILL) {
        // rank directly (first flags recomputed on the fly below)
        if (first) {
          int pos = 0;
          for (int j = 0; j < m; ++j) {
            int32_t u = cand(j);
            if (u < v) {
              bool jf = true;
              for (int k = 0; k < j; ++k) if (cand(k) == u) { jf = false; break; }
              pos += jf;
            }
          }
          col_idx[row_ptr[r] + pos] = v;
        }
      }
    }
    __syncwarp();
    if (!FILL && lane == 0) row_len[r] = nfirst;
    __syncwarp();
  }
}

// fill pass for the short rows: one thread per kept slot
__global__ void k_rows_copy(int64_t nrows, const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ keep,
                            int32_t* __restrict__ col_idx) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nrows * kRowKeep; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / kRowKeep;
    const int k = (int)(i % kRowKeep);
    const int64_t base = row_ptr[r];
    const int len = (int)(row_ptr[r + 1] - base);
    if (len <= kRowKeep && k < len) col_idx[base + k] = keep[i];
  }
}

// position map: for cell c and local pair (a,b): offset of column cells[c][b] inside row cells[c][a]
__global__ void k_posmap(const int32_t* __restrict__ cells, int64_t ncells, int nl, const int64_t* __restrict__ row_ptr,
                         const int32_t* __restrict__ col_idx, uint8_t* __restrict__ posmap) {
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncells; c += (int64_t)gridDim.x * blockDim.x) {
    int v[10];
    for (int a = 0; a < nl; ++a) v[a] = cells[c * nl + a];
    for (int a = 0; a < nl; ++a) {
      const int64_t base = row_ptr[v[a]];
      const int len = (int)(row_ptr[v[a] + 1] - base);
      int lo = 0;
      for (int b = 0; b < nl; ++b) {
        if (b > 0 && v[b] < v[b - 1]) lo = 0;      // P2 node lists are ascending only within vertices / within runs of edges
        lo = row_find(col_idx + base, lo, len, v[b]);
        posmap[c * nl * nl + a * nl + b] = (uint8_t)lo;
        ++lo;
      }
    }
  }
}

__global__ void k_max_i32(const int32_t* __restrict__ in, int64_t n, int32_t* __restrict__ out) {
  int m = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = max(m, in[i]);
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// ------------------------------------------------------------------------------------ matrix objects
extern "C" void fsb_mat_destroy(fsb_mat* A) {
  if (!A) return;
  fsb_dfree(A->ctx, A->row_ptr);
  fsb_dfree(A->ctx, A->col_idx);
  fsb_dfree(A->ctx, A->vals);
  fsb_dfree(A->ctx, A->posmap);
  fsb_dfree(A->ctx, A->plan_rank); fsb_dfree(A->ctx, A->plan_head); fsb_dfree(A->ctx, A->plan_run_ptr); fsb_dfree(A->ctx, A->plan_dest);
  fsb_dfree(A->ctx, A->tile_row);
  fsb_dfree(A->ctx, A->bc_flag);
  fsb_dfree(A->ctx, A->bc_val);
  fsb_dfree(A->ctx, A->bc_dofs);
  fsb_dfree(A->ctx, A->bc_vals);
  for (double* w : A->work) fsb_dfree(A->ctx, w);
  fsb_dist_release_mat(A);
  if (A->sq) fsb_mat_destroy(A->sq);
  delete A;
}

extern "C" int fsb_mat_create(fsb_mesh* mesh, int32_t ncomp, fsb_mat** out) {
  if (!mesh || !out) return FSB_ERR_ARG;
  fsb_ctx* ctx = mesh->ctx;
  if (ncomp < 1 || ncomp > 3) FSB_FAIL(ctx, FSB_ERR_ARG, "ncomp must be 1..3");
  const int nl = mesh->nl;
  const int64_t nv = mesh->nnodes, nc = mesh->ncells;
  const int cap = ctx->sm_count * 16;
  fsb_mat* A = new fsb_mat();
  A->ctx = ctx; A->mesh = mesh; A->bs = ncomp; A->nbrows = nv; A->own0 = 0; A->own1 = nv;
  int32_t *deg = nullptr, *v2c = nullptr, *d_max = nullptr, *keep = nullptr;
  int64_t* vptr = nullptr;
  int rc = FSB_OK;
  auto cleanup = [&]() {
    fsb_dfree(ctx, deg); fsb_dfree(ctx, d_max); fsb_dfree(ctx, keep);
    if (v2c != mesh->v2c) fsb_dfree(ctx, v2c);          // the mesh owns a kept adjacency
    if (vptr != mesh->v2c_ptr) fsb_dfree(ctx, vptr);
    v2c = nullptr; vptr = nullptr;
  };
#define TRY(x) do { rc = (x); if (rc) { cleanup(); fsb_mat_destroy(A); return rc; } } while (0)
#define TRYCUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { ctx->err = std::string(#x) + ": " + cudaGetErrorString(e_); cleanup(); fsb_mat_destroy(A); return FSB_ERR_CUDA; } } while (0)
  TRY(fsb_dmalloc(ctx, &deg, (size_t)nv + 1));
  TRY(fsb_dmalloc(ctx, &d_max, 1));
  // vertex -> cell adjacency.  A degree-1 mesh keeps it (sorted per vertex) for the row-gather assembly kernels; a second
  // matrix on the same mesh reuses it
  const bool keep_adj = mesh->degree == 1;
  if (keep_adj && mesh->v2c) {
    vptr = mesh->v2c_ptr; v2c = mesh->v2c;
  } else {
    TRY(fsb_dmalloc(ctx, &vptr, (size_t)nv + 1));
    TRY(fsb_dmalloc(ctx, &v2c, (size_t)nc * nl));
    TRYCUDA(cudaMemsetAsync(deg, 0, sizeof(int32_t) * (nv + 1), ctx->stream));
    k_v2c_count<<<fsb_grid(nc * nl, 256, cap), 256, 0, ctx->stream>>>(mesh->cell_nodes, nc * nl, deg);
    ctx->launches++; TRYCUDA(cudaGetLastError());
    TRY(fsb_exclusive_scan(ctx, deg, vptr, nv));
    TRYCUDA(cudaMemsetAsync(deg, 0, sizeof(int32_t) * (nv + 1), ctx->stream));
    k_v2c_fill<<<fsb_grid(nc, 256, cap), 256, 0, ctx->stream>>>(mesh->cell_nodes, nc, nl, vptr, deg, v2c);
    ctx->launches++; TRYCUDA(cudaGetLastError());
    if (keep_adj) { mesh->v2c_ptr = vptr; mesh->v2c = v2c; mesh->v2c_sorted = false; }
  }
  // row lengths -> row_ptr
  TRY(fsb_dmalloc(ctx, &A->row_ptr, (size_t)nv + 1));
  // scratch for the sorted short rows (128 B per row); without it the fill pass simply sorts every row again
  if (fsb_dmalloc(ctx, &keep, (size_t)nv * kRowKeep) != FSB_OK) { keep = nullptr; ctx->err.clear(); }
  k_rows<false><<<fsb_grid(nv, kRowWarps, cap), kRowWarps * 32, 0, ctx->stream>>>(mesh->cell_nodes, nl, vptr, v2c, nv, deg, nullptr, nullptr, keep);
  ctx->launches++; TRYCUDA(cudaGetLastError());
  TRYCUDA(cudaMemsetAsync(d_max, 0, sizeof(int32_t), ctx->stream));
  k_max_i32<<<fsb_grid(nv, 256, cap), 256, 0, ctx->stream>>>(deg, nv, d_max);
  ctx->launches++; TRYCUDA(cudaGetLastError());
  TRY(fsb_exclusive_scan(ctx, deg, A->row_ptr, nv));
  int64_t nnzb = 0;
  int32_t maxlen = 0;
  TRYCUDA(cudaMemcpyAsync(&nnzb, A->row_ptr + nv, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  TRYCUDA(cudaMemcpyAsync(&maxlen, d_max, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  TRYCUDA(cudaStreamSynchronize(ctx->stream));
  A->nnzb = nnzb;
  A->max_row_len = maxlen;
  TRY(fsb_dmalloc(ctx, &A->col_idx, (size_t)nnzb));
  if (keep) {
    k_rows_copy<<<fsb_grid(nv * kRowKeep, 256, (int64_t)ctx->sm_count * 32), 256, 0, ctx->stream>>>(nv, A->row_ptr, keep, A->col_idx);
    ctx->launches++; TRYCUDA(cudaGetLastError());
  }
  if (!keep || maxlen > kRowKeep) {
    k_rows<true><<<fsb_grid(nv, kRowWarps, cap), kRowWarps * 32, 0, ctx->stream>>>(mesh->cell_nodes, nl, vptr, v2c, nv, nullptr, A->row_ptr, A->col_idx, keep);
    ctx->launches++; TRYCUDA(cudaGetLastError());
  }
  TRYCUDA(cudaStreamSynchronize(ctx->stream));
  if (v2c != mesh->v2c) fsb_dfree(ctx, v2c);
  if (vptr != mesh->v2c_ptr) fsb_dfree(ctx, vptr);
  v2c = nullptr; vptr = nullptr;
  fsb_dfree(ctx, keep); keep = nullptr;
  TRY(fsb_dmalloc(ctx, &A->vals, (size_t)nnzb * ncomp * ncomp));
  TRYCUDA(cudaMemsetAsync(A->vals, 0, sizeof(double) * nnzb * ncomp * ncomp + 512, ctx->stream));
  if (maxlen <= 256) {
    TRY(fsb_dmalloc(ctx, &A->posmap, (size_t)nc * nl * nl));
    k_posmap<<<fsb_grid(nc, 256, cap), 256, 0, ctx->stream>>>(mesh->cell_nodes, nc, nl, A->row_ptr, A->col_idx, A->posmap);
    ctx->launches++; TRYCUDA(cudaGetLastError());
  }
  TRY(fsb_mat_setup_tiles(A));
  TRYCUDA(cudaStreamSynchronize(ctx->stream));
  cleanup();
#undef TRY
#undef TRYCUDA
  *out = A;
  return FSB_OK;
}

extern "C" int fsb_mat_from_csr(fsb_ctx* ctx, int64_t nrows, const int64_t* row_ptr, const int32_t* col_idx,
                                const double* vals, fsb_mat** out) {
  if (!ctx || !out || !row_ptr || !col_idx || !vals || nrows <= 0) return FSB_ERR_ARG;
  fsb_mat* A = new fsb_mat();
  A->ctx = ctx; A->bs = 1; A->nbrows = nrows; A->own0 = 0; A->own1 = nrows;
  A->nnzb = row_ptr[nrows];
  int maxlen = 0;
  for (int64_t r = 0; r < nrows; ++r) maxlen = std::max<int64_t>(maxlen, row_ptr[r + 1] - row_ptr[r]);
  A->max_row_len = maxlen;
  int rc = fsb_dmalloc(ctx, &A->row_ptr, (size_t)nrows + 1);
  if (!rc) rc = fsb_dmalloc(ctx, &A->col_idx, (size_t)A->nnzb);
  if (!rc) rc = fsb_dmalloc(ctx, &A->vals, (size_t)A->nnzb);
  if (rc) { fsb_mat_destroy(A); return rc; }
  cudaMemsetAsync(A->col_idx + A->nnzb, 0, 512, ctx->stream);
  cudaMemsetAsync(A->vals + A->nnzb, 0, 512, ctx->stream);
  FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(A->row_ptr, row_ptr, sizeof(int64_t) * (nrows + 1), cudaMemcpyHostToDevice, ctx->stream));
  FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(A->col_idx, col_idx, sizeof(int32_t) * A->nnzb, cudaMemcpyHostToDevice, ctx->stream));
  FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(A->vals, vals, sizeof(double) * A->nnzb, cudaMemcpyHostToDevice, ctx->stream));
  FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  rc = fsb_mat_setup_tiles(A);
  if (rc) { fsb_mat_destroy(A); return rc; }
  *out = A;
  return FSB_OK;
}

extern "C" int fsb_mat_sizes(fsb_mat* A, int64_t* nrows, int64_t* nnz, int32_t* bs, int64_t* nnzb) {
  if (!A) return FSB_ERR_ARG;
  if (nrows) *nrows = A->nbrows * A->bs;
  if (nnz) *nnz = A->nnzb * A->bs * A->bs;
  if (bs) *bs = A->bs;
  if (nnzb) *nnzb = A->nnzb;
  return FSB_OK;
}

extern "C" int fsb_mat_zero(fsb_mat* A) {
  if (!A) return FSB_ERR_ARG;
  FSB_CHECK_CUDA(A->ctx, cudaMemsetAsync(A->vals, 0, sizeof(double) * A->nnzb * A->bs * A->bs, A->ctx->stream));
  return FSB_OK;
}

extern "C" int fsb_mat_set_owned_rows(fsb_mat* A, int64_t row0, int64_t row1) {
  if (!A) return FSB_ERR_ARG;
  if (row0 < 0 || row1 > A->nbrows || row0 > row1) FSB_FAIL(A->ctx, FSB_ERR_ARG, "bad owned row range");
  A->own0 = row0; A->own1 = row1;
  return fsb_mat_setup_tiles(A);
}

// scalar CSR view of the block matrix: scalar row bs*R+i holds, for each block (R,C) in order, the
// columns bs*C+0..bs-1, so columns stay sorted.
extern "C" int fsb_mat_download_csr(fsb_mat* A, int64_t* row_ptr, int32_t* col_idx, double* vals) {
  if (!A) return FSB_ERR_ARG;
  fsb_ctx* ctx = A->ctx;
  const int bs = A->bs;
  std::vector<int64_t> rp((size_t)A->nbrows + 1);
  FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(rp.data(), A->row_ptr, sizeof(int64_t) * (A->nbrows + 1), cudaMemcpyDeviceToHost, ctx->stream));
  FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (bs == 1) {
    if (row_ptr) memcpy(row_ptr, rp.data(), sizeof(int64_t) * (A->nbrows + 1));
    if (col_idx) FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(col_idx, A->col_idx, sizeof(int32_t) * A->nnzb, cudaMemcpyDeviceToHost, ctx->stream));
    if (vals) FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(vals, A->vals, sizeof(double) * A->nnzb, cudaMemcpyDeviceToHost, ctx->stream));
    FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FSB_OK;
  }
  std::vector<int32_t> ci((size_t)A->nnzb);
  std::vector<double> bv;
  FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(ci.data(), A->col_idx, sizeof(int32_t) * A->nnzb, cudaMemcpyDeviceToHost, ctx->stream));
  if (vals) {
    bv.resize((size_t)A->nnzb * bs * bs);
    FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(bv.data(), A->vals, sizeof(double) * bv.size(), cudaMemcpyDeviceToHost, ctx->stream));
  }
  FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int64_t R = 0; R < A->nbrows; ++R) {
    const int64_t len = rp[R + 1] - rp[R];
    for (int i = 0; i < bs; ++i) {
      const int64_t srow = R * bs + i;
      const int64_t sbase = rp[R] * bs * bs + (int64_t)i * len * bs;
      if (row_ptr) row_ptr[srow] = sbase;
      for (int64_t k = 0; k < len; ++k)
        for (int j = 0; j < bs; ++j) {
          if (col_idx) col_idx[sbase + k * bs + j] = ci[rp[R] + k] * bs + j;
          if (vals) vals[sbase + k * bs + j] = bv[(rp[R] + k) * bs * bs + i * bs + j];
        }
    }
  }
  if (row_ptr) row_ptr[A->nbrows * bs] = A->nnzb * bs * bs;
  return FSB_OK;
}
