// Internal definitions shared by the libfsb translation units (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <chrono>
#include <cstdlib>
#include <string>
#include <unordered_map>
#include <vector>

#include "fsb.h"

struct fsb_dist;   // fsb_dist.cu

struct fsb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sm_count = 148;
  std::string err;
  int64_t launches = 0;
  // options
  int asm_mode = 1;      // 0 search+atomics, 1 position map + atomics (default), 2 row-gather (owner computes, no atomics, bitwise
                         // reproducible; measured 2x slower than 1 at 256^3, profiles/asm_ab_r2.txt) for the degree-1 scalar forms,
                         // 3 per-warp combine plan (sorted contributions, one RED per distinct slot of a warp) for P1 tetrahedra
  int spmv_mode = 0;     // 0 TMA-staged tiles v2, 1 plain row-per-thread, 2 TMA-staged v1 (one thread per row)
  int spmv_lpr = 0;      // lanes per row in the staged kernel (1, 2 or 4; 0 = best per block size)
  int spmv_rows = 0;     // scalar rows per tile: 256 or 128 (x 3/4 for 3x3 blocks; 0 = best per block size)
  int spmv_stages = 0;   // TMA pipeline depth (2..4; 0 = 2)
  int spmv_flat = 0;     // two-phase tiles (products over non-zeros, then row sums): 0 = long rows of 3x3 blocks, 1 always, 2 never
  int dist_p2p = 1;      // distributed CG through peer-memory mailboxes/halo (1) or NCCL collectives (0)
  int profile = 0;
  int use_graph = 1;
  int check_every = 32;
  int spmv_hint = 1;     // L2 evict-first on the matrix stream + streaming stores of y (0: plain)
  int cg_variant = 0;    // 0 auto (persistent kernel where it applies, else classic on one GPU / single-reduction when distributed),
                         // 1 classic 3-kernel chain, 2 single-reduction 2-kernel chain, 3 persistent kernel (fsb_cgp.cu)
  int cg_minb = 0;       // persistent kernel, one-lane configuration: CTAs per SM (2 or 4; 0 = by slab size)
  int cg_umode = 1;      // persistent kernel, update phase: 0 contiguous row slice per CTA, 1 grid-stride (measured faster)
  int vec_skew = 0;      // bytes between the start offsets of consecutive Krylov work vectors inside their blocks (multiple of 256, <= 8192)
  int cg_debug = 0;      // persistent kernel timing experiments (wrong results): 1 no peer stores, 2 no system fence after them
  int cg_timeout_s = 30; // watchdog of the persistent kernel's spin loops (a lost peer traps instead of hanging the GPU)
  int drop_zeros = 2;    // Krylov SpMVs run on a copy without the exactly-zero blocks (fsb_squeeze.cu): 0 never, 1 always,
                         // 2 (default) when at least 20 % of the stored blocks are exactly zero
  // scratch for reductions: per-CTA partials + a few scalars, and a pinned host mirror
  double* d_partials = nullptr;   // [kMaxPartials * 4]
  double* d_scalars = nullptr;    // [64]
  unsigned* d_counters = nullptr; // [16] last-block counters (self-resetting)
  int* d_state = nullptr;         // [8] Krylov state: done, iterations, outcome
  double* h_pinned = nullptr;     // [64]
  int* h_state = nullptr;         // [16] pinned mirror of d_state (two polling slots)
  void* h_stage[2] = {nullptr, nullptr};   // pinned staging chunks for large downloads into pageable memory
  cudaEvent_t stage_done[2] = {nullptr, nullptr};
  fsb_dist* dist = nullptr;
  void* cg_comm = nullptr;            // single-GPU mailbox of the persistent CG kernel (a CommBuf)
  unsigned long long cg_seq = 1;      // its next unused sequence number
  // exact-size block cache in front of the stream-ordered pool (see fsb_dmalloc)
  std::unordered_multimap<size_t, void*> free_blocks;
  std::unordered_map<void*, size_t> live_blocks;
  size_t cached_bytes = 0;
  size_t cache_limit = (size_t)96 << 30;
  // page-locked host blocks handed out by fsb_host_alloc, and the released ones kept for exact-size reuse
  std::unordered_map<void*, size_t> host_live;
  std::unordered_multimap<size_t, void*> host_free;
  size_t host_cached = 0;
};

static constexpr int kMaxPartials = 4096;

struct fsb_mesh {
  fsb_ctx* ctx;
  int gdim, tdim;
  int64_t nverts, ncells;
  double* xyz = nullptr;     // [nverts][gdim]
  int32_t* cells = nullptr;  // [ncells][tdim+1] sorted per cell
  // Lagrange node layout the pattern and the element kernels work on.  Degree 1: the vertices themselves
  // (cell_nodes == cells, nl == tdim+1, nnodes == nverts).  Degree 2: vertices then edge nodes.
  int degree = 1;
  int nl = 0;                      // nodes per cell
  int64_t nnodes = 0;
  int32_t* cell_nodes = nullptr;   // [ncells][nl]; owned only when degree == 2
  double* p2_tables = nullptr;     // device copy of the P2 reference tensors for this dimension (degree 2)
  // vertex -> cell adjacency of a degree-1 mesh (built with the first matrix, cell ids ascending per vertex): the row-gather
  // assembly kernels walk it
  int64_t* v2c_ptr = nullptr;      // [nverts+1]
  int32_t* v2c = nullptr;          // [ncells*(tdim+1)]
  bool v2c_sorted = false;         // cell ids ascending inside every vertex's list (fsb_mesh_sort_adjacency)
  // K1 (fsb_facets.cu): exterior facets in lexicographic order, made on first request
  int64_t nbf = -1, nfacets = 0;   // exterior facets / distinct facets of the mesh
  int32_t* bf_verts = nullptr;     // [nbf][tdim] sorted vertex tuples
  int32_t* bf_opp = nullptr;       // [nbf] vertex of the cell opposite the facet
  int32_t* bf_cell = nullptr;      // [nbf] the one cell holding the facet
  int64_t* bf_id = nullptr;        // [nbf] dolfin facet index = rank among all distinct facets (complete once bf_id_ranked)
  bool bf_id_ranked = false;
  int64_t nbv = -1;                // distinct boundary vertices (fsb_mesh_boundary_geometry)
  int32_t* bg_verts = nullptr;     // [nbv] ascending
  int32_t* bg_finv = nullptr;      // [nbf][tdim] facet vertices as indices into bg_verts
  double* bg_xyz = nullptr;        // [nbv][gdim]
  double* bg_mid = nullptr;        // [nbf][gdim] facet midpoints
};

struct fsb_vec {
  fsb_ctx* ctx;
  int64_t n;
  double* d = nullptr;
};

struct fsb_mat {
  fsb_ctx* ctx;
  fsb_mesh* mesh = nullptr;    // pattern source (may be null for from_csr)
  int bs = 1;                  // block size (ncomp)
  int64_t nbrows = 0;          // block rows
  int64_t nnzb = 0;            // blocks
  int64_t* row_ptr = nullptr;  // [nbrows+1]
  int32_t* col_idx = nullptr;  // [nnzb] (+pad)
  double* vals = nullptr;      // [nnzb][bs][bs] (+pad)
  uint8_t* posmap = nullptr;   // [ncells][(tdim+1)^2] in-row offsets, or null
  // per-warp combine plan of the degree-1 tetrahedron scatter (asm_mode 3, fsb_assemble.cu), built on first use
  uint16_t* plan_rank = nullptr;   // [nchunk][16][32]
  uint16_t* plan_head = nullptr;   // [nchunk][32]
  int64_t* plan_run_ptr = nullptr; // [nchunk+1]
  uint32_t* plan_dest = nullptr;   // [plan_runs]
  int64_t plan_runs = 0;
  int max_row_len = 0;
  int64_t own0 = 0, own1 = 0;  // owned block rows
  // SpMV tiling (TMA-staged): tiles of ~tile_nnz blocks snapped to row boundaries
  int tile_nnz = 0;
  int64_t ntiles = 0;
  int64_t* tile_row = nullptr; // [2*(ntiles+1)]: first block row of each tile, then row_ptr at that row
  int tile_cap = 0;            // smem capacity in blocks per stage
  int tile_rows = 0;           // scalar rows per tile this tiling was built for
  double avg_row = 0.0;        // blocks per owned block row (picks the SpMV configuration)
  size_t stage_bytes = 0;      // shared memory per pipeline stage (values + columns + row_ptr slice)
  // dirichlet scratch
  uint8_t* bc_flag = nullptr;  // [nbrows*bs]
  double* bc_val = nullptr;
  int64_t* bc_dofs = nullptr;  // staging for the uploaded Dirichlet list (capacity bc_cap)
  double* bc_vals = nullptr;
  int64_t bc_cap = 0;
  // distributed CG over peer memory: the search direction lives in a cudaMalloc'ed (IPC-exportable) vector
  // whose ghost planes the neighbours write directly   [fsb_dist.cu]
  double* p_dist = nullptr;
  int64_t p_dist_n = 0;
  double* p_lo_remote = nullptr;   // rank-1's p vector (peer mapping) or null
  double* p_hi_remote = nullptr;   // rank+1's p vector
  int64_t lo_ghost_offset = 0;     // index of rank-1's upper ghost plane inside its p vector
  // Krylov work vectors, kept across solves (transient runs re-solve every step)
  double* work[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int work_count = 0;
  int last_iters = 0;          // iterations of the previous Krylov solve on this matrix (sizes the first launch batch)
  // "drop_zeros": compacted copy (blocks with a non-zero entry only) the Krylov SpMVs run on   [fsb_squeeze.cu]
  fsb_mat* sq = nullptr;
  int64_t sq_cap = 0;          // capacity (blocks) of sq's col_idx / vals, kept on the squeezed matrix itself
};

#define FSB_CHECK_CUDA(ctx, call)                                                         \
  do {                                                                                    \
    cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + \
                   ":" + std::to_string(__LINE__) + ")";                                  \
      return e__ == cudaErrorMemoryAllocation ? FSB_ERR_NOMEM : FSB_ERR_CUDA;             \
    }                                                                                     \
  } while (0)

#define FSB_FAIL(ctx, code, msg) \
  do {                           \
    (ctx)->err = (msg);          \
    return (code);               \
  } while (0)

#define FSB_LAUNCH_CHECK(ctx)                   \
  do {                                          \
    (ctx)->launches++;                          \
    FSB_CHECK_CUDA(ctx, cudaGetLastError());    \
  } while (0)

// Device buffers.  Every kernel of a context runs on ctx->stream, so a block released by the host can be handed
// to the next request of the same size without waiting: work already queued on the old owner is ordered before
// anything the new owner enqueues.  A solver rebuilt on the same mesh (every transient restart, every bench
// step) asks for exactly the sizes the previous one released, so it is served from this cache and never reaches
// the driver allocator (whose pool grows by mapping fresh physical memory: ~0.1 ms per 2 MB, measured as
// 0.4-1.0 s per 256^3 solver construction).  Misses fall through to cudaMallocAsync.
static inline void fsb_cache_flush(fsb_ctx* ctx) {
  for (auto& kv : ctx->free_blocks) cudaFreeAsync(kv.second, ctx->stream);
  ctx->free_blocks.clear();
  ctx->cached_bytes = 0;
}

static inline int fsb_dmalloc_bytes(fsb_ctx* ctx, void** p, size_t bytes) {
  auto hit = ctx->free_blocks.find(bytes);
  if (hit != ctx->free_blocks.end()) {
    *p = hit->second;
    ctx->free_blocks.erase(hit);
    ctx->cached_bytes -= bytes;
    ctx->live_blocks[*p] = bytes;
    return FSB_OK;
  }
  static const bool trace = getenv("FSB_ALLOC_TRACE") != nullptr;
  auto t0 = std::chrono::steady_clock::now();
  void* q = nullptr;
  cudaError_t e = cudaMallocAsync(&q, bytes, ctx->stream);
  if (e != cudaSuccess && !ctx->free_blocks.empty()) {     // give the cached blocks back and try once more
    cudaGetLastError();
    fsb_cache_flush(ctx);
    cudaStreamSynchronize(ctx->stream);
    e = cudaMallocAsync(&q, bytes, ctx->stream);
  }
  if (trace) {
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (ms > 0.5) fprintf(stderr, "libfsb: cudaMallocAsync(%zu MB) took %.2f ms\n", bytes >> 20, ms);
  }
  if (e != cudaSuccess) {
    ctx->err = std::string("cudaMalloc of ") + std::to_string(bytes) + " bytes: " + cudaGetErrorString(e);
    *p = nullptr;
    cudaGetLastError();
    return FSB_ERR_NOMEM;
  }
  *p = q;
  ctx->live_blocks[q] = bytes;
  return FSB_OK;
}

// pad: TMA tiles may over-read a few entries
template <typename T>
static inline int fsb_dmalloc(fsb_ctx* ctx, T** p, size_t count, size_t pad_bytes = 512) {
  void* q = nullptr;
  int rc = fsb_dmalloc_bytes(ctx, &q, count * sizeof(T) + pad_bytes);
  *p = (T*)q;
  return rc;
}

static inline void fsb_dfree(fsb_ctx* ctx, void* p) {
  if (!p) return;
  auto it = ctx->live_blocks.find(p);
  if (it == ctx->live_blocks.end()) { cudaFreeAsync(p, ctx->stream); return; }
  const size_t bytes = it->second;
  ctx->live_blocks.erase(it);
  if (ctx->cached_bytes + bytes > ctx->cache_limit) { cudaFreeAsync(p, ctx->stream); return; }
  ctx->free_blocks.emplace(bytes, p);
  ctx->cached_bytes += bytes;
}

static inline unsigned fsb_grid(int64_t n, int block, int64_t cap = (1ll << 31) - 1) {
  int64_t g = (n + block - 1) / block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (unsigned)g;
}

// device-wide exclusive scan int32 -> int64 (out has n+1 entries, out[n] = total)   [fsb_pattern.cu]
int fsb_exclusive_scan(fsb_ctx* ctx, const int32_t* in, int64_t* out, int64_t n);
int fsb_mesh_sort_adjacency(fsb_mesh* mesh);      // ascending cell ids per vertex in mesh->v2c (once; no-op afterwards)
// SpMV tiling setup after row_ptr / owned range are known, and y = A x (+ fused dots d0 = y.w, d1 = y.y
// written to out[0..2)) on the ctx stream; `done` is an optional device early-exit flag   [fsb_spmv.cu]
int fsb_mat_setup_tiles(fsb_mat* A);
struct fsb_spmv_dist;   // fsb_device.cuh: peer-memory wait/post instructions for one launch
int fsb_launch_spmv(fsb_mat* A, const double* x, double* y, const double* w, int want_yy, double* out, const int* done,
                    const fsb_spmv_dist* dd = nullptr, const double* w2 = nullptr);
bool fsb_spmv_supports_p2p(fsb_mat* A);
int fsb_mat_squeeze(fsb_mat* A, fsb_mat** out);   // [fsb_squeeze.cu]
// degree-2 element kernels [fsb_assemble_p2.cu]; A == nullptr selects the matrix-free action y += (...) x
int fsb_p2_scalar(fsb_mesh* mesh, fsb_mat* A, fsb_vec* x, fsb_vec* y, double kscale, const double* ktensor, double mass, double adv,
                  const double* vel);
int fsb_p2_elasticity(fsb_mesh* mesh, fsb_mat* A, double mu, double lambda);
int fsb_p2_source(fsb_mesh* mesh, double* b, int ncomp, const double* S_const, const double* S_nodal, double scale, const int32_t* d_tags,
                  int tag);
int fsb_p2_facet_load(fsb_mesh* mesh, double* b, int ncomp, int64_t nf, const int32_t* d_fnodes, const int32_t* d_opp, int mode,
                      const double* g, double scale);
int fsb_p2_facet_mass(fsb_mesh* mesh, fsb_mat* A, int64_t nf, const int32_t* d_fnodes, double h);
int fsb_p2_thermal_load(fsb_mesh* mesh, double* b, const double* T, double T_const, double T_ref, double w);
int fsb_p2_von_mises_load(fsb_mesh* mesh, const double* u, double mu, double lambda, double* b);
// collapsed Gauss-Legendre rule on the reference d-simplex, n points per axis: barycentric points [np][d+1] and
// weights normalised to sum 1 (so int f = |T| sum w f)   [fsb_assemble_p2.cu]
void fsb_simplex_rule(int d, int n, std::vector<double>& bary, std::vector<double>& w);
// distributed hooks [fsb_dist.cu]
bool fsb_dist_active(fsb_ctx* ctx);
int fsb_dist_halo_raw(fsb_ctx* ctx, double* v, int64_t n);
int fsb_dist_allreduce_sum_dev(fsb_ctx* ctx, double* d_vals, int count);
void fsb_dist_owned_range(fsb_ctx* ctx, int64_t n, int64_t* o0, int64_t* o1);
void fsb_dist_destroy(fsb_ctx* ctx);
// peer-memory CG support [fsb_dist.cu]: is the IPC mailbox mapped on every rank; collective set-up of the
// shared search-direction vector of a matrix; release; sequence numbers (identical on every rank)
struct PeerComm;
bool fsb_dist_p2p_ready(fsb_ctx* ctx);
int fsb_dist_share_p(fsb_mat* A, int64_t n);
void fsb_dist_release_mat(fsb_mat* A);
int fsb_dist_peer_comm(fsb_mat* A, PeerComm* pc);
unsigned long long fsb_dist_seq_reserve(fsb_ctx* ctx, unsigned long long count);

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__
// matrix scatter: adding an exact zero changes nothing, so such a contribution needs no atomic at all
// (right-angled meshes: 6 of the 16 entries of a P1 tet's Laplace matrix are products of orthogonal gradients)
__device__ __forceinline__ void add_nz(double* p, double v) {
  if (v != 0.0) atomicAdd(p, v);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum, result valid in thread 0 (deterministic order). smem: >= 32 doubles
__device__ __forceinline__ double block_sum(double v, double* smem) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) smem[w] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? smem[threadIdx.x] : 0.0;
  if (w == 0) v = warp_sum(v);
  return v;
}

// von Mises stress of the small-strain state with displacement gradient H = grad u:
// sigma = 2 mu sym(H) + lambda tr(H) I, s = sigma - (1/3) tr(sigma) I, sqrt(3/2 s:s)   (LinearElasticitySolver.py:71-76;
// the 1/3 is the reference's in 2D as well)
template <int D>
__device__ __forceinline__ double fsb_von_mises(const double (&H)[D][D], double mu, double lambda) {
  double tr = 0.0;
#pragma unroll
  for (int i = 0; i < D; ++i) tr += H[i][i];
  double sig[D][D], trs = 0.0;
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) sig[i][j] = mu * (H[i][j] + H[j][i]) + (i == j ? lambda * tr : 0.0);
#pragma unroll
  for (int i = 0; i < D; ++i) trs += sig[i][i];
  double ss = 0.0;
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) {
      const double sij = sig[i][j] - (i == j ? trs * (1.0 / 3.0) : 0.0);
      ss += sij * sij;
    }
  return sqrt(1.5 * ss);
}

// in-row search: position of `col` in the sorted list cols[0..len), starting the search at lo
__device__ __forceinline__ int row_find(const int32_t* __restrict__ cols, int lo, int len, int32_t col) {
  int hi = len;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (__ldg(cols + mid) < col) lo = mid + 1; else hi = mid;
  }
  return lo;
}
#endif
