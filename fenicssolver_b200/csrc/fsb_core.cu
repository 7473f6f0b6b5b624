// Context, vectors and meshes (upload + on-device dolfin-layout box generator).
#include "fsb_internal.cuh"
#include <algorithm>

// ------------------------------------------------------------------------------------ context
extern "C" int fsb_init(int device, void* stream, fsb_ctx** out) {
  if (!out) return FSB_ERR_ARG;
  *out = nullptr;
  fsb_ctx* ctx = new fsb_ctx();
  ctx->device = device;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    // no context to carry the message: report on stderr, the Python host raises SolverError
    fprintf(stderr, "libfsb: cudaSetDevice(%d) failed: %s\n", device, cudaGetErrorString(e));
    delete ctx;
    return FSB_ERR_CUDA;
  }
  if (stream) {
    ctx->stream = (cudaStream_t)stream;
  } else {
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete ctx; return FSB_ERR_CUDA; }
    ctx->own_stream = true;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
  // device buffers come from the default stream-ordered pool; keep freed blocks instead of returning them
  // to the OS, so building a second solver on the same GPU reuses the first one's memory at no cost
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t keep = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) ctx->cache_limit = total_b / 3;   // bound on blocks held for reuse
  bool ok = cudaMalloc((void**)&ctx->d_partials, sizeof(double) * kMaxPartials * 4) == cudaSuccess &&
            cudaMalloc((void**)&ctx->d_scalars, sizeof(double) * 64) == cudaSuccess &&
            cudaMalloc((void**)&ctx->d_counters, sizeof(unsigned) * 16) == cudaSuccess &&
            cudaMalloc((void**)&ctx->d_state, sizeof(int) * 8) == cudaSuccess &&
            cudaMallocHost((void**)&ctx->h_state, sizeof(int) * 16) == cudaSuccess &&
            cudaMallocHost((void**)&ctx->h_pinned, sizeof(double) * 64) == cudaSuccess;
  if (!ok) { fsb_destroy(ctx); return FSB_ERR_NOMEM; }
  cudaMemsetAsync(ctx->d_counters, 0, sizeof(unsigned) * 16, ctx->stream);
  cudaMemsetAsync(ctx->d_scalars, 0, sizeof(double) * 64, ctx->stream);
  cudaMemsetAsync(ctx->d_state, 0, sizeof(int) * 8, ctx->stream);
  *out = ctx;
  return FSB_OK;
}

extern "C" void fsb_destroy(fsb_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  fsb_dist_destroy(ctx);
  fsb_cache_flush(ctx);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(ctx->d_partials);
  cudaFree(ctx->d_scalars);
  cudaFree(ctx->d_counters);
  cudaFree(ctx->d_state);
  cudaFree(ctx->cg_comm);
  cudaFreeHost(ctx->h_state);
  for (int k = 0; k < 2; ++k) {
    if (ctx->h_stage[k]) cudaFreeHost(ctx->h_stage[k]);
    if (ctx->stage_done[k]) cudaEventDestroy(ctx->stage_done[k]);
  }
  cudaFreeHost(ctx->h_pinned);
  for (auto& kv : ctx->host_free) cudaFreeHost(kv.second);
  // blocks still handed out stay mapped: arrays that view them may outlive the context
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" const char* fsb_last_error(fsb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int fsb_sync(fsb_ctx* ctx) {
  if (!ctx) return FSB_ERR_ARG;
  FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FSB_OK;
}

extern "C" int fsb_device_info(fsb_ctx* ctx, int32_t* sm_count, int64_t* free_bytes, int64_t* total_bytes) {
  if (!ctx) return FSB_ERR_ARG;
  size_t f = 0, t = 0;
  FSB_CHECK_CUDA(ctx, cudaMemGetInfo(&f, &t));
  if (sm_count) *sm_count = ctx->sm_count;
  if (free_bytes) *free_bytes = (int64_t)f;
  if (total_bytes) *total_bytes = (int64_t)t;
  return FSB_OK;
}

extern "C" int fsb_set_option(fsb_ctx* ctx, const char* name, int64_t value) {
  if (!ctx || !name) return FSB_ERR_ARG;
  std::string s(name);
  if (s == "asm_mode") ctx->asm_mode = (int)value;
  else if (s == "drop_zeros") ctx->drop_zeros = (int)value;
  else if (s == "cg_variant") ctx->cg_variant = (int)value;
  else if (s == "cg_umode") ctx->cg_umode = (int)value;
  else if (s == "cg_minb") ctx->cg_minb = (int)value;
  else if (s == "vec_skew") ctx->vec_skew = (int)(value < 0 ? 0 : (value > 8192 ? 8192 : (value & ~255ll)));
  else if (s == "cg_debug") ctx->cg_debug = (int)value;
  else if (s == "cg_timeout_s") ctx->cg_timeout_s = value < 1 ? 1 : (int)value;
  else if (s == "spmv_hint") ctx->spmv_hint = (int)value;
  else if (s == "spmv_mode") ctx->spmv_mode = (int)value;
  else if (s == "dist_p2p") ctx->dist_p2p = (int)value;
  else if (s == "spmv_lpr") ctx->spmv_lpr = (int)value;
  else if (s == "spmv_rows") ctx->spmv_rows = (int)value;
  else if (s == "spmv_stages") ctx->spmv_stages = (int)value;
  else if (s == "spmv_flat") ctx->spmv_flat = (int)value;
  else if (s == "profile") ctx->profile = (int)value;
  else if (s == "alloc_cache_mb") { ctx->cache_limit = (size_t)(value < 0 ? 0 : value) << 20; if (!value) fsb_cache_flush(ctx); }
  else if (s == "graph") ctx->use_graph = (int)value;
  else if (s == "check_every") ctx->check_every = value < 1 ? 1 : (int)value;
  else FSB_FAIL(ctx, FSB_ERR_ARG, "unknown option " + s);
  return FSB_OK;
}

extern "C" int64_t fsb_launch_count(fsb_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ------------------------------------------------------------------------------------ vectors
__global__ void k_fill(double* __restrict__ p, int64_t n, double v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void k_axpy(double* __restrict__ y, double a, const double* __restrict__ x, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] += a * x[i];
}

__global__ void k_add_entries(double* __restrict__ v, int64_t n, const int64_t* __restrict__ idx, const double* __restrict__ val) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) atomicAdd(v + idx[i], val[i]);
}

extern "C" int fsb_vec_add_entries(fsb_vec* v, int64_t n, const int64_t* idx, const double* vals) {
  if (!v || (n > 0 && (!idx || !vals))) return FSB_ERR_ARG;
  fsb_ctx* ctx = v->ctx;
  for (int64_t i = 0; i < n; ++i)
    if (idx[i] < 0 || idx[i] >= v->n) FSB_FAIL(ctx, FSB_ERR_ARG, "entry index out of range");
  if (n == 0) return FSB_OK;
  int64_t* d_idx = nullptr;
  double* d_val = nullptr;
  int rc = fsb_dmalloc(ctx, &d_idx, (size_t)n);
  if (!rc) rc = fsb_dmalloc(ctx, &d_val, (size_t)n);
  if (rc) { fsb_dfree(ctx, d_idx); return rc; }
  FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(d_idx, idx, sizeof(int64_t) * n, cudaMemcpyHostToDevice, ctx->stream));
  FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(d_val, vals, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  k_add_entries<<<fsb_grid(n, 128, 1024), 128, 0, ctx->stream>>>(v->d, n, d_idx, d_val);
  FSB_LAUNCH_CHECK(ctx);
  FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));     // idx/vals are caller-owned pageable memory
  fsb_dfree(ctx, d_idx);
  fsb_dfree(ctx, d_val);
  return FSB_OK;
}

extern "C" int fsb_vec_create(fsb_ctx* ctx, int64_t n, fsb_vec** out) {
  if (!ctx || !out || n < 0) return FSB_ERR_ARG;
  fsb_vec* v = new fsb_vec{ctx, n, nullptr};
  int rc = fsb_dmalloc(ctx, &v->d, (size_t)n);
  if (rc) { delete v; return rc; }
  FSB_CHECK_CUDA(ctx, cudaMemsetAsync(v->d, 0, sizeof(double) * n + 64, ctx->stream));
  *out = v;
  return FSB_OK;
}
extern "C" int fsb_vec_fill(fsb_vec* v, double value) {
  if (!v) return FSB_ERR_ARG;
  if (v->n == 0) return FSB_OK;
  k_fill<<<fsb_grid(v->n, 256, 148 * 16), 256, 0, v->ctx->stream>>>(v->d, v->n, value);
  FSB_LAUNCH_CHECK(v->ctx);
  return FSB_OK;
}
extern "C" int fsb_vec_upload(fsb_vec* v, const double* host, int64_t n) {
  if (!v || !host || n != v->n) return FSB_ERR_ARG;
  FSB_CHECK_CUDA(v->ctx, cudaMemcpyAsync(v->d, host, sizeof(double) * n, cudaMemcpyHostToDevice, v->ctx->stream));
  FSB_CHECK_CUDA(v->ctx, cudaStreamSynchronize(v->ctx->stream));
  return FSB_OK;
}
// Page-locked host blocks for results (fsb_host_alloc / fsb_host_free): a released block goes to an exact-size free list
// and is handed to the next request of that size, so the solve of every time step downloads into memory that is already
// pinned and already touched (a fresh pageable array costs a page fault per 4 KB: 26 ms for the 136 MB of a 256^3 field
// against 2.5 ms of DMA).
extern "C" int fsb_host_alloc(fsb_ctx* ctx, int64_t bytes, void** out) {
  if (!ctx || !out || bytes <= 0) return FSB_ERR_ARG;
  auto hit = ctx->host_free.find((size_t)bytes);
  if (hit != ctx->host_free.end()) {
    *out = hit->second;
    ctx->host_free.erase(hit);
    ctx->host_cached -= (size_t)bytes;
  } else {
    void* p = nullptr;
    cudaError_t e = cudaMallocHost(&p, (size_t)bytes);
    if (e != cudaSuccess) {
      cudaGetLastError();
      for (auto& kv : ctx->host_free) cudaFreeHost(kv.second);       // give the cached blocks back and try once more
      ctx->host_free.clear(); ctx->host_cached = 0;
      e = cudaMallocHost(&p, (size_t)bytes);
    }
    if (e != cudaSuccess) { cudaGetLastError(); FSB_FAIL(ctx, FSB_ERR_NOMEM, "cudaMallocHost failed"); }
    *out = p;
  }
  ctx->host_live[*out] = (size_t)bytes;
  return FSB_OK;
}

extern "C" int fsb_host_free(fsb_ctx* ctx, void* p) {
  if (!ctx || !p) return FSB_ERR_ARG;
  auto it = ctx->host_live.find(p);
  if (it == ctx->host_live.end()) FSB_FAIL(ctx, FSB_ERR_ARG, "fsb_host_free: not a block of fsb_host_alloc");
  const size_t bytes = it->second;
  ctx->host_live.erase(it);
  if (ctx->host_cached + bytes <= ((size_t)2 << 30)) {      // at most 2 GB of released host blocks wait for reuse
    ctx->host_free.emplace(bytes, p);
    ctx->host_cached += bytes;
  } else {
    cudaFreeHost(p);
  }
  return FSB_OK;
}

extern "C" int fsb_vec_download(fsb_vec* v, double* host, int64_t n) {
  if (!v || !host || n != v->n) return FSB_ERR_ARG;
  fsb_ctx* ctx = v->ctx;
  const size_t bytes = sizeof(double) * (size_t)n;
  constexpr size_t kChunk = 8u << 20;
  bool pinned = false;
  for (auto& kv : ctx->host_live)       // a handful of blocks at most
    if ((char*)host >= (char*)kv.first && (char*)host + bytes <= (char*)kv.first + kv.second) { pinned = true; break; }
  if (pinned || bytes < 4 * kChunk) {
    FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(host, v->d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FSB_OK;
  }
  // Large result into pageable caller memory: DMA into two pinned staging chunks while the host copies the
  // previous chunk out (a direct pageable cudaMemcpy runs at ~2.5 GB/s, this at host-memcpy speed).
  if (!ctx->h_stage[0]) {
    for (int k = 0; k < 2; ++k) FSB_CHECK_CUDA(ctx, cudaMallocHost(&ctx->h_stage[k], kChunk));
    for (int k = 0; k < 2; ++k) FSB_CHECK_CUDA(ctx, cudaEventCreateWithFlags(&ctx->stage_done[k], cudaEventDisableTiming));
  }
  const size_t nchunks = (bytes + kChunk - 1) / kChunk;
  const char* src = reinterpret_cast<const char*>(v->d);
  char* dst = reinterpret_cast<char*>(host);
  for (size_t c = 0; c <= nchunks; ++c) {
    if (c < nchunks) {
      const size_t len = std::min(kChunk, bytes - c * kChunk);
      FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(ctx->h_stage[c & 1], src + c * kChunk, len, cudaMemcpyDeviceToHost, ctx->stream));
      FSB_CHECK_CUDA(ctx, cudaEventRecord(ctx->stage_done[c & 1], ctx->stream));
    }
    if (c > 0) {
      const size_t p = c - 1, len = std::min(kChunk, bytes - p * kChunk);
      FSB_CHECK_CUDA(ctx, cudaEventSynchronize(ctx->stage_done[p & 1]));
      memcpy(dst + p * kChunk, ctx->h_stage[p & 1], len);
    }
  }
  return FSB_OK;
}
extern "C" int fsb_vec_copy(fsb_vec* dst, fsb_vec* src) {
  if (!dst || !src || dst->n != src->n) return FSB_ERR_ARG;
  FSB_CHECK_CUDA(dst->ctx, cudaMemcpyAsync(dst->d, src->d, sizeof(double) * src->n, cudaMemcpyDeviceToDevice, dst->ctx->stream));
  return FSB_OK;
}
extern "C" int fsb_vec_axpy(fsb_vec* y, double a, fsb_vec* x) {
  if (!y || !x || y->n != x->n) return FSB_ERR_ARG;
  if (y->n == 0) return FSB_OK;
  k_axpy<<<fsb_grid(y->n, 256, 148 * 16), 256, 0, y->ctx->stream>>>(y->d, a, x->d, y->n);
  FSB_LAUNCH_CHECK(y->ctx);
  return FSB_OK;
}
extern "C" int fsb_vec_size(fsb_vec* v, int64_t* n) {
  if (!v || !n) return FSB_ERR_ARG;
  *n = v->n;
  return FSB_OK;
}
extern "C" void* fsb_vec_ptr(fsb_vec* v) { return v ? (void*)v->d : nullptr; }
extern "C" void fsb_vec_destroy(fsb_vec* v) {
  if (!v) return;
  fsb_dfree(v->ctx, v->d);
  delete v;
}

// ------------------------------------------------------------------------------------ meshes
__global__ void k_shift_i32(int32_t* __restrict__ p, int64_t n, int32_t delta) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] += delta;
}

extern "C" int fsb_mesh_upload(fsb_ctx* ctx, int32_t gdim, int32_t tdim, int64_t nverts, const double* xyz,
                               int64_t ncells, const int32_t* cells, fsb_mesh** out) {
  return fsb_mesh_upload_part(ctx, gdim, tdim, nverts, xyz, ncells, cells, 0, out);
}

extern "C" int fsb_mesh_upload_part(fsb_ctx* ctx, int32_t gdim, int32_t tdim, int64_t nverts, const double* xyz,
                                    int64_t ncells, const int32_t* cells, int64_t vertex_offset, fsb_mesh** out) {
  if (!ctx || !out || !xyz || !cells) return FSB_ERR_ARG;
  if (gdim != tdim || (tdim != 2 && tdim != 3)) FSB_FAIL(ctx, FSB_ERR_ARG, "only gdim==tdim in {2,3} is supported");
  if (nverts <= 0 || ncells <= 0 || nverts > 0x7fffffffll) FSB_FAIL(ctx, FSB_ERR_ARG, "bad mesh sizes");
  fsb_mesh* m = new fsb_mesh{ctx, gdim, tdim, nverts, ncells};
  int rc = fsb_dmalloc(ctx, &m->xyz, (size_t)nverts * gdim);
  if (!rc) rc = fsb_dmalloc(ctx, &m->cells, (size_t)ncells * (tdim + 1));
  if (rc) { fsb_mesh_destroy(m); return rc; }
  FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(m->xyz, xyz, sizeof(double) * nverts * gdim, cudaMemcpyHostToDevice, ctx->stream));
  FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(m->cells, cells, sizeof(int32_t) * ncells * (tdim + 1), cudaMemcpyHostToDevice, ctx->stream));
  if (vertex_offset) {      // the caller passed a slice of a global cell table: renumber on the device, not on the host
    k_shift_i32<<<fsb_grid(ncells * (tdim + 1), 256, (int64_t)ctx->sm_count * 16), 256, 0, ctx->stream>>>(m->cells, ncells * (tdim + 1), (int32_t)-vertex_offset);
    ctx->launches++;
  }
  FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  m->degree = 1; m->nl = m->tdim + 1; m->nnodes = m->nverts; m->cell_nodes = m->cells;
  *out = m;
  return FSB_OK;
}

// vertex id = ix + iy*(nx+1) (+ iz*(nx+1)(ny+1)); x = p0 + ix*(p1-p0)/n evaluated in that operation
// order without contraction so the coordinates are bit-identical to the numpy oracle.
template <int D>
__global__ void k_box_coords(double* __restrict__ xyz, int64_t nverts, int nx, int ny, int nz, int layer0,
                             double x0, double y0, double z0, double x1, double y1, double z1) {
  const int64_t px = nx + 1, py = px * (ny + 1);
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nverts; v += (int64_t)gridDim.x * blockDim.x) {
    if (D == 2) {
      int64_t iy = v / px + layer0, ix = v % px;
      xyz[2 * v + 0] = __dadd_rn(x0, __ddiv_rn(__dmul_rn((double)ix, __dsub_rn(x1, x0)), (double)nx));
      xyz[2 * v + 1] = __dadd_rn(y0, __ddiv_rn(__dmul_rn((double)iy, __dsub_rn(y1, y0)), (double)ny));
    } else {
      int64_t iz = v / py + layer0, rem = v % py, iy = rem / px, ix = rem % px;
      xyz[3 * v + 0] = __dadd_rn(x0, __ddiv_rn(__dmul_rn((double)ix, __dsub_rn(x1, x0)), (double)nx));
      xyz[3 * v + 1] = __dadd_rn(y0, __ddiv_rn(__dmul_rn((double)iy, __dsub_rn(y1, y0)), (double)ny));
      xyz[3 * v + 2] = __dadd_rn(z0, __ddiv_rn(__dmul_rn((double)iz, __dsub_rn(z1, z0)), (double)nz));
    }
  }
}

// six tets per hex sharing the v0-v7 diagonal (dolfin BoxMesh), written already sorted:
// (0,1,3,7) (0,1,5,7) (0,4,5,7) (0,2,3,7) (0,4,6,7) (0,2,6,7) in local hex-corner numbering.
__global__ void k_box_cells3(int4* __restrict__ cells, int64_t nhex, int nx, int ny) {
  const int64_t px = nx + 1, py = px * (ny + 1);
  for (int64_t h = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; h < nhex; h += (int64_t)gridDim.x * blockDim.x) {
    int64_t cz = h / ((int64_t)nx * ny), rem = h % ((int64_t)nx * ny), cy = rem / nx, cx = rem % nx;
    int v0 = (int)(cx + cy * px + cz * py);
    int c[8] = {v0, v0 + 1, v0 + (int)px, v0 + (int)px + 1, v0 + (int)py, v0 + (int)py + 1, v0 + (int)(py + px), v0 + (int)(py + px) + 1};
    int4* o = cells + 6 * h;
    o[0] = make_int4(c[0], c[1], c[3], c[7]);
    o[1] = make_int4(c[0], c[1], c[5], c[7]);
    o[2] = make_int4(c[0], c[4], c[5], c[7]);
    o[3] = make_int4(c[0], c[2], c[3], c[7]);
    o[4] = make_int4(c[0], c[4], c[6], c[7]);
    o[5] = make_int4(c[0], c[2], c[6], c[7]);
  }
}

// RectangleMesh diagonal "right": (v0,v1,v3), (v0,v2,v3)
__global__ void k_box_cells2(int32_t* __restrict__ cells, int64_t nquad, int nx) {
  const int64_t px = nx + 1;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nquad; q += (int64_t)gridDim.x * blockDim.x) {
    int64_t cy = q / nx, cx = q % nx;
    int v0 = (int)(cy * px + cx), v1 = v0 + 1, v2 = v0 + (int)px, v3 = v2 + 1;
    int32_t* o = cells + 6 * q;
    o[0] = v0; o[1] = v1; o[2] = v3;
    o[3] = v0; o[4] = v2; o[5] = v3;
  }
}

extern "C" int fsb_mesh_box(fsb_ctx* ctx, int32_t tdim, const int32_t* n, const double* p0, const double* p1,
                            int32_t layer0, int32_t layer1, fsb_mesh** out) {
  if (!ctx || !out || !n || !p0 || !p1) return FSB_ERR_ARG;
  if (tdim != 2 && tdim != 3) FSB_FAIL(ctx, FSB_ERR_ARG, "tdim must be 2 or 3");
  for (int i = 0; i < tdim; ++i)
    if (n[i] < 1) FSB_FAIL(ctx, FSB_ERR_ARG, "box divisions must be >= 1");
  const int nlast = n[tdim - 1];
  if (layer0 < 0 || layer1 > nlast || layer0 >= layer1) FSB_FAIL(ctx, FSB_ERR_ARG, "bad layer range");
  const int nx = n[0], ny = n[1], nz = tdim == 3 ? n[2] : 1;
  const int64_t nl = layer1 - layer0;
  int64_t plane = tdim == 3 ? (int64_t)(nx + 1) * (ny + 1) : (nx + 1);
  int64_t nverts = plane * (nl + 1);
  int64_t nbox = tdim == 3 ? (int64_t)nx * ny * nl : (int64_t)nx * nl;
  int64_t ncells = tdim == 3 ? 6 * nbox : 2 * nbox;
  if (nverts > 0x7fffffffll) FSB_FAIL(ctx, FSB_ERR_ARG, "mesh too large for int32 vertex ids");
  fsb_mesh* m = new fsb_mesh{ctx, tdim, tdim, nverts, ncells};
  int rc = fsb_dmalloc(ctx, &m->xyz, (size_t)nverts * tdim);
  if (!rc) rc = fsb_dmalloc(ctx, &m->cells, (size_t)ncells * (tdim + 1));
  if (rc) { fsb_mesh_destroy(m); return rc; }
  const int cap = ctx->sm_count * 16;
  if (tdim == 3) {
    k_box_coords<3><<<fsb_grid(nverts, 256, cap), 256, 0, ctx->stream>>>(m->xyz, nverts, nx, ny, nz, layer0, p0[0], p0[1], p0[2], p1[0], p1[1], p1[2]);
    FSB_LAUNCH_CHECK(ctx);
    k_box_cells3<<<fsb_grid(nbox, 256, cap), 256, 0, ctx->stream>>>((int4*)m->cells, nbox, nx, ny);
    FSB_LAUNCH_CHECK(ctx);
  } else {
    k_box_coords<2><<<fsb_grid(nverts, 256, cap), 256, 0, ctx->stream>>>(m->xyz, nverts, nx, ny, 1, layer0, p0[0], p0[1], 0.0, p1[0], p1[1], 0.0);
    FSB_LAUNCH_CHECK(ctx);
    k_box_cells2<<<fsb_grid(nbox, 256, cap), 256, 0, ctx->stream>>>(m->cells, nbox, nx);
    FSB_LAUNCH_CHECK(ctx);
  }
  m->degree = 1; m->nl = m->tdim + 1; m->nnodes = m->nverts; m->cell_nodes = m->cells;
  *out = m;
  return FSB_OK;
}

extern "C" int fsb_mesh_sizes(fsb_mesh* m, int32_t* gdim, int32_t* tdim, int64_t* nverts, int64_t* ncells) {
  if (!m) return FSB_ERR_ARG;
  if (gdim) *gdim = m->gdim;
  if (tdim) *tdim = m->tdim;
  if (nverts) *nverts = m->nverts;
  if (ncells) *ncells = m->ncells;
  return FSB_OK;
}

extern "C" int fsb_mesh_download(fsb_mesh* m, double* xyz, int32_t* cells) {
  if (!m) return FSB_ERR_ARG;
  fsb_ctx* ctx = m->ctx;
  if (xyz) FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(xyz, m->xyz, sizeof(double) * m->nverts * m->gdim, cudaMemcpyDeviceToHost, ctx->stream));
  if (cells) FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(cells, m->cells, sizeof(int32_t) * m->ncells * (m->tdim + 1), cudaMemcpyDeviceToHost, ctx->stream));
  FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FSB_OK;
}

// FunctionSpace(mesh, "Lagrange", 2): the mesh plus a degree-2 node layout (vertices, then edge nodes).
extern "C" int fsb_mesh_upload_p2(fsb_ctx* ctx, int32_t gdim, int32_t tdim, int64_t nverts, const double* xyz, int64_t ncells,
                                  const int32_t* cell_nodes, int64_t nnodes, fsb_mesh** out) {
  if (!ctx || !out || !xyz || !cell_nodes) return FSB_ERR_ARG;
  if (gdim != tdim || (tdim != 2 && tdim != 3)) FSB_FAIL(ctx, FSB_ERR_ARG, "only gdim==tdim in {2,3} is supported");
  if (nverts <= 0 || ncells <= 0 || nnodes < nverts || nnodes > 0x7fffffffll) FSB_FAIL(ctx, FSB_ERR_ARG, "bad mesh sizes");
  const int nv_loc = tdim + 1, nl = (tdim + 1) * (tdim + 2) / 2;
  for (int64_t c = 0; c < ncells; ++c)
    for (int a = 0; a < nl; ++a) {
      const int32_t v = cell_nodes[c * nl + a];
      if (v < 0 || v >= nnodes || (a < nv_loc && v >= nverts)) FSB_FAIL(ctx, FSB_ERR_ARG, "cell_nodes entry out of range");
    }
  fsb_mesh* m = new fsb_mesh{ctx, gdim, tdim, nverts, ncells};
  int rc = fsb_dmalloc(ctx, &m->xyz, (size_t)nverts * gdim);
  if (!rc) rc = fsb_dmalloc(ctx, &m->cells, (size_t)ncells * nv_loc);
  if (!rc) rc = fsb_dmalloc(ctx, &m->cell_nodes, (size_t)ncells * nl);
  if (rc) { fsb_mesh_destroy(m); return rc; }
  m->degree = 2; m->nl = nl; m->nnodes = nnodes;
  std::vector<int32_t> verts((size_t)ncells * nv_loc);
  for (int64_t c = 0; c < ncells; ++c)
    for (int a = 0; a < nv_loc; ++a) verts[c * nv_loc + a] = cell_nodes[c * nl + a];
  FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(m->xyz, xyz, sizeof(double) * nverts * gdim, cudaMemcpyHostToDevice, ctx->stream));
  FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(m->cells, verts.data(), sizeof(int32_t) * verts.size(), cudaMemcpyHostToDevice, ctx->stream));
  FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(m->cell_nodes, cell_nodes, sizeof(int32_t) * ncells * nl, cudaMemcpyHostToDevice, ctx->stream));
  FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *out = m;
  return FSB_OK;
}

extern "C" void fsb_mesh_destroy(fsb_mesh* m) {
  if (!m) return;
  fsb_dfree(m->ctx, m->xyz);
  if (m->cell_nodes != m->cells) fsb_dfree(m->ctx, m->cell_nodes);
  fsb_dfree(m->ctx, m->cells);
  fsb_dfree(m->ctx, m->p2_tables);
  fsb_dfree(m->ctx, m->v2c_ptr); fsb_dfree(m->ctx, m->v2c);
  fsb_dfree(m->ctx, m->bg_verts); fsb_dfree(m->ctx, m->bg_finv); fsb_dfree(m->ctx, m->bg_xyz); fsb_dfree(m->ctx, m->bg_mid);
  fsb_dfree(m->ctx, m->bf_verts); fsb_dfree(m->ctx, m->bf_opp); fsb_dfree(m->ctx, m->bf_cell); fsb_dfree(m->ctx, m->bf_id);
  delete m;
}
