// K8: distribution — halo exchange and scalar all-reduce over NCCL (NVLink/NVSwitch).  Two layouts:
//  * z-slabs of a box mesh: [ghost plane below | owned planes | ghost plane above], neighbours rank-1 / rank+1
//    (fsb_dist_set_slab; this is the layout the peer-memory CG kernels understand);
//  * any partition of the nodes (fsb_dist_set_halo): local numbering [owned nodes | ghosts grouped by owner rank],
//    per-neighbour send lists packed by a gather kernel, received straight into the ghost range (PETSc VecScatter
//    style; unstructured meshes, degree-2 spaces).
//
// NCCL is bound at run time (dlopen "libnccl.so.2": inside a torch process this resolves to the copy
// torch already loaded) so libfsb.so has no link-time dependency on it; only the handful of entry
// points used here are declared.  Vectors on a distributed context are laid out in natural plane
// order [ghost plane below | owned planes | ghost plane above]; the neighbours are rank-1 and rank+1.
#include "fsb_device.cuh"
#include <dlfcn.h>

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[FSB_NCCL_UID_BYTES]; } ncclUniqueId_t;
enum { kNcclSuccess = 0, kNcclFloat64 = 8, kNcclSum = 0, kNcclMax = 2 };

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(ncclUniqueId_t*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId_t, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};

static NcclApi& nccl() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) return api;
#define LOAD(field, sym) *(void**)(&api.field) = dlsym(api.handle, sym); if (!api.field) return api
  LOAD(GetUniqueId, "ncclGetUniqueId");
  LOAD(CommInitRank, "ncclCommInitRank");
  LOAD(CommDestroy, "ncclCommDestroy");
  LOAD(AllReduce, "ncclAllReduce");
  LOAD(AllGather, "ncclAllGather");
  LOAD(Send, "ncclSend");
  LOAD(Recv, "ncclRecv");
  LOAD(GroupStart, "ncclGroupStart");
  LOAD(GroupEnd, "ncclGroupEnd");
  LOAD(GetErrorString, "ncclGetErrorString");
#undef LOAD
  api.ok = true;
  return api;
}

struct fsb_dist {
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  int ghost_lo = 0, ghost_hi = 0;
  int64_t owned_planes = 0;
  bool slab_set = false;
  // general halo plan (fsb_dist_set_halo)
  bool halo_set = false;
  int64_t n_owned = 0, n_local = 0;
  std::vector<int> nb_rank;
  std::vector<int64_t> send_ptr, recv_off, recv_cnt;     // per neighbour, in nodes
  int64_t* d_send_idx = nullptr;                          // concatenated send lists (local node ids)
  double* d_sendbuf = nullptr;
  size_t sendbuf_cap = 0;                                 // doubles
  // peer memory (CUDA IPC): this rank's CommBuf and every rank's mapping of it
  CommBuf* comm_local = nullptr;
  CommBuf* comm_of[kMaxRanks] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool p2p = false;
  unsigned long long seq = 1;     // next unused sequence number; advances identically on every rank
  // The halo-carrying Krylov vector of the peer-memory path (p / u), exported by CUDA IPC.  It belongs to the CONTEXT: export,
  // import and release of IPC memory cost tens of milliseconds and synchronise the device, so a matrix (one per solver object,
  // one per time-loop restart) borrows this buffer for the length of a solve instead of owning one (measured at 8 GPUs: 60-100 ms
  // spikes per solver construction / release, 36 % of an end-to-end 256^3 step)
  double* p_shared = nullptr;
  long long p_cap = 0;            // doubles
  double* p_lo_map = nullptr;     // rank-1's buffer mapped here
  double* p_hi_map = nullptr;     // rank+1's buffer
  char* d_rec = nullptr;          // device staging of allgather_records
};

#define FSB_CHECK_NCCL(ctx, call)                                                             \
  do {                                                                                        \
    int r__ = (call);                                                                         \
    if (r__ != kNcclSuccess) {                                                                \
      (ctx)->err = std::string(#call) + ": " + (nccl().GetErrorString ? nccl().GetErrorString(r__) : "nccl error"); \
      return FSB_ERR_NCCL;                                                                    \
    }                                                                                         \
  } while (0)

static int fsb_dist_map_mailboxes(fsb_ctx* ctx);

extern "C" int fsb_dist_unique_id(void* uid128) {
  if (!uid128) return FSB_ERR_ARG;
  if (!nccl().ok) return FSB_ERR_NCCL;
  ncclUniqueId_t id;
  if (nccl().GetUniqueId(&id) != kNcclSuccess) return FSB_ERR_NCCL;
  memcpy(uid128, &id, FSB_NCCL_UID_BYTES);
  return FSB_OK;
}

extern "C" int fsb_dist_init(fsb_ctx* ctx, int32_t rank, int32_t nranks, const void* uid128) {
  if (!ctx || !uid128 || nranks < 1 || rank < 0 || rank >= nranks) return FSB_ERR_ARG;
  if (!nccl().ok) FSB_FAIL(ctx, FSB_ERR_NCCL, "libnccl.so.2 could not be loaded");
  if (ctx->dist) FSB_FAIL(ctx, FSB_ERR_STATE, "context is already distributed");
  fsb_dist* d = new fsb_dist();
  d->rank = rank; d->nranks = nranks;
  ncclUniqueId_t id;
  memcpy(&id, uid128, FSB_NCCL_UID_BYTES);
  FSB_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  int r = nccl().CommInitRank(&d->comm, nranks, id, rank);
  if (r != kNcclSuccess) {
    ctx->err = std::string("ncclCommInitRank: ") + nccl().GetErrorString(r);
    delete d;
    return FSB_ERR_NCCL;
  }
  ctx->dist = d;
  // peer-memory mailboxes (best effort: without them the solvers fall back to NCCL collectives)
  if (nranks > 1 && nranks <= kMaxRanks && ctx->dist_p2p) {
    int rc = fsb_dist_map_mailboxes(ctx);
    if (rc) { d->p2p = false; ctx->err.clear(); }
  }
  return FSB_OK;
}

void fsb_dist_destroy(fsb_ctx* ctx) {
  if (!ctx || !ctx->dist) return;
  for (int r = 0; r < kMaxRanks; ++r)
    if (ctx->dist->comm_of[r] && r != ctx->dist->rank) cudaIpcCloseMemHandle(ctx->dist->comm_of[r]);
  cudaFree(ctx->dist->comm_local);
  if (ctx->dist->p_lo_map) cudaIpcCloseMemHandle(ctx->dist->p_lo_map);
  if (ctx->dist->p_hi_map) cudaIpcCloseMemHandle(ctx->dist->p_hi_map);
  cudaFree(ctx->dist->p_shared);
  cudaFree(ctx->dist->d_rec);
  cudaFree(ctx->dist->d_send_idx);
  cudaFree(ctx->dist->d_sendbuf);
  if (ctx->dist->comm && nccl().ok) nccl().CommDestroy(ctx->dist->comm);
  delete ctx->dist;
  ctx->dist = nullptr;
}

bool fsb_dist_active(fsb_ctx* ctx) { return ctx && ctx->dist && ctx->dist->nranks > 1; }

// ------------------------------------------------------------------------------------ peer memory set-up
// Records are exchanged with ncclAllGather (bytes), so no host-side rendezvous beyond NCCL's is needed.
struct ShareRec {
  cudaIpcMemHandle_t handle;    // 64 bytes
  long long ghost_lo, owned, n, plane;
  long long cap;                // capacity of the rank's shared buffer (doubles) when the record was made
  char pad[128 - 64 - 40];
};
static_assert(sizeof(ShareRec) == 128, "ShareRec must be 128 bytes");

static int allgather_records(fsb_ctx* ctx, const ShareRec& mine, std::vector<ShareRec>& all) {
  fsb_dist* d = ctx->dist;
  if (!d->d_rec) FSB_CHECK_CUDA(ctx, cudaMalloc((void**)&d->d_rec, sizeof(ShareRec) * (kMaxRanks + 1)));      // kept: cudaMalloc/cudaFree synchronise
  char* dbuf = d->d_rec;
  FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(dbuf, &mine, sizeof(ShareRec), cudaMemcpyHostToDevice, ctx->stream));
  int r = nccl().AllGather(dbuf, dbuf + sizeof(ShareRec), sizeof(ShareRec), 0 /* ncclInt8 */, d->comm, ctx->stream);
  if (r != kNcclSuccess) FSB_FAIL(ctx, FSB_ERR_NCCL, std::string("ncclAllGather: ") + nccl().GetErrorString(r));
  all.resize(d->nranks);
  FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(all.data(), dbuf + sizeof(ShareRec), sizeof(ShareRec) * d->nranks, cudaMemcpyDeviceToHost, ctx->stream));
  FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FSB_OK;
}

static int fsb_dist_map_mailboxes(fsb_ctx* ctx) {
  fsb_dist* d = ctx->dist;
  FSB_CHECK_CUDA(ctx, cudaMalloc((void**)&d->comm_local, sizeof(CommBuf)));
  FSB_CHECK_CUDA(ctx, cudaMemset(d->comm_local, 0, sizeof(CommBuf)));
  ShareRec mine;
  memset(&mine, 0, sizeof(mine));
  FSB_CHECK_CUDA(ctx, cudaIpcGetMemHandle(&mine.handle, d->comm_local));
  std::vector<ShareRec> all;
  int rc = allgather_records(ctx, mine, all);
  if (rc) return rc;
  for (int r = 0; r < d->nranks; ++r) {
    if (r == d->rank) { d->comm_of[r] = d->comm_local; continue; }
    void* q = nullptr;
    FSB_CHECK_CUDA(ctx, cudaIpcOpenMemHandle(&q, all[r].handle, cudaIpcMemLazyEnablePeerAccess));
    d->comm_of[r] = (CommBuf*)q;
  }
  d->p2p = true;
  return FSB_OK;
}

bool fsb_dist_p2p_ready(fsb_ctx* ctx) { return fsb_dist_active(ctx) && ctx->dist->p2p && ctx->dist->slab_set && ctx->dist_p2p; }

unsigned long long fsb_dist_seq_reserve(fsb_ctx* ctx, unsigned long long count) {
  unsigned long long base = ctx->dist->seq;
  ctx->dist->seq += count;
  return base;
}

// Collective: (re)allocate the IPC-exportable search-direction vector of A and map the neighbours' copies.
static void release_shared_p(fsb_dist* d) {
  if (d->p_lo_map) cudaIpcCloseMemHandle(d->p_lo_map);
  if (d->p_hi_map) cudaIpcCloseMemHandle(d->p_hi_map);
  if (d->p_shared) cudaFree(d->p_shared);
  d->p_lo_map = d->p_hi_map = d->p_shared = nullptr;
  d->p_cap = 0;
}

// Collective.  Lends the context's exported buffer to A for the coming solve.  Every call all-gathers (n, capacity, slab layout) of
// all ranks — one small NCCL all-gather — and every rank takes the same decision from the same data: if any rank's buffer is too
// small (first use, or a larger problem) ALL ranks free, re-allocate, re-export and re-import; otherwise nothing is allocated.  The
// decision never depends on when a rank's garbage collector released an earlier matrix.
int fsb_dist_share_p(fsb_mat* A, int64_t n) {
  fsb_ctx* ctx = A->ctx;
  fsb_dist* d = ctx->dist;
  const long long planes = d->ghost_lo + d->owned_planes + d->ghost_hi;
  ShareRec mine;
  memset(&mine, 0, sizeof(mine));
  mine.ghost_lo = d->ghost_lo; mine.owned = d->owned_planes; mine.n = n; mine.plane = n / planes; mine.cap = d->p_cap;
  std::vector<ShareRec> all;
  int rc = allgather_records(ctx, mine, all);
  if (rc) return rc;
  bool grow = false;
  for (int r = 0; r < d->nranks; ++r) grow = grow || all[r].cap < all[r].n;
  if (grow) {
    release_shared_p(d);
    const long long cap = n + n / 8 + 64;                      // head room: a slightly larger slab does not re-export
    FSB_CHECK_CUDA(ctx, cudaMalloc((void**)&d->p_shared, sizeof(double) * cap + 512));
    d->p_cap = cap;
    FSB_CHECK_CUDA(ctx, cudaIpcGetMemHandle(&mine.handle, d->p_shared));
    mine.cap = cap;
    std::vector<ShareRec> handles;
    if ((rc = allgather_records(ctx, mine, handles))) return rc;
    if (d->rank > 0) {
      void* q = nullptr;
      FSB_CHECK_CUDA(ctx, cudaIpcOpenMemHandle(&q, handles[d->rank - 1].handle, cudaIpcMemLazyEnablePeerAccess));
      d->p_lo_map = (double*)q;
    }
    if (d->rank < d->nranks - 1) {
      void* q = nullptr;
      FSB_CHECK_CUDA(ctx, cudaIpcOpenMemHandle(&q, handles[d->rank + 1].handle, cudaIpcMemLazyEnablePeerAccess));
      d->p_hi_map = (double*)q;
    }
  }
  A->p_dist = d->p_shared;
  A->p_dist_n = n;
  A->p_lo_remote = d->ghost_lo ? d->p_lo_map : nullptr;
  A->p_hi_remote = d->ghost_hi ? d->p_hi_map : nullptr;
  A->lo_ghost_offset = d->ghost_lo ? (all[d->rank - 1].ghost_lo + all[d->rank - 1].owned) * all[d->rank - 1].plane : 0;
  FSB_CHECK_CUDA(ctx, cudaMemsetAsync(A->p_dist, 0, sizeof(double) * n + 512, ctx->stream));
  return FSB_OK;
}

// the buffer is the context's: a matrix only forgets it
void fsb_dist_release_mat(fsb_mat* A) {
  A->p_lo_remote = A->p_hi_remote = A->p_dist = nullptr;
  A->p_dist_n = 0;
}

int fsb_dist_peer_comm(fsb_mat* A, PeerComm* pc) {
  fsb_ctx* ctx = A->ctx;
  fsb_dist* d = ctx->dist;
  memset(pc, 0, sizeof(*pc));
  pc->rank = d->rank; pc->nranks = d->nranks;
  for (int r = 0; r < d->nranks; ++r) pc->buf[r] = d->comm_of[r];
  const long long planes = d->ghost_lo + d->owned_planes + d->ghost_hi;
  pc->plane = A->p_dist_n / planes;
  pc->lo_dst = A->p_lo_remote ? A->p_lo_remote + A->lo_ghost_offset : nullptr;
  pc->hi_dst = A->p_hi_remote ? A->p_hi_remote : nullptr;     // rank+1's lower ghost plane is its plane 0
  return FSB_OK;
}


extern "C" int fsb_dist_set_slab(fsb_ctx* ctx, int32_t ghost_lo, int32_t ghost_hi, int64_t owned_planes) {
  if (!ctx) return FSB_ERR_ARG;
  if (!ctx->dist) FSB_FAIL(ctx, FSB_ERR_STATE, "fsb_dist_init has not been called");
  if (ghost_lo < 0 || ghost_lo > 1 || ghost_hi < 0 || ghost_hi > 1 || owned_planes < 1) FSB_FAIL(ctx, FSB_ERR_ARG, "bad slab description");
  fsb_dist* d = ctx->dist;
  if ((ghost_lo && d->rank == 0) || (ghost_hi && d->rank == d->nranks - 1)) FSB_FAIL(ctx, FSB_ERR_ARG, "ghost plane without a neighbour rank");
  d->ghost_lo = ghost_lo; d->ghost_hi = ghost_hi; d->owned_planes = owned_planes; d->slab_set = true;
  d->halo_set = false;
  return FSB_OK;
}

extern "C" int fsb_dist_set_halo(fsb_ctx* ctx, int64_t n_owned, int64_t n_local, int32_t nneigh, const int32_t* neigh_rank,
                                 const int64_t* send_ptr, const int64_t* send_idx, const int64_t* recv_off, const int64_t* recv_cnt) {
  if (!ctx || n_owned < 0 || n_local < n_owned || nneigh < 0 || (nneigh > 0 && (!neigh_rank || !send_ptr || !recv_off || !recv_cnt)))
    return FSB_ERR_ARG;
  if (!ctx->dist) FSB_FAIL(ctx, FSB_ERR_STATE, "fsb_dist_init has not been called");
  fsb_dist* d = ctx->dist;
  const int64_t nsend = nneigh ? send_ptr[nneigh] : 0;
  if (nsend > 0 && !send_idx) return FSB_ERR_ARG;
  for (int i = 0; i < nneigh; ++i) {
    if (neigh_rank[i] < 0 || neigh_rank[i] >= d->nranks || neigh_rank[i] == d->rank) FSB_FAIL(ctx, FSB_ERR_ARG, "bad neighbour rank");
    if (send_ptr[i + 1] < send_ptr[i] || recv_off[i] < n_owned || recv_off[i] + recv_cnt[i] > n_local)
      FSB_FAIL(ctx, FSB_ERR_ARG, "halo lists out of range");
  }
  for (int64_t k = 0; k < nsend; ++k)
    if (send_idx[k] < 0 || send_idx[k] >= n_owned) FSB_FAIL(ctx, FSB_ERR_ARG, "send list entry is not an owned node");
  d->slab_set = false;
  d->halo_set = true;
  d->n_owned = n_owned; d->n_local = n_local;
  d->nb_rank.assign(neigh_rank, neigh_rank + nneigh);
  d->send_ptr.assign(send_ptr, send_ptr + (nneigh ? nneigh + 1 : 0));
  d->recv_off.assign(recv_off, recv_off + nneigh);
  d->recv_cnt.assign(recv_cnt, recv_cnt + nneigh);
  cudaFree(d->d_send_idx);
  d->d_send_idx = nullptr;
  if (nsend > 0) {
    FSB_CHECK_CUDA(ctx, cudaMalloc((void**)&d->d_send_idx, sizeof(int64_t) * nsend));
    FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(d->d_send_idx, send_idx, sizeof(int64_t) * nsend, cudaMemcpyHostToDevice, ctx->stream));
    FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return FSB_OK;
}

void fsb_dist_owned_range(fsb_ctx* ctx, int64_t n, int64_t* o0, int64_t* o1) {
  *o0 = 0; *o1 = n;
  if (fsb_dist_active(ctx) && ctx->dist->halo_set && ctx->dist->n_local > 0) {
    *o1 = ctx->dist->n_owned * (n / ctx->dist->n_local);      // vectors hold ncomp values per node
    return;
  }
  if (!fsb_dist_active(ctx) || !ctx->dist->slab_set) return;
  fsb_dist* d = ctx->dist;
  const int64_t planes = d->ghost_lo + d->owned_planes + d->ghost_hi;
  const int64_t plane = n / planes;
  *o0 = d->ghost_lo * plane;
  *o1 = *o0 + d->owned_planes * plane;
}

__global__ void k_pack(const double* __restrict__ v, const int64_t* __restrict__ idx, int64_t nsend, int bs, double* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nsend * bs; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = v[idx[i / bs] * bs + i % bs];
}

// general partition: gather the send lists into one buffer, one send + one receive per neighbour; the ghosts of a
// neighbour are contiguous in the local numbering, so the receive lands in place
static int halo_general(fsb_ctx* ctx, double* v, int64_t n) {
  fsb_dist* d = ctx->dist;
  if (d->n_local <= 0 || n % d->n_local) FSB_FAIL(ctx, FSB_ERR_ARG, "vector length is not a multiple of the local node count");
  const int bs = (int)(n / d->n_local);
  const int nn = (int)d->nb_rank.size();
  if (nn == 0) return FSB_OK;
  const int64_t nsend = d->send_ptr[nn];
  if ((size_t)(nsend * bs) > d->sendbuf_cap) {
    cudaFree(d->d_sendbuf);
    d->d_sendbuf = nullptr; d->sendbuf_cap = 0;
    FSB_CHECK_CUDA(ctx, cudaMalloc((void**)&d->d_sendbuf, sizeof(double) * nsend * bs));
    d->sendbuf_cap = (size_t)(nsend * bs);
  }
  if (nsend > 0) {
    k_pack<<<fsb_grid(nsend * bs, 256, (int64_t)ctx->sm_count * 8), 256, 0, ctx->stream>>>(v, d->d_send_idx, nsend, bs, d->d_sendbuf);
    FSB_LAUNCH_CHECK(ctx);
  }
  FSB_CHECK_NCCL(ctx, nccl().GroupStart());
  for (int i = 0; i < nn; ++i) {
    const int64_t cnt = d->send_ptr[i + 1] - d->send_ptr[i];
    if (cnt > 0) FSB_CHECK_NCCL(ctx, nccl().Send(d->d_sendbuf + d->send_ptr[i] * bs, (size_t)(cnt * bs), kNcclFloat64, d->nb_rank[i], d->comm, ctx->stream));
    if (d->recv_cnt[i] > 0) FSB_CHECK_NCCL(ctx, nccl().Recv(v + d->recv_off[i] * bs, (size_t)(d->recv_cnt[i] * bs), kNcclFloat64, d->nb_rank[i], d->comm, ctx->stream));
  }
  FSB_CHECK_NCCL(ctx, nccl().GroupEnd());
  return FSB_OK;
}

// v has (ghost_lo + owned + ghost_hi) planes; the plane size follows from n
static int halo(fsb_ctx* ctx, double* v, int64_t n) {
  fsb_dist* d = ctx->dist;
  if (d->halo_set) return halo_general(ctx, v, n);
  if (!d->slab_set) FSB_FAIL(ctx, FSB_ERR_STATE, "fsb_dist_set_slab has not been called");
  const int64_t planes = d->ghost_lo + d->owned_planes + d->ghost_hi;
  if (n % planes) FSB_FAIL(ctx, FSB_ERR_ARG, "vector length is not a whole number of planes");
  const size_t plane = (size_t)(n / planes);
  if (!d->ghost_lo && !d->ghost_hi) return FSB_OK;
  FSB_CHECK_NCCL(ctx, nccl().GroupStart());
  if (d->ghost_lo) {
    FSB_CHECK_NCCL(ctx, nccl().Recv(v, plane, kNcclFloat64, d->rank - 1, d->comm, ctx->stream));
    FSB_CHECK_NCCL(ctx, nccl().Send(v + plane * d->ghost_lo, plane, kNcclFloat64, d->rank - 1, d->comm, ctx->stream));
  }
  if (d->ghost_hi) {
    FSB_CHECK_NCCL(ctx, nccl().Send(v + plane * (d->ghost_lo + d->owned_planes - 1), plane, kNcclFloat64, d->rank + 1, d->comm, ctx->stream));
    FSB_CHECK_NCCL(ctx, nccl().Recv(v + plane * (d->ghost_lo + d->owned_planes), plane, kNcclFloat64, d->rank + 1, d->comm, ctx->stream));
  }
  FSB_CHECK_NCCL(ctx, nccl().GroupEnd());
  return FSB_OK;
}

int fsb_dist_halo_raw(fsb_ctx* ctx, double* v, int64_t n) {
  if (!fsb_dist_active(ctx)) return FSB_OK;
  return halo(ctx, v, n);
}

extern "C" int fsb_dist_halo(fsb_vec* v) {
  if (!v) return FSB_ERR_ARG;
  if (!fsb_dist_active(v->ctx)) return FSB_OK;
  return halo(v->ctx, v->d, v->n);
}

int fsb_dist_allreduce_sum_dev(fsb_ctx* ctx, double* d_vals, int count) {
  if (!fsb_dist_active(ctx)) return FSB_OK;
  FSB_CHECK_NCCL(ctx, nccl().AllReduce(d_vals, d_vals, (size_t)count, kNcclFloat64, kNcclSum, ctx->dist->comm, ctx->stream));
  return FSB_OK;
}

extern "C" int fsb_dist_allreduce_max(fsb_ctx* ctx, double* value) {
  if (!ctx || !value) return FSB_ERR_ARG;
  if (!fsb_dist_active(ctx)) return FSB_OK;
  double* d = ctx->d_scalars + 40;
  FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(d, value, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  FSB_CHECK_NCCL(ctx, nccl().AllReduce(d, d, 1, kNcclFloat64, kNcclMax, ctx->dist->comm, ctx->stream));
  FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(value, d, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FSB_OK;
}
