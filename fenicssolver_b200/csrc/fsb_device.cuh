// Device-side helpers shared by the SpMV and Krylov translation units: mbarrier / TMA bulk-copy PTX wrappers
// and the deterministic last-CTA reduction of per-CTA partial sums.
#pragma once
#include "fsb_internal.cuh"

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

// the same with an L2 eviction-priority hint: the matrix streams through once per SpMV and should not push the
// gathered vector out of L2
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
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


// as finish_partials; returns true in every thread of the CTA that combined the partials (out[] is written)
template <int NV>
__device__ __forceinline__ bool finish_partials_last(const double (&mine)[NV], double* __restrict__ partials, int stride,
                                                     double* __restrict__ out, unsigned* counter, double* sm) {
  __shared__ bool is_last_f;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) partials[k * stride + blockIdx.x] = mine[k];
    __threadfence();
    unsigned t = atomicAdd(counter, 1u);
    is_last_f = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last_f) return false;
  __threadfence();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double s = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) s += __ldcg(partials + k * stride + i);
    s = block_sum(s, sm);
    if (threadIdx.x == 0) out[k] = s;
  }
  if (threadIdx.x == 0) { *counter = 0; __threadfence(); }
  __syncthreads();
  return true;
}


// ------------------------------------------------------------------------------------ peer memory (one node)
// Distributed CG without collective launches.  Every rank owns a small CommBuf that all ranks map through
// CUDA IPC (NVLink / NVSwitch peer memory):
//   * reductions: the last CTA of the producing kernel stores this rank's partial sums into the mailbox
//     of every rank (peer stores), fenced and tagged with a sequence number; the consuming kernel waits for
//     all ranks' entries in its own mailbox and adds them in rank order, so every rank gets bitwise the same
//     value and takes the same convergence decision;
//   * halo: the p-update kernel writes its boundary planes straight into the neighbours' ghost planes and
//     the last CTA raises a flag in the neighbour's CommBuf; the next SpMV waits for that flag.
// Slots are reused every second iteration; a rank can only overwrite a slot after it has consumed a value
// its peers produce after their own reads of that slot, so two parities are enough.
static constexpr int kMaxRanks = 8;
static constexpr int kMailSlots = 4;      // {p.q, (r.z, z.z)} x iteration parity
enum { MAIL_PQ = 0, MAIL_RZ = 2 };
struct MailEntry { double v[3]; unsigned long long seq; };
struct CommBuf {
  MailEntry mail[kMailSlots][kMaxRanks];
  unsigned long long halo_flag[2];        // [0] raised by rank-1, [1] raised by rank+1
};
struct PeerComm {
  int rank, nranks;                       // nranks <= 1: single GPU, scalars stay in ctx->d_scalars
  CommBuf* buf[kMaxRanks];                // buf[r]: rank r's CommBuf (peer mapping); buf[rank] is local
  double* lo_dst;                         // rank-1's upper ghost plane inside its p vector (or null)
  double* hi_dst;                         // rank+1's lower ghost plane inside its p vector (or null)
  long long plane;                        // dofs per vertex plane
};

// what one SpMV launch of the distributed CG waits for and posts
struct fsb_spmv_dist {
  PeerComm pc;
  unsigned long long halo_seq;
  int mail_slot;
  unsigned long long mail_seq;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// last-CTA reduction as finish_partials, but the totals go to every rank's mailbox slot instead of `out`
template <int NV>
__device__ __forceinline__ void finish_partials_mail(const double (&mine)[NV], double* __restrict__ partials, int stride,
                                                     unsigned* counter, double* sm, const PeerComm& pc, int slot,
                                                     unsigned long long seq) {
  __shared__ bool is_last_m;
  __shared__ double tot[3];
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) partials[k * stride + blockIdx.x] = mine[k];
    __threadfence();
    unsigned t = atomicAdd(counter, 1u);
    is_last_m = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last_m) return;
  __threadfence();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double s = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) s += __ldcg(partials + k * stride + i);
    s = block_sum(s, sm);
    if (threadIdx.x == 0) tot[k] = s;
  }
  __syncthreads();
  if ((int)threadIdx.x < pc.nranks) {
    MailEntry* e = &pc.buf[threadIdx.x]->mail[slot][pc.rank];
#pragma unroll
    for (int k = 0; k < NV; ++k) e->v[k] = tot[k];
    __threadfence_system();
    st_release_sys(&e->seq, seq);
  }
  if (threadIdx.x == 0) *counter = 0;
}

// wait for all ranks' entries of `slot` (sequence >= seq) and add them in rank order; every thread gets v[].
// One lane per rank polls (the waits overlap instead of queueing behind each other).
template <int NV>
__device__ __forceinline__ void mail_sum(const PeerComm& pc, int slot, unsigned long long seq, double (&v)[NV], double* /*unused*/) {
  __shared__ double s_mail[kMaxRanks][3];
  if ((int)threadIdx.x < pc.nranks) {
    const MailEntry* e = &pc.buf[pc.rank]->mail[slot][threadIdx.x];
    while (ld_acquire_sys(&e->seq) < seq) { }
#pragma unroll
    for (int k = 0; k < NV; ++k) s_mail[threadIdx.x][k] = *reinterpret_cast<const volatile double*>(&e->v[k]);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double acc = 0.0;
    for (int r = 0; r < pc.nranks; ++r) acc += s_mail[r][k];
    v[k] = acc;
  }
  __syncthreads();
}

// two slots at once (lanes [0, kMaxRanks) poll slot A, lanes [kMaxRanks, 2 kMaxRanks) slot B)
template <int NA, int NB>
__device__ __forceinline__ void mail_sum_pair(const PeerComm& pc, int slot_a, unsigned long long seq_a, double (&a)[NA], int slot_b,
                                              unsigned long long seq_b, double (&b)[NB]) {
  __shared__ double s_pair[2][kMaxRanks][3];
  const int which = threadIdx.x / kMaxRanks, r = threadIdx.x % kMaxRanks;
  if (which < 2 && r < pc.nranks) {
    const MailEntry* e = &pc.buf[pc.rank]->mail[which ? slot_b : slot_a][r];
    const unsigned long long seq = which ? seq_b : seq_a;
    while (ld_acquire_sys(&e->seq) < seq) { }
#pragma unroll
    for (int k = 0; k < 3; ++k) s_pair[which][r][k] = *reinterpret_cast<const volatile double*>(&e->v[k]);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NA; ++k) {
    double acc = 0.0;
    for (int q = 0; q < pc.nranks; ++q) acc += s_pair[0][q][k];
    a[k] = acc;
  }
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    double acc = 0.0;
    for (int q = 0; q < pc.nranks; ++q) acc += s_pair[1][q][k];
    b[k] = acc;
  }
  __syncthreads();
}
