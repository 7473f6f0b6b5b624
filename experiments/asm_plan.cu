// EXPERIMENT, NOT PART OF libfsb.so (not listed in fenicssolver_b200/build.py) and NOT YET RUN ON HARDWARE: round 1 ended with the GPU
// budget spent.  CUDA counterpart of tools/asm_plan_prototype.py (which is verified in numpy: identical values, 10 -> 4.33 REDs per
// tet), kept here so that the next round starts from code instead of from a description.  It compiles for sm_100a:
//     nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -I include -I fenicssolver_b200/csrc -c experiments/asm_plan.cu
//
// Per-warp combine plan for the P1 tetrahedron matrix scatter.  The kernel k_scalar_form is bound by the rate of scalar fp64 REDs
// (profiles/assembly_r1.txt); the 32 consecutive tets of a warp hit only ~200 distinct matrix slots with their 512 contributions.
//   symbolic phase, once per mesh (k_plan_sort<COUNT>, k_plan_sort<FILL>):
//     rank[chunk][e][lane]  uint16   position of contribution (lane, e) in the chunk's contributions sorted by destination slot
//     head[chunk][lane]     uint16   bit i: sorted position lane*16 + i starts a run (bit 0 always set: runs do not cross lanes)
//     run_ptr[chunk]        int64    first entry of the chunk in dest[]
//     dest[run]             uint32   destination slot of the run (0xffffffff: padding of the last chunk)
//   numeric phase (k_scalar_form_plan): each lane forms its 16 local entries, stores them to shared memory at rank, then walks its
//   own 16 sorted positions, adds up each run and issues one RED per run.
#include "fsb_internal.cuh"
#include "fsb_p1.cuh"

namespace {

constexpr int kE = 16;                 // local entries of a P1 tet
constexpr int kChunk = 32 * kE;        // contributions per warp
constexpr unsigned kNoSlot = 0xffffffffu;

// sorted positions are stored with one pad word per 16 so that lane j reading [17 j + i] is conflict-free per half-warp
__device__ __forceinline__ int padded(int p) { return p + (p >> 4); }

// One warp per chunk.  Sorts the chunk's 512 (slot, k) keys in shared memory (bitonic) and emits rank/head and either the run count
// (FILL = false) or the destinations (FILL = true).
template <bool FILL>
__global__ void __launch_bounds__(128)
k_plan_sort(int64_t ncells, const int32_t* __restrict__ cells, const int64_t* __restrict__ row_ptr, const uint8_t* __restrict__ posmap,
            uint16_t* __restrict__ rank, uint16_t* __restrict__ head, int32_t* __restrict__ run_count, const int64_t* __restrict__ run_ptr,
            uint32_t* __restrict__ dest) {
  __shared__ unsigned long long s_key[4][kChunk];
  __shared__ uint16_t s_rank[4][kChunk];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t nchunk = (ncells + 31) / 32;
  for (int64_t ch = (int64_t)blockIdx.x * 4 + w; ch < nchunk; ch += (int64_t)gridDim.x * 4) {
    const int64_t c = ch * 32 + lane;
    unsigned long long* key = s_key[w];
    if (c < ncells) {
      int v[4];
      load_cell<3>(cells, c, v);
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(posmap) + c);
      const unsigned pw[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const unsigned long long base = (unsigned long long)__ldg(row_ptr + v[a]);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int k = lane * kE + a * 4 + b;
          key[k] = ((base + ((pw[a] >> (8 * b)) & 0xff)) << 9) | (unsigned long long)k;
        }
      }
    } else {
#pragma unroll
      for (int e = 0; e < kE; ++e) key[lane * kE + e] = ((unsigned long long)kNoSlot << 9) | (unsigned long long)(lane * kE + e);
    }
    __syncwarp();
    for (int k = 2; k <= kChunk; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int t = lane; t < (kChunk >> 1); t += 32) {
          const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1)), hi = lo | j;
          const unsigned long long a = key[lo], b = key[hi];
          if ((a > b) == ((lo & k) == 0)) { key[lo] = b; key[hi] = a; }
        }
        __syncwarp();
      }
    // rank of every contribution, heads and runs of this lane's 16 sorted positions
    for (int p = lane; p < kChunk; p += 32) s_rank[w][(int)(key[p] & 511)] = (uint16_t)p;
    __syncwarp();
    unsigned mask = 0;
    int nruns = 0;
#pragma unroll
    for (int i = 0; i < kE; ++i) {
      const int p = lane * kE + i;
      const bool h = i == 0 || (key[p] >> 9) != (key[p - 1] >> 9);
      mask |= (unsigned)h << i;
      nruns += h;
    }
    // exclusive scan of the run counts over the warp
    int incl = nruns;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int tprev = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += tprev;
    }
    if (!FILL) {
      if (lane == 31) run_count[ch] = incl;
    } else {
#pragma unroll
      for (int e = 0; e < kE; ++e) rank[(ch * kE + e) * 32 + lane] = s_rank[w][lane * kE + e];
      head[ch * 32 + lane] = (uint16_t)mask;
      int64_t run = run_ptr[ch] + (incl - nruns);
#pragma unroll
      for (int i = 0; i < kE; ++i)
        if ((mask >> i) & 1) dest[run++] = (uint32_t)(key[lane * kE + i] >> 9);
    }
    __syncwarp();
  }
}

template __global__ void k_plan_sort<false>(int64_t, const int32_t*, const int64_t*, const uint8_t*, uint16_t*, uint16_t*, int32_t*, const int64_t*, uint32_t*);
template __global__ void k_plan_sort<true>(int64_t, const int32_t*, const int64_t*, const uint8_t*, uint16_t*, uint16_t*, int32_t*, const int64_t*, uint32_t*);

struct ScalarFormP {
  double kscale, K[9], mass, adv, vel[3];
};

// numeric phase: same local matrix as k_scalar_form<3, false> (fsb_assemble.cu), scatter through the plan
__global__ void __launch_bounds__(128)
k_scalar_form_plan(int64_t ncells, const int32_t* __restrict__ cells, const double* __restrict__ xyz, ScalarFormP f,
                   const uint16_t* __restrict__ rank, const uint16_t* __restrict__ head, const int64_t* __restrict__ run_ptr,
                   const uint32_t* __restrict__ dest, double* __restrict__ vals) {
  __shared__ double s_val[4][kChunk + kChunk / 16];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t nchunk = (ncells + 31) / 32;
  for (int64_t ch = (int64_t)blockIdx.x * 4 + w; ch < nchunk; ch += (int64_t)gridDim.x * 4) {
    const int64_t c = ch * 32 + lane;
    double Ke[kE];
#pragma unroll
    for (int e = 0; e < kE; ++e) Ke[e] = 0.0;
    if (c < ncells) {
      int v[4];
      load_cell<3>(cells, c, v);
      Geo<3> g;
      p1_geometry<3>(xyz, v, g);
      double KG[4][3], vg[4];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        vg[b] = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          double s = 0.0;
#pragma unroll
          for (int j = 0; j < 3; ++j) s += f.K[i * 3 + j] * g.G[b][j];
          KG[b][i] = s;
          vg[b] += f.vel[i] * g.G[b][i];
        }
      }
      const double kw = f.kscale * g.vol, mw = f.mass * g.vol / 20.0, aw = f.adv * g.vol / 4.0;
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          double s = 0.0;
#pragma unroll
          for (int i = 0; i < 3; ++i) s += g.G[a][i] * KG[b][i];
          Ke[a * 4 + b] = kw * s + mw * (a == b ? 2.0 : 1.0) + aw * vg[b];
        }
    }
    double* sv = s_val[w];
#pragma unroll
    for (int e = 0; e < kE; ++e) sv[padded(__ldg(rank + (ch * kE + e) * 32 + lane))] = Ke[e];
    const unsigned mask = __ldg(head + ch * 32 + lane);
    const int nruns = __popc(mask);
    int incl = nruns;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int tprev = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += tprev;
    }
    __syncwarp();
    int64_t run = __ldg(run_ptr + ch) + (incl - nruns);
    double acc = 0.0;
    unsigned d = kNoSlot;
#pragma unroll
    for (int i = 0; i < kE; ++i) {
      if ((mask >> i) & 1) {
        if (i > 0 && acc != 0.0 && d != kNoSlot) atomicAdd(vals + d, acc);
        acc = 0.0;
        d = __ldg(dest + run++);
      }
      acc += sv[lane * 17 + i];
    }
    if (acc != 0.0 && d != kNoSlot) atomicAdd(vals + d, acc);
    __syncwarp();
  }
}

}  // namespace

// host side, to be merged into fsb_assemble.cu / fsb_pattern.cu: build the plan lazily on the second assembly of a matrix
// (count pass, fsb_exclusive_scan, fill pass), then launch k_scalar_form_plan instead of k_scalar_form<3, false> when
// ctx->asm_mode == 2 and the matrix has fewer than 2^32 slots.
int fsb_experiment_plan_sizes(int64_t ncells, int64_t* rank_bytes, int64_t* head_bytes) {
  const int64_t nchunk = (ncells + 31) / 32;
  if (rank_bytes) *rank_bytes = nchunk * kChunk * (int64_t)sizeof(uint16_t);
  if (head_bytes) *head_bytes = nchunk * 32 * (int64_t)sizeof(uint16_t);
  return 0;
}
