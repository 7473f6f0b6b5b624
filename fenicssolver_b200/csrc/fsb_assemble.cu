// K3+K4+K5+K6: P1 element kernels with scatter-add into the block-CSR, facet kernels, Dirichlet rows.
//
// One thread per cell: connectivity is read as one vector load, the (L2-resident) vertex coordinates
// are gathered, the closed-form P1 local matrix is formed in registers (exact for every integrand on
// the hot path) and its entries are added with fire-and-forget fp64 reductions
// (REDG.E.ADD.F64) at positions taken from the uint8 position map (asm_mode 1) or an in-row binary
// search (asm_mode 0).
#include "fsb_internal.cuh"
#include "fsb_p1.cuh"

struct ScalarForm {
  double kscale;
  double K[9];      // row-major DxD conductivity tensor
  double mass;
  double adv;
  double vel[3];
};

template <int D, bool ACTION>
__global__ void __launch_bounds__(128)
k_scalar_form(int64_t ncells, const int32_t* __restrict__ cells, const double* __restrict__ xyz, ScalarForm f,
              const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx, double* __restrict__ vals,
              const uint8_t* __restrict__ posmap, const double* __restrict__ x, double* __restrict__ y) {
  constexpr int NL = D + 1;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncells; c += (int64_t)gridDim.x * blockDim.x) {
    int v[NL];
    load_cell<D>(cells, c, v);
    Geo<D> g;
    p1_geometry<D>(xyz, v, g);
    // KG[b] = K grad phi_b ; vg[b] = vel . grad phi_b
    double KG[NL][D], vg[NL];
#pragma unroll
    for (int b = 0; b < NL; ++b) {
      vg[b] = 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) s += f.K[i * D + j] * g.G[b][j];
        KG[b][i] = s;
        vg[b] += f.vel[i] * g.G[b][i];
      }
    }
    const double kw = f.kscale * g.vol, mw = f.mass * g.vol / (double)(NL * (NL + 1)), aw = f.adv * g.vol / (double)NL;
    double Ke[NL][NL];
#pragma unroll
    for (int a = 0; a < NL; ++a)
#pragma unroll
      for (int b = 0; b < NL; ++b) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) s += g.G[a][i] * KG[b][i];
        Ke[a][b] = kw * s + mw * (a == b ? 2.0 : 1.0) + aw * vg[b];
      }
    if (ACTION) {
      double xl[NL];
#pragma unroll
      for (int b = 0; b < NL; ++b) xl[b] = __ldg(x + v[b]);
#pragma unroll
      for (int a = 0; a < NL; ++a) {
        double s = 0.0;
#pragma unroll
        for (int b = 0; b < NL; ++b) s += Ke[a][b] * xl[b];
        atomicAdd(y + v[a], s);
      }
    } else {
      int64_t base[NL];
      int pos[NL][NL];
      entry_positions<D>(posmap, c, v, row_ptr, col_idx, base, pos);
#pragma unroll
      for (int a = 0; a < NL; ++a)
#pragma unroll
        for (int b = 0; b < NL; ++b) add_nz(vals + base[a] + pos[a][b], Ke[a][b]);
    }
  }
}

// ------------------------------------------------------------------------------------ row-gather form of the same kernel
// asm_mode 2 (opt-in; the A/B below decided against it as the default).  ncu on the scatter kernel above (profiles/assembly_r1.txt): 0.17 of the HBM roofline, bound by the RATE of
// scalar fp64 REDs (9.2 per tet, SM-side issue ~1.3 cycles per lane), DRAM traffic 1.9x the algorithmic bytes because every RED that
// misses L2 fetches its sector and A has to be zero-filled first.  Here the OWNER of a row computes it: one thread per row walks the
// cells around its vertex (the sorted vertex->cell adjacency the symbolic phase keeps), recomputes their affine geometry (4x
// redundant fp64 work — cheap next to the REDs it replaces), accumulates row `a` of every local matrix in thread-private shared
// memory slots addressed by the position map, and writes the finished row once: no atomics, no zero-fill, no read-modify-write of
// A, and the sum order is fixed (ascending cell index), so the assembled matrix is bitwise reproducible run to run.
// MEASURED (tools/asm_ab.py, profiles/asm_ab_r2.txt, 256^3): 10.6 ms against the scatter kernel's 5.4 ms (zero-fill included), the
// matrix-free action 4.7 ms against 2.2 ms: 403 M cell visits of dependent gathers (cell -> 4 coordinates) cost more than the 926 M
// REDs they replace.  Kept for runs that need reproducible sums.
// MODE 0: vals += row, 1: vals = row (caller would otherwise zero A first), 2: y[row] += (row of K_e) . x (matrix-free action).
constexpr int kRowSlots = 32;       // longest row the shared-memory slots hold (host falls back to the scatter kernel beyond)

template <int D, int MODE>
__global__ void __launch_bounds__(128)
k_scalar_rows(int64_t nrows, const int32_t* __restrict__ cells, const double* __restrict__ xyz, ScalarForm f,
              const int64_t* __restrict__ vptr, const int32_t* __restrict__ v2c, const int64_t* __restrict__ row_ptr,
              double* __restrict__ vals, const uint8_t* __restrict__ posmap, const double* __restrict__ x, double* __restrict__ y) {
  constexpr int NL = D + 1;
  __shared__ double acc[MODE == 2 ? 1 : kRowSlots][128];
  const int tid = threadIdx.x;
  for (int64_t base = blockIdx.x * (int64_t)128; base < nrows; base += (int64_t)gridDim.x * 128) {
    const int64_t row = base + tid;
    if (row >= nrows) continue;
    int64_t k0 = 0;
    int len = 0;
    if (MODE != 2) {
      k0 = __ldg(row_ptr + row);
      len = (int)(__ldg(row_ptr + row + 1) - k0);
      for (int s = 0; s < len; ++s) acc[s][tid] = 0.0;
    }
    double sum = 0.0;
    const int64_t p0 = __ldg(vptr + row), p1 = __ldg(vptr + row + 1);
    for (int64_t p = p0; p < p1; ++p) {
      const int64_t c = __ldg(v2c + p);
      int v[NL];
      load_cell<D>(cells, c, v);
      Geo<D> g;
      p1_geometry<D>(xyz, v, g);
      int la = 0;
#pragma unroll
      for (int a = 1; a < NL; ++a) la = (v[a] == (int)row) ? a : la;
      double Ga[D];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double t = g.G[0][i];
#pragma unroll
        for (int a = 1; a < NL; ++a) t = (la == a) ? g.G[a][i] : t;
        Ga[i] = t;
      }
      const double kw = f.kscale * g.vol, mw = f.mass * g.vol / (double)(NL * (NL + 1)), aw = f.adv * g.vol / (double)NL;
      double ke[NL];
#pragma unroll
      for (int b = 0; b < NL; ++b) {
        double s = 0.0, vg = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) {
          double kg = 0.0;
#pragma unroll
          for (int j = 0; j < D; ++j) kg += f.K[i * D + j] * g.G[b][j];
          s += Ga[i] * kg;
          vg += f.vel[i] * g.G[b][i];
        }
        ke[b] = kw * s + mw * (la == b ? 2.0 : 1.0) + aw * vg;
      }
      if (MODE == 2) {
#pragma unroll
        for (int b = 0; b < NL; ++b) sum += ke[b] * __ldg(x + v[b]);
      } else if (D == 3) {
        const unsigned w = __ldg(reinterpret_cast<const unsigned*>(posmap) + c * NL + la);
#pragma unroll
        for (int b = 0; b < NL; ++b) acc[(w >> (8 * b)) & 0xff][tid] += ke[b];
      } else {
#pragma unroll
        for (int b = 0; b < NL; ++b) acc[__ldg(posmap + c * NL * NL + la * NL + b)][tid] += ke[b];
      }
    }
    if (MODE == 2) {
      y[row] += sum;
    } else if (MODE == 1) {
      for (int s = 0; s < len; ++s) vals[k0 + s] = acc[s][tid];
    } else {
      for (int s = 0; s < len; ++s) vals[k0 + s] += acc[s][tid];
    }
  }
}

// ------------------------------------------------------------------------------------ per-warp combine plan (asm_mode 3)
// The scatter kernel is bound by the RATE of scalar fp64 REDs, and the 32 consecutive tetrahedra of a warp hit only ~200 distinct
// matrix slots with their 512 contributions.  Symbolic phase, once per matrix (k_plan_sort<COUNT>, k_plan_sort<FILL>, built on the
// first assembly in this mode): for every chunk of 32 cells the contributions are sorted by destination slot and the plan stores
//     rank[chunk][e][lane]  uint16   position of contribution (lane, e) in the sorted order
//     head[chunk][lane]     uint16   bit i: sorted position lane*16 + i starts a run (bit 0 always: runs do not cross lanes)
//     run_ptr[chunk]        int64    first entry of the chunk in dest[];   dest[run]  uint32  the run's slot in vals
// Numeric phase (k_scalar_form_plan): each lane forms its 16 local entries, stores them to shared memory at `rank`, then walks its
// own 16 sorted positions, adds up each run and issues ONE RED per run whose sum is not exactly zero.
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

// numeric phase: same local matrix as k_scalar_form<3, false> (fsb_assemble.cu), scatter through the plan
__global__ void __launch_bounds__(128)
k_scalar_form_plan(int64_t ncells, const int32_t* __restrict__ cells, const double* __restrict__ xyz, ScalarForm f,
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

static void plan_release(fsb_mat* A) {
  fsb_ctx* ctx = A->ctx;
  fsb_dfree(ctx, A->plan_rank); fsb_dfree(ctx, A->plan_head); fsb_dfree(ctx, A->plan_run_ptr); fsb_dfree(ctx, A->plan_dest);
  A->plan_rank = nullptr; A->plan_head = nullptr; A->plan_run_ptr = nullptr; A->plan_dest = nullptr; A->plan_runs = 0;
}

// plan for (mesh, A): tetrahedra, degree 1, scalar matrix with a position map and fewer than 2^32 slots
static bool plan_path(fsb_ctx* ctx, fsb_mesh* mesh, fsb_mat* A) {
  return ctx->asm_mode == 3 && mesh->degree == 1 && mesh->tdim == 3 && A && A->mesh == mesh && A->bs == 1 && A->posmap &&
         A->nnzb < (int64_t)0xffffffffll;
}

static int plan_build(fsb_mesh* mesh, fsb_mat* A) {
  fsb_ctx* ctx = mesh->ctx;
  if (A->plan_rank) return FSB_OK;
  const int64_t nc = mesh->ncells, nchunk = (nc + 31) / 32;
  const unsigned grid = fsb_grid(nchunk, 4, (int64_t)ctx->sm_count * 16);
  int32_t* count = nullptr;
  int rc;
  if ((rc = fsb_dmalloc(ctx, &count, (size_t)nchunk + 1))) return rc;
  if ((rc = fsb_dmalloc(ctx, &A->plan_run_ptr, (size_t)nchunk + 1))) { fsb_dfree(ctx, count); return rc; }
  k_plan_sort<false><<<grid, 128, 0, ctx->stream>>>(nc, mesh->cells, A->row_ptr, A->posmap, nullptr, nullptr, count, nullptr, nullptr);
  ctx->launches++;
  rc = fsb_exclusive_scan(ctx, count, A->plan_run_ptr, nchunk);
  fsb_dfree(ctx, count);
  if (rc) { plan_release(A); return rc; }
  int64_t nruns = 0;
  FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(&nruns, A->plan_run_ptr + nchunk, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if ((rc = fsb_dmalloc(ctx, &A->plan_rank, (size_t)nchunk * kChunk)) || (rc = fsb_dmalloc(ctx, &A->plan_head, (size_t)nchunk * 32)) ||
      (rc = fsb_dmalloc(ctx, &A->plan_dest, (size_t)nruns + 1))) { plan_release(A); return rc; }
  k_plan_sort<true><<<grid, 128, 0, ctx->stream>>>(nc, mesh->cells, A->row_ptr, A->posmap, A->plan_rank, A->plan_head, nullptr, A->plan_run_ptr, A->plan_dest);
  FSB_LAUNCH_CHECK(ctx);
  A->plan_runs = nruns;
  return FSB_OK;
}

// the row-gather kernels apply to a degree-1 mesh whose adjacency was kept, a position map and rows that fit the slots
static bool rows_path(fsb_ctx* ctx, fsb_mesh* mesh, fsb_mat* A) {
  if (ctx->asm_mode != 2 || mesh->degree != 1 || !mesh->v2c) return false;
  if (A && (A->mesh != mesh || !A->posmap || A->max_row_len > kRowSlots)) return false;
  return fsb_mesh_sort_adjacency(mesh) == FSB_OK;       // fixed (ascending cell) summation order
}

// K_e[(a,i),(b,j)] = |T| ( mu (G_a.G_b d_ij + G_a[j] G_b[i]) + lambda G_a[i] G_b[j] ), DxD blocks
template <int D>
__global__ void __launch_bounds__(128)
k_elasticity(int64_t ncells, const int32_t* __restrict__ cells, const double* __restrict__ xyz, double mu, double lambda,
             const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx, double* __restrict__ vals,
             const uint8_t* __restrict__ posmap) {
  constexpr int NL = D + 1;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncells; c += (int64_t)gridDim.x * blockDim.x) {
    int v[NL];
    load_cell<D>(cells, c, v);
    Geo<D> g;
    p1_geometry<D>(xyz, v, g);
    int64_t base[NL];
    int pos[NL][NL];
    entry_positions<D>(posmap, c, v, row_ptr, col_idx, base, pos);
    const double wmu = mu * g.vol, wl = lambda * g.vol;
#pragma unroll
    for (int a = 0; a < NL; ++a)
#pragma unroll
      for (int b = 0; b < NL; ++b) {
        double gg = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) gg += g.G[a][i] * g.G[b][i];
        double* blk = vals + (base[a] + pos[a][b]) * (D * D);
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
          for (int j = 0; j < D; ++j)
            add_nz(blk + i * D + j, wmu * ((i == j ? gg : 0.0) + g.G[a][j] * g.G[b][i]) + wl * g.G[a][i] * g.G[b][j]);
      }
  }
}

struct Vec3 { double v[3]; };

template <int D>
__global__ void k_source_const(int64_t ncells, const int32_t* __restrict__ cells, const double* __restrict__ xyz,
                               int ncomp, Vec3 S, double scale, const int32_t* __restrict__ tags, int tag,
                               double* __restrict__ b) {
  constexpr int NL = D + 1;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncells; c += (int64_t)gridDim.x * blockDim.x) {
    if (tags && tags[c] != tag) continue;
    int v[NL];
    load_cell<D>(cells, c, v);
    Geo<D> g;
    p1_geometry<D>(xyz, v, g);
    const double w = scale * g.vol / (double)NL;
    for (int a = 0; a < NL; ++a)
      for (int k = 0; k < ncomp; ++k) atomicAdd(b + (int64_t)v[a] * ncomp + k, w * S.v[k]);
  }
}

template <int D>
__global__ void k_source_nodal(int64_t ncells, const int32_t* __restrict__ cells, const double* __restrict__ xyz,
                               int ncomp, const double* __restrict__ S, double scale, double* __restrict__ b) {
  constexpr int NL = D + 1;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncells; c += (int64_t)gridDim.x * blockDim.x) {
    int v[NL];
    load_cell<D>(cells, c, v);
    Geo<D> g;
    p1_geometry<D>(xyz, v, g);
    const double w = scale * g.vol / (double)(NL * (NL + 1));
    for (int k = 0; k < ncomp; ++k) {
      double sl[NL], tot = 0.0;
      for (int e = 0; e < NL; ++e) { sl[e] = S[(int64_t)v[e] * ncomp + k]; tot += sl[e]; }
      for (int a = 0; a < NL; ++a) atomicAdd(b + (int64_t)v[a] * ncomp + k, w * (tot + sl[a]));
    }
  }
}

template <int D>
__global__ void k_facet_load(int64_t nf, const int32_t* __restrict__ fverts, const int32_t* __restrict__ opp,
                             const double* __restrict__ xyz, int ncomp, int mode, Vec3 gval, double scale,
                             double* __restrict__ b) {
  for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nf; f += (int64_t)gridDim.x * blockDim.x) {
    const int32_t* fv = fverts + f * D;
    double n[3], x0[3];
    const double meas = facet_geom<D>(xyz, fv, n, x0);
    double gl[3] = {gval.v[0], gval.v[1], gval.v[2]};
    if (mode == 1) {
      double nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
      double d = 0.0;
      for (int i = 0; i < D; ++i) d += n[i] * (xyz[(int64_t)opp[f] * D + i] - x0[i]);
      double sgn = d > 0 ? -1.0 : 1.0;
      for (int i = 0; i < D; ++i) gl[i] = gval.v[0] * sgn * n[i] / nn;
    }
    const double w = scale * meas / (double)D;
    for (int a = 0; a < D; ++a)
      for (int k = 0; k < ncomp; ++k) atomicAdd(b + (int64_t)fv[a] * ncomp + k, w * gl[k]);
  }
}

template <int D>
__global__ void k_facet_mass(int64_t nf, const int32_t* __restrict__ fverts, const double* __restrict__ xyz, double h,
                             const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx,
                             double* __restrict__ vals) {
  for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nf; f += (int64_t)gridDim.x * blockDim.x) {
    const int32_t* fv = fverts + f * D;
    double n[3], x0[3];
    const double w = h * facet_geom<D>(xyz, fv, n, x0) / (double)(D * (D + 1));
    for (int a = 0; a < D; ++a) {
      const int64_t base = row_ptr[fv[a]];
      const int len = (int)(row_ptr[fv[a] + 1] - base);
      for (int b = 0; b < D; ++b) {
        int p = row_find(col_idx + base, 0, len, fv[b]);
        atomicAdd(vals + base + p, w * (a == b ? 2.0 : 1.0));
      }
    }
  }
}

template <int D>
__global__ void k_facet_area(int64_t nf, const int32_t* __restrict__ fverts, int stride, const double* __restrict__ xyz,
                             double* __restrict__ partials) {
  __shared__ double sm[32];
  double s = 0.0;
  for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nf; f += (int64_t)gridDim.x * blockDim.x) {
    double n[3], x0[3];
    s += facet_geom<D>(xyz, fverts + f * stride, n, x0);     // the facet's vertices come first in its node list
  }
  s = block_sum(s, sm);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

// ------------------------------------------------------------------------------------ thermal stress, von Mises, radiation
// b[(a,i)] += w |T| mean_a(T_a - T_ref) G_a[i]: the load of sigma_t = beta (T - T_ref) I tested with grad v
// (LinearElasticitySolver.py:78-85, 232-238), T the P1 interpolant of the nodal temperatures (or a constant)
template <int D>
__global__ void k_thermal_load(int64_t ncells, const int32_t* __restrict__ cells, const double* __restrict__ xyz,
                               const double* __restrict__ T, double T_const, double T_ref, double w, double* __restrict__ b) {
  constexpr int NL = D + 1;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncells; c += (int64_t)gridDim.x * blockDim.x) {
    int v[NL];
    load_cell<D>(cells, c, v);
    Geo<D> g;
    p1_geometry<D>(xyz, v, g);
    double dT = 0.0;
#pragma unroll
    for (int a = 0; a < NL; ++a) dT += (T ? __ldg(T + v[a]) : T_const) - T_ref;
    const double f = w * g.vol * dT / (double)NL;
#pragma unroll
    for (int a = 0; a < NL; ++a)
#pragma unroll
      for (int i = 0; i < D; ++i) atomicAdd(b + (int64_t)v[a] * D + i, f * g.G[a][i]);
  }
}

// b_a += |T|/(D+1) vm(u_h): load of the L2 projection of the (cell-wise constant) von Mises stress onto P1
// (LinearElasticitySolver.py:71-76)
template <int D>
__global__ void k_von_mises_load(int64_t ncells, const int32_t* __restrict__ cells, const double* __restrict__ xyz,
                                 const double* __restrict__ u, double mu, double lambda, double* __restrict__ b) {
  constexpr int NL = D + 1;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncells; c += (int64_t)gridDim.x * blockDim.x) {
    int v[NL];
    load_cell<D>(cells, c, v);
    Geo<D> g;
    p1_geometry<D>(xyz, v, g);
    double H[D][D];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int k = 0; k < D; ++k) H[i][k] = 0.0;
#pragma unroll
    for (int a = 0; a < NL; ++a)
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const double ua = __ldg(u + (int64_t)v[a] * D + i);
#pragma unroll
        for (int k = 0; k < D; ++k) H[i][k] += ua * g.G[a][k];
      }
    const double w = g.vol / (double)NL * fsb_von_mises<D>(H, mu, lambda);
#pragma unroll
    for (int a = 0; a < NL; ++a) atomicAdd(b + v[a], w);
  }
}

// Newton terms of the grey-body boundary flux m (Ta^4 - T^4) (ScalarTransportSolver.py:334-359, 361-374):
//   A_ab += int 4 m T_h^3 phi_a phi_b ds      r_a += rscale int m (T_h^4 - Ta^4) phi_a ds
// T_h is linear on the facet, so the integrands are polynomials of degree 5: the rule integrates them exactly.
struct FacetRule { int n; double l[16][3]; double w[16]; };

template <int D>
__global__ void k_facet_radiation(int64_t nf, const int32_t* __restrict__ fverts, const double* __restrict__ xyz,
                                  const double* __restrict__ T, double m, double Ta4, double rscale, FacetRule q,
                                  const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx,
                                  double* __restrict__ vals, double* __restrict__ r) {
  for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nf; f += (int64_t)gridDim.x * blockDim.x) {
    const int32_t* fv = fverts + f * D;
    double n[3], x0[3];
    const double meas = facet_geom<D>(xyz, fv, n, x0);
    double t[D], Jab[D][D], ra[D];
    for (int a = 0; a < D; ++a) {
      t[a] = T[fv[a]];
      ra[a] = 0.0;
      for (int b = 0; b < D; ++b) Jab[a][b] = 0.0;
    }
    for (int p = 0; p < q.n; ++p) {
      double Tq = 0.0;
      for (int a = 0; a < D; ++a) Tq += t[a] * q.l[p][a];
      const double T3 = Tq * Tq * Tq, wp = q.w[p] * meas;
      for (int a = 0; a < D; ++a) {
        ra[a] += wp * m * (T3 * Tq - Ta4) * q.l[p][a];
        for (int b = 0; b < D; ++b) Jab[a][b] += wp * 4.0 * m * T3 * q.l[p][a] * q.l[p][b];
      }
    }
    for (int a = 0; a < D; ++a) {
      if (r) atomicAdd(r + fv[a], rscale * ra[a]);
      if (vals) {
        const int64_t base = row_ptr[fv[a]];
        const int len = (int)(row_ptr[fv[a] + 1] - base);
        for (int b = 0; b < D; ++b) atomicAdd(vals + base + row_find(col_idx + base, 0, len, fv[b]), Jab[a][b]);
      }
    }
  }
}

// Temperature-dependent conductivity k(T) (ScalarTransportSolver.py:228-233, 284-285 with `conductivity` a function of T;
// examples/test_heat_transfer.py:53-56): k_h is the P1 interpolant of the nodal values k(T_a), so int k_h = |T| mean k_a.
//   r_a  += rscale * w |T| kbar (G_a . grad T_h)
//   A_ab += w |T| ( kbar G_a.G_b + k'(T_b)/(D+1) (G_a . grad T_h) )        (the Gateaux derivative, :352-353)
template <int D>
__global__ void __launch_bounds__(128)
k_scalar_nonlinear_k(int64_t ncells, const int32_t* __restrict__ cells, const double* __restrict__ xyz,
                     const double* __restrict__ T, const double* __restrict__ kn, const double* __restrict__ dkn, double w,
                     double rscale, const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx,
                     double* __restrict__ vals, const uint8_t* __restrict__ posmap, double* __restrict__ r) {
  constexpr int NL = D + 1;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncells; c += (int64_t)gridDim.x * blockDim.x) {
    int v[NL];
    load_cell<D>(cells, c, v);
    Geo<D> g;
    p1_geometry<D>(xyz, v, g);
    double gT[D], kbar = 0.0, dk[NL];
#pragma unroll
    for (int i = 0; i < D; ++i) gT[i] = 0.0;
#pragma unroll
    for (int a = 0; a < NL; ++a) {
      const double Ta = __ldg(T + v[a]);
      kbar += __ldg(kn + v[a]);
      dk[a] = __ldg(dkn + v[a]) / (double)NL;
#pragma unroll
      for (int i = 0; i < D; ++i) gT[i] += Ta * g.G[a][i];
    }
    kbar /= (double)NL;
    const double wv = w * g.vol;
    double flux[NL];
#pragma unroll
    for (int a = 0; a < NL; ++a) {
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i) s += g.G[a][i] * gT[i];
      flux[a] = s;
      if (r) atomicAdd(r + v[a], rscale * wv * kbar * s);
    }
    if (vals) {
      int64_t base[NL];
      int pos[NL][NL];
      entry_positions<D>(posmap, c, v, row_ptr, col_idx, base, pos);
#pragma unroll
      for (int a = 0; a < NL; ++a)
#pragma unroll
        for (int b = 0; b < NL; ++b) {
          double s = 0.0;
#pragma unroll
          for (int i = 0; i < D; ++i) s += g.G[a][i] * g.G[b][i];
          add_nz(vals + base[a] + pos[a][b], wv * (kbar * s + dk[b] * flux[a]));
        }
    }
  }
}

// Convection by a velocity FIELD (ScalarTransportSolver.py:130-139 get_convective_velocity_function, :311): v_h is the P1
// interpolant of the nodal velocities, so int (v_h . grad phi_b) phi_a = |T|/((D+1)(D+2)) sum_c (1 + delta_ac) v_c . G_b.
//   A_ab += w * that            or (ACTION)   y_a += w * sum_b (...) x_b
template <int D, bool ACTION>
__global__ void __launch_bounds__(128)
k_advection_nodal(int64_t ncells, const int32_t* __restrict__ cells, const double* __restrict__ xyz, const double* __restrict__ vel,
                  double w, const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx, double* __restrict__ vals,
                  const uint8_t* __restrict__ posmap, const double* __restrict__ x, double* __restrict__ y) {
  constexpr int NL = D + 1;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncells; c += (int64_t)gridDim.x * blockDim.x) {
    int v[NL];
    load_cell<D>(cells, c, v);
    Geo<D> g;
    p1_geometry<D>(xyz, v, g);
    double vn[NL][D], vs[D];
#pragma unroll
    for (int i = 0; i < D; ++i) vs[i] = 0.0;
#pragma unroll
    for (int a = 0; a < NL; ++a)
#pragma unroll
      for (int i = 0; i < D; ++i) { vn[a][i] = __ldg(vel + (int64_t)v[a] * D + i); vs[i] += vn[a][i]; }
    const double f = w * g.vol / (double)(NL * (NL + 1));
    double Ce[NL][NL];
#pragma unroll
    for (int a = 0; a < NL; ++a)
#pragma unroll
      for (int b = 0; b < NL; ++b) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) s += (vs[i] + vn[a][i]) * g.G[b][i];
        Ce[a][b] = f * s;
      }
    if (ACTION) {
      double xl[NL];
#pragma unroll
      for (int b = 0; b < NL; ++b) xl[b] = __ldg(x + v[b]);
#pragma unroll
      for (int a = 0; a < NL; ++a) {
        double s = 0.0;
#pragma unroll
        for (int b = 0; b < NL; ++b) s += Ce[a][b] * xl[b];
        atomicAdd(y + v[a], s);
      }
    } else {
      int64_t base[NL];
      int pos[NL][NL];
      entry_positions<D>(posmap, c, v, row_ptr, col_idx, base, pos);
#pragma unroll
      for (int a = 0; a < NL; ++a)
#pragma unroll
        for (int b = 0; b < NL; ++b) add_nz(vals + base[a] + pos[a][b], Ce[a][b]);
    }
  }
}

// ------------------------------------------------------------------------------------ Dirichlet
__global__ void k_bc_scatter(int64_t nbc, const int64_t* __restrict__ dofs, const double* __restrict__ g,
                             uint8_t* __restrict__ flag, double* __restrict__ val, double* __restrict__ x) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nbc; i += (int64_t)gridDim.x * blockDim.x) {
    flag[dofs[i]] = 1;
    val[dofs[i]] = g[i];
    if (x) x[dofs[i]] = g[i];
  }
}

// one thread per scalar row (R = block row, i = component)
template <int BS>
__global__ void k_dirichlet(int64_t nrows, const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx,
                            double* __restrict__ vals, const uint8_t* __restrict__ flag, const double* __restrict__ gval,
                            double* __restrict__ b, int symmetric) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < nrows; r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t R = r / BS;
    const int i = (int)(r % BS);
    const int64_t k0 = row_ptr[R], k1 = row_ptr[R + 1];
    if (flag[r]) {
      for (int64_t k = k0; k < k1; ++k) {
        const int64_t C = col_idx[k];
        for (int j = 0; j < BS; ++j) vals[k * BS * BS + i * BS + j] = (C * BS + j == r) ? 1.0 : 0.0;
      }
      b[r] = gval[r];
    } else if (symmetric) {
      double corr = 0.0;
      for (int64_t k = k0; k < k1; ++k) {
        const int64_t C = col_idx[k];
        for (int j = 0; j < BS; ++j)
          if (flag[C * BS + j]) {
            double* p = vals + k * BS * BS + i * BS + j;
            corr += *p * gval[C * BS + j];
            *p = 0.0;
          }
      }
      if (corr != 0.0) b[r] -= corr;
    }
  }
}

// ------------------------------------------------------------------------------------ ABI
static void fill_form(ScalarForm& f, int D, double kscale, const double* ktensor, double mass, double adv, const double* vel) {
  memset(&f, 0, sizeof(f));
  f.kscale = kscale; f.mass = mass; f.adv = adv;
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < D; ++j) f.K[i * D + j] = ktensor ? ktensor[i * D + j] : (i == j ? 1.0 : 0.0);
  if (vel) for (int i = 0; i < D; ++i) f.vel[i] = vel[i];
}

static int assemble_scalar_impl(fsb_mesh* mesh, fsb_mat* A, double kscale, const double* ktensor, double mass, double adv, const double* vel,
                                int overwrite) {
  if (!mesh || !A) return FSB_ERR_ARG;
  fsb_ctx* ctx = mesh->ctx;
  if (A->bs != 1 || A->nbrows != mesh->nnodes) FSB_FAIL(ctx, FSB_ERR_ARG, "matrix does not belong to a scalar P1 space on this mesh");
  if (adv != 0.0 && !vel) FSB_FAIL(ctx, FSB_ERR_ARG, "advection needs a velocity");
  ScalarForm f;
  if (mesh->degree == 1 && rows_path(ctx, mesh, A)) {
    fill_form(f, mesh->tdim, kscale, ktensor, mass, adv, vel);
    const unsigned grid = fsb_grid(mesh->nnodes, 128, (int64_t)ctx->sm_count * 64);
#define ROWS_LAUNCH(D, MODE) k_scalar_rows<D, MODE><<<grid, 128, 0, ctx->stream>>>(mesh->nnodes, mesh->cells, mesh->xyz, f, mesh->v2c_ptr, mesh->v2c, \
                                                                                  A->row_ptr, A->vals, A->posmap, nullptr, nullptr)
    if (mesh->tdim == 3) { if (overwrite) ROWS_LAUNCH(3, 1); else ROWS_LAUNCH(3, 0); }
    else { if (overwrite) ROWS_LAUNCH(2, 1); else ROWS_LAUNCH(2, 0); }
#undef ROWS_LAUNCH
    FSB_LAUNCH_CHECK(ctx);
    return FSB_OK;
  }
  if (overwrite) {
    int rc = fsb_mat_zero(A);
    if (rc) return rc;
  }
  if (mesh->degree == 2) return fsb_p2_scalar(mesh, A, nullptr, nullptr, kscale, ktensor, mass, adv, vel);
  fill_form(f, mesh->tdim, kscale, ktensor, mass, adv, vel);
  if (plan_path(ctx, mesh, A)) {
    int rc = plan_build(mesh, A);
    if (rc) return rc;
    const int64_t nchunk = (mesh->ncells + 31) / 32;
    k_scalar_form_plan<<<fsb_grid(nchunk, 4, (int64_t)ctx->sm_count * 64), 128, 0, ctx->stream>>>(mesh->ncells, mesh->cells, mesh->xyz, f, A->plan_rank,
                                                                                             A->plan_head, A->plan_run_ptr, A->plan_dest, A->vals);
    FSB_LAUNCH_CHECK(ctx);
    return FSB_OK;
  }
  const uint8_t* pm = (ctx->asm_mode >= 1 && A->mesh == mesh) ? A->posmap : nullptr;
  const unsigned grid = fsb_grid(mesh->ncells, 128, (int64_t)ctx->sm_count * 64);
  if (mesh->tdim == 3)
    k_scalar_form<3, false><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cells, mesh->xyz, f, A->row_ptr, A->col_idx, A->vals, pm, nullptr, nullptr);
  else
    k_scalar_form<2, false><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cells, mesh->xyz, f, A->row_ptr, A->col_idx, A->vals, pm, nullptr, nullptr);
  FSB_LAUNCH_CHECK(ctx);
  return FSB_OK;
}

extern "C" int fsb_assemble_scalar(fsb_mesh* mesh, fsb_mat* A, double kscale, const double* ktensor, double mass,
                                   double adv, const double* vel) {
  return assemble_scalar_impl(mesh, A, kscale, ktensor, mass, adv, vel, 0);
}

extern "C" int fsb_assemble_scalar_set(fsb_mesh* mesh, fsb_mat* A, double kscale, const double* ktensor, double mass,
                                       double adv, const double* vel) {
  return assemble_scalar_impl(mesh, A, kscale, ktensor, mass, adv, vel, 1);
}

extern "C" int fsb_apply_scalar(fsb_mesh* mesh, fsb_vec* x, fsb_vec* y, double kscale, const double* ktensor, double mass,
                                double adv, const double* vel) {
  if (!mesh || !x || !y) return FSB_ERR_ARG;
  fsb_ctx* ctx = mesh->ctx;
  if (x->n != mesh->nnodes || y->n != mesh->nnodes || x == y) FSB_FAIL(ctx, FSB_ERR_ARG, "vector sizes do not match the mesh");
  if (mesh->degree == 2) return fsb_p2_scalar(mesh, nullptr, x, y, kscale, ktensor, mass, adv, vel);
  ScalarForm f;
  fill_form(f, mesh->tdim, kscale, ktensor, mass, adv, vel);
  if (rows_path(ctx, mesh, nullptr)) {
    const unsigned g2 = fsb_grid(mesh->nnodes, 128, (int64_t)ctx->sm_count * 64);
    if (mesh->tdim == 3)
      k_scalar_rows<3, 2><<<g2, 128, 0, ctx->stream>>>(mesh->nnodes, mesh->cells, mesh->xyz, f, mesh->v2c_ptr, mesh->v2c, nullptr, nullptr, nullptr, x->d, y->d);
    else
      k_scalar_rows<2, 2><<<g2, 128, 0, ctx->stream>>>(mesh->nnodes, mesh->cells, mesh->xyz, f, mesh->v2c_ptr, mesh->v2c, nullptr, nullptr, nullptr, x->d, y->d);
    FSB_LAUNCH_CHECK(ctx);
    return FSB_OK;
  }
  const unsigned grid = fsb_grid(mesh->ncells, 128, (int64_t)ctx->sm_count * 64);
  if (mesh->tdim == 3)
    k_scalar_form<3, true><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cells, mesh->xyz, f, nullptr, nullptr, nullptr, nullptr, x->d, y->d);
  else
    k_scalar_form<2, true><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cells, mesh->xyz, f, nullptr, nullptr, nullptr, nullptr, x->d, y->d);
  FSB_LAUNCH_CHECK(ctx);
  return FSB_OK;
}

extern "C" int fsb_assemble_elasticity(fsb_mesh* mesh, fsb_mat* A, double mu, double lambda) {
  if (!mesh || !A) return FSB_ERR_ARG;
  fsb_ctx* ctx = mesh->ctx;
  if (A->bs != mesh->tdim || A->nbrows != mesh->nnodes) FSB_FAIL(ctx, FSB_ERR_ARG, "elasticity needs a matrix with ncomp == dim on this mesh");
  if (mesh->degree == 2) return fsb_p2_elasticity(mesh, A, mu, lambda);
  const uint8_t* pm = (ctx->asm_mode >= 1 && A->mesh == mesh) ? A->posmap : nullptr;
  const unsigned grid = fsb_grid(mesh->ncells, 128, (int64_t)ctx->sm_count * 64);
  if (mesh->tdim == 3)
    k_elasticity<3><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cells, mesh->xyz, mu, lambda, A->row_ptr, A->col_idx, A->vals, pm);
  else
    k_elasticity<2><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cells, mesh->xyz, mu, lambda, A->row_ptr, A->col_idx, A->vals, pm);
  FSB_LAUNCH_CHECK(ctx);
  return FSB_OK;
}

extern "C" int fsb_assemble_source(fsb_mesh* mesh, fsb_vec* b, int32_t ncomp, const double* S, double scale,
                                   const int32_t* cell_tags, int32_t tag) {
  if (!mesh || !b || !S) return FSB_ERR_ARG;
  fsb_ctx* ctx = mesh->ctx;
  if (ncomp < 1 || ncomp > 3 || b->n != mesh->nnodes * ncomp) FSB_FAIL(ctx, FSB_ERR_ARG, "rhs size does not match mesh*ncomp");
  Vec3 s{{0, 0, 0}};
  for (int k = 0; k < ncomp; ++k) s.v[k] = S[k];
  int32_t* d_tags = nullptr;
  if (cell_tags) {
    int rc = fsb_dmalloc(ctx, &d_tags, (size_t)mesh->ncells);
    if (rc) return rc;
    FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(d_tags, cell_tags, sizeof(int32_t) * mesh->ncells, cudaMemcpyHostToDevice, ctx->stream));
  }
  const unsigned grid = fsb_grid(mesh->ncells, 128, (int64_t)ctx->sm_count * 64);
  if (mesh->degree == 2) {
    int rc = fsb_p2_source(mesh, b->d, ncomp, s.v, nullptr, scale, d_tags, tag);
    if (rc) { fsb_dfree(ctx, d_tags); return rc; }
  } else if (mesh->tdim == 3)
    k_source_const<3><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cells, mesh->xyz, ncomp, s, scale, d_tags, tag, b->d);
  else
    k_source_const<2><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cells, mesh->xyz, ncomp, s, scale, d_tags, tag, b->d);
  FSB_LAUNCH_CHECK(ctx);
  if (d_tags) {
    FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    fsb_dfree(ctx, d_tags);
  }
  return FSB_OK;
}

extern "C" int fsb_assemble_source_nodal(fsb_mesh* mesh, fsb_vec* b, int32_t ncomp, fsb_vec* S, double scale) {
  if (!mesh || !b || !S) return FSB_ERR_ARG;
  fsb_ctx* ctx = mesh->ctx;
  if (ncomp < 1 || ncomp > 3 || b->n != mesh->nnodes * ncomp || S->n != b->n) FSB_FAIL(ctx, FSB_ERR_ARG, "vector sizes do not match mesh*ncomp");
  if (mesh->degree == 2) return fsb_p2_source(mesh, b->d, ncomp, nullptr, S->d, scale, nullptr, 0);
  const unsigned grid = fsb_grid(mesh->ncells, 128, (int64_t)ctx->sm_count * 64);
  if (mesh->tdim == 3)
    k_source_nodal<3><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cells, mesh->xyz, ncomp, S->d, scale, b->d);
  else
    k_source_nodal<2><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cells, mesh->xyz, ncomp, S->d, scale, b->d);
  FSB_LAUNCH_CHECK(ctx);
  return FSB_OK;
}

// facets arrive as host arrays; they are small (boundary only) so a temporary upload per call is fine
struct FacetUpload {
  fsb_ctx* ctx;
  int32_t* fverts = nullptr;
  int32_t* opp = nullptr;
  explicit FacetUpload(fsb_ctx* c) : ctx(c) {}
  ~FacetUpload() { fsb_dfree(ctx, fverts); fsb_dfree(ctx, opp); }
};

static int upload_facets(fsb_mesh* mesh, int64_t nf, const int32_t* fverts, const int32_t* opp, FacetUpload& up) {
  fsb_ctx* ctx = mesh->ctx;
  // nodes per facet: tdim vertices (degree 1) or the facet's P2 nodes, vertices then edges (degree 2)
  const int nfn = mesh->degree == 2 ? mesh->tdim * (mesh->tdim + 1) / 2 : mesh->tdim;
  int rc = fsb_dmalloc(ctx, &up.fverts, (size_t)nf * nfn);
  if (rc) return rc;
  FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(up.fverts, fverts, sizeof(int32_t) * nf * nfn, cudaMemcpyHostToDevice, ctx->stream));
  if (opp) {
    rc = fsb_dmalloc(ctx, &up.opp, (size_t)nf);
    if (rc) return rc;
    FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(up.opp, opp, sizeof(int32_t) * nf, cudaMemcpyHostToDevice, ctx->stream));
  }
  return FSB_OK;
}

extern "C" int fsb_assemble_facet_load(fsb_mesh* mesh, fsb_vec* b, int32_t ncomp, int64_t nf, const int32_t* fverts,
                                       const int32_t* opp, int32_t mode, const double* g, double scale) {
  if (!mesh || !b || !g || (nf > 0 && !fverts)) return FSB_ERR_ARG;
  fsb_ctx* ctx = mesh->ctx;
  if (ncomp < 1 || ncomp > 3 || b->n != mesh->nnodes * ncomp) FSB_FAIL(ctx, FSB_ERR_ARG, "rhs size does not match mesh*ncomp");
  if (mode == 1 && (!opp || ncomp != mesh->tdim)) FSB_FAIL(ctx, FSB_ERR_ARG, "normal loads need opposite vertices and ncomp == dim");
  if (nf == 0) return FSB_OK;
  FacetUpload up(ctx);
  int rc = upload_facets(mesh, nf, fverts, opp, up);
  if (rc) return rc;
  Vec3 gv{{0, 0, 0}};
  for (int k = 0; k < (mode == 1 ? 1 : ncomp); ++k) gv.v[k] = g[k];
  const unsigned grid = fsb_grid(nf, 128, (int64_t)ctx->sm_count * 16);
  if (mesh->degree == 2) {
    rc = fsb_p2_facet_load(mesh, b->d, ncomp, nf, up.fverts, up.opp, mode, g, scale);
    if (rc) return rc;
  } else if (mesh->tdim == 3)
    k_facet_load<3><<<grid, 128, 0, ctx->stream>>>(nf, up.fverts, up.opp, mesh->xyz, ncomp, mode, gv, scale, b->d);
  else
    k_facet_load<2><<<grid, 128, 0, ctx->stream>>>(nf, up.fverts, up.opp, mesh->xyz, ncomp, mode, gv, scale, b->d);
  FSB_LAUNCH_CHECK(ctx);
  FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FSB_OK;
}

extern "C" int fsb_assemble_facet_mass(fsb_mesh* mesh, fsb_mat* A, int64_t nf, const int32_t* fverts, double h) {
  if (!mesh || !A || (nf > 0 && !fverts)) return FSB_ERR_ARG;
  fsb_ctx* ctx = mesh->ctx;
  if (A->bs != 1 || A->nbrows != mesh->nnodes) FSB_FAIL(ctx, FSB_ERR_ARG, "facet mass needs the scalar matrix of this mesh");
  if (nf == 0) return FSB_OK;
  FacetUpload up(ctx);
  int rc = upload_facets(mesh, nf, fverts, nullptr, up);
  if (rc) return rc;
  const unsigned grid = fsb_grid(nf, 128, (int64_t)ctx->sm_count * 16);
  if (mesh->degree == 2) {
    rc = fsb_p2_facet_mass(mesh, A, nf, up.fverts, h);
    if (rc) return rc;
  } else if (mesh->tdim == 3)
    k_facet_mass<3><<<grid, 128, 0, ctx->stream>>>(nf, up.fverts, mesh->xyz, h, A->row_ptr, A->col_idx, A->vals);
  else
    k_facet_mass<2><<<grid, 128, 0, ctx->stream>>>(nf, up.fverts, mesh->xyz, h, A->row_ptr, A->col_idx, A->vals);
  FSB_LAUNCH_CHECK(ctx);
  FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FSB_OK;
}

extern "C" int fsb_facet_area(fsb_mesh* mesh, int64_t nf, const int32_t* fverts, double* area) {
  if (!mesh || !area || (nf > 0 && !fverts)) return FSB_ERR_ARG;
  fsb_ctx* ctx = mesh->ctx;
  *area = 0.0;
  if (nf == 0) return FSB_OK;
  FacetUpload up(ctx);
  int rc = upload_facets(mesh, nf, fverts, nullptr, up);
  if (rc) return rc;
  const unsigned grid = fsb_grid(nf, 256, 256);
  const int stride = mesh->degree == 2 ? mesh->tdim * (mesh->tdim + 1) / 2 : mesh->tdim;
  if (mesh->tdim == 3)
    k_facet_area<3><<<grid, 256, 0, ctx->stream>>>(nf, up.fverts, stride, mesh->xyz, ctx->d_partials);
  else
    k_facet_area<2><<<grid, 256, 0, ctx->stream>>>(nf, up.fverts, stride, mesh->xyz, ctx->d_partials);
  FSB_LAUNCH_CHECK(ctx);
  std::vector<double> part(grid);
  FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(part.data(), ctx->d_partials, sizeof(double) * grid, cudaMemcpyDeviceToHost, ctx->stream));
  FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  double s = 0.0;
  for (double p : part) s += p;
  *area = s;
  return FSB_OK;
}

extern "C" int fsb_assemble_thermal_load(fsb_mesh* mesh, fsb_vec* b, double beta, fsb_vec* T, double T_const, double T_ref,
                                         double scale) {
  if (!mesh || !b) return FSB_ERR_ARG;
  fsb_ctx* ctx = mesh->ctx;
  if (b->n != mesh->nnodes * mesh->tdim) FSB_FAIL(ctx, FSB_ERR_ARG, "thermal load needs a vector rhs with ncomp == dim");
  if (T && T->n != mesh->nnodes) FSB_FAIL(ctx, FSB_ERR_ARG, "temperature must be a nodal scalar field of the same space");
  const double* Tp = T ? T->d : nullptr;
  if (mesh->degree == 2) return fsb_p2_thermal_load(mesh, b->d, Tp, T_const, T_ref, beta * scale);
  const unsigned grid = fsb_grid(mesh->ncells, 128, (int64_t)ctx->sm_count * 64);
  if (mesh->tdim == 3) k_thermal_load<3><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cells, mesh->xyz, Tp, T_const, T_ref, beta * scale, b->d);
  else k_thermal_load<2><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cells, mesh->xyz, Tp, T_const, T_ref, beta * scale, b->d);
  FSB_LAUNCH_CHECK(ctx);
  return FSB_OK;
}

extern "C" int fsb_assemble_von_mises_load(fsb_mesh* mesh, fsb_vec* u, double mu, double lambda, fsb_vec* b) {
  if (!mesh || !u || !b) return FSB_ERR_ARG;
  fsb_ctx* ctx = mesh->ctx;
  if (u->n != mesh->nnodes * mesh->tdim) FSB_FAIL(ctx, FSB_ERR_ARG, "displacement must have ncomp == dim on this mesh");
  if (b->n != mesh->nverts) FSB_FAIL(ctx, FSB_ERR_ARG, "the projection load lives on the vertices (P1)");
  if (mesh->degree == 2) return fsb_p2_von_mises_load(mesh, u->d, mu, lambda, b->d);
  const unsigned grid = fsb_grid(mesh->ncells, 128, (int64_t)ctx->sm_count * 64);
  if (mesh->tdim == 3) k_von_mises_load<3><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cells, mesh->xyz, u->d, mu, lambda, b->d);
  else k_von_mises_load<2><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cells, mesh->xyz, u->d, mu, lambda, b->d);
  FSB_LAUNCH_CHECK(ctx);
  return FSB_OK;
}

extern "C" int fsb_assemble_facet_radiation(fsb_mesh* mesh, fsb_mat* A, fsb_vec* r, fsb_vec* T, int64_t nf, const int32_t* fverts,
                                            double m, double T_ambient, double rscale) {
  if (!mesh || !T || (!A && !r) || (nf > 0 && !fverts)) return FSB_ERR_ARG;
  fsb_ctx* ctx = mesh->ctx;
  if (mesh->degree != 1) FSB_FAIL(ctx, FSB_ERR_ARG, "radiation is implemented for degree-1 spaces");
  if (T->n != mesh->nnodes || (r && r->n != mesh->nnodes) || (A && (A->bs != 1 || A->nbrows != mesh->nnodes)))
    FSB_FAIL(ctx, FSB_ERR_ARG, "radiation needs the scalar matrix / vectors of this mesh");
  if (nf == 0) return FSB_OK;
  FacetUpload up(ctx);
  int rc = upload_facets(mesh, nf, fverts, nullptr, up);
  if (rc) return rc;
  std::vector<double> bary, w;
  const int fd = mesh->tdim - 1;
  fsb_simplex_rule(fd, fd == 1 ? 3 : 4, bary, w);
  FacetRule q;
  memset(&q, 0, sizeof(q));
  q.n = (int)w.size();
  for (int p = 0; p < q.n; ++p) {
    for (int a = 0; a <= fd; ++a) q.l[p][a] = bary[(size_t)p * (fd + 1) + a];
    q.w[p] = w[p];
  }
  const double Ta2 = T_ambient * T_ambient;
  const unsigned grid = fsb_grid(nf, 128, (int64_t)ctx->sm_count * 16);
  if (mesh->tdim == 3)
    k_facet_radiation<3><<<grid, 128, 0, ctx->stream>>>(nf, up.fverts, mesh->xyz, T->d, m, Ta2 * Ta2, rscale, q, A ? A->row_ptr : nullptr,
                                                        A ? A->col_idx : nullptr, A ? A->vals : nullptr, r ? r->d : nullptr);
  else
    k_facet_radiation<2><<<grid, 128, 0, ctx->stream>>>(nf, up.fverts, mesh->xyz, T->d, m, Ta2 * Ta2, rscale, q, A ? A->row_ptr : nullptr,
                                                        A ? A->col_idx : nullptr, A ? A->vals : nullptr, r ? r->d : nullptr);
  FSB_LAUNCH_CHECK(ctx);
  FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FSB_OK;
}

extern "C" int fsb_assemble_scalar_nonlinear_k(fsb_mesh* mesh, fsb_mat* A, fsb_vec* r, fsb_vec* T, fsb_vec* k, fsb_vec* dk,
                                               double scale, double rscale) {
  if (!mesh || !T || !k || !dk || (!A && !r)) return FSB_ERR_ARG;
  fsb_ctx* ctx = mesh->ctx;
  if (mesh->degree != 1) FSB_FAIL(ctx, FSB_ERR_ARG, "temperature-dependent conductivity is implemented for degree-1 spaces");
  const int64_t n = mesh->nnodes;
  if (T->n != n || k->n != n || dk->n != n || (r && r->n != n) || (A && (A->bs != 1 || A->nbrows != n)))
    FSB_FAIL(ctx, FSB_ERR_ARG, "nonlinear conductivity needs the scalar matrix / nodal vectors of this mesh");
  const uint8_t* pm = (A && ctx->asm_mode >= 1 && A->mesh == mesh) ? A->posmap : nullptr;
  const unsigned grid = fsb_grid(mesh->ncells, 128, (int64_t)ctx->sm_count * 64);
  if (mesh->tdim == 3)
    k_scalar_nonlinear_k<3><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cells, mesh->xyz, T->d, k->d, dk->d, scale, rscale,
                                                           A ? A->row_ptr : nullptr, A ? A->col_idx : nullptr, A ? A->vals : nullptr, pm, r ? r->d : nullptr);
  else
    k_scalar_nonlinear_k<2><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cells, mesh->xyz, T->d, k->d, dk->d, scale, rscale,
                                                           A ? A->row_ptr : nullptr, A ? A->col_idx : nullptr, A ? A->vals : nullptr, pm, r ? r->d : nullptr);
  FSB_LAUNCH_CHECK(ctx);
  return FSB_OK;
}

extern "C" int fsb_assemble_advection_nodal(fsb_mesh* mesh, fsb_mat* A, fsb_vec* x, fsb_vec* y, fsb_vec* vel, double scale) {
  if (!mesh || !vel || (!A && (!x || !y))) return FSB_ERR_ARG;
  fsb_ctx* ctx = mesh->ctx;
  if (mesh->degree != 1) FSB_FAIL(ctx, FSB_ERR_ARG, "a nodal velocity field is implemented for degree-1 spaces");
  const int64_t n = mesh->nnodes;
  if (vel->n != n * mesh->tdim) FSB_FAIL(ctx, FSB_ERR_ARG, "the velocity field needs dim values per vertex");
  if (A && (A->bs != 1 || A->nbrows != n)) FSB_FAIL(ctx, FSB_ERR_ARG, "matrix does not belong to a scalar P1 space on this mesh");
  if (!A && (x->n != n || y->n != n || x == y)) FSB_FAIL(ctx, FSB_ERR_ARG, "vector sizes do not match the mesh");
  const uint8_t* pm = (A && ctx->asm_mode >= 1 && A->mesh == mesh) ? A->posmap : nullptr;
  const unsigned grid = fsb_grid(mesh->ncells, 128, (int64_t)ctx->sm_count * 64);
  if (mesh->tdim == 3) {
    if (A) k_advection_nodal<3, false><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cells, mesh->xyz, vel->d, scale, A->row_ptr, A->col_idx, A->vals, pm, nullptr, nullptr);
    else k_advection_nodal<3, true><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cells, mesh->xyz, vel->d, scale, nullptr, nullptr, nullptr, nullptr, x->d, y->d);
  } else {
    if (A) k_advection_nodal<2, false><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cells, mesh->xyz, vel->d, scale, A->row_ptr, A->col_idx, A->vals, pm, nullptr, nullptr);
    else k_advection_nodal<2, true><<<grid, 128, 0, ctx->stream>>>(mesh->ncells, mesh->cells, mesh->xyz, vel->d, scale, nullptr, nullptr, nullptr, nullptr, x->d, y->d);
  }
  FSB_LAUNCH_CHECK(ctx);
  return FSB_OK;
}

extern "C" int fsb_apply_dirichlet(fsb_mat* A, fsb_vec* b, fsb_vec* x, int64_t nbc, const int64_t* dofs,
                                   const double* vals, int32_t symmetric) {
  if (!A || !b || (nbc > 0 && (!dofs || !vals))) return FSB_ERR_ARG;
  fsb_ctx* ctx = A->ctx;
  const int64_t n = A->nbrows * A->bs;
  if (b->n != n || (x && x->n != n)) FSB_FAIL(ctx, FSB_ERR_ARG, "vector sizes do not match the matrix");
  for (int64_t i = 0; i < nbc; ++i)
    if (dofs[i] < 0 || dofs[i] >= n) FSB_FAIL(ctx, FSB_ERR_ARG, "Dirichlet dof out of range");
  if (nbc == 0) {
    // no constrained dof this time: flags left by an earlier call must not survive (the multigrid transfers read them)
    if (A->bc_flag) FSB_CHECK_CUDA(ctx, cudaMemsetAsync(A->bc_flag, 0, (size_t)n, ctx->stream));
    return FSB_OK;
  }
  if (!A->bc_flag) {
    int rc = fsb_dmalloc(ctx, &A->bc_flag, (size_t)n);
    if (!rc) rc = fsb_dmalloc(ctx, &A->bc_val, (size_t)n);
    if (rc) return rc;
  }
  FSB_CHECK_CUDA(ctx, cudaMemsetAsync(A->bc_flag, 0, (size_t)n, ctx->stream));
  if (nbc > A->bc_cap) {      // staging buffers are kept on the matrix: a transient run applies the BCs every step
    fsb_dfree(A->ctx, A->bc_dofs); fsb_dfree(A->ctx, A->bc_vals);
    A->bc_dofs = nullptr; A->bc_vals = nullptr; A->bc_cap = 0;
    int rc = fsb_dmalloc(ctx, &A->bc_dofs, (size_t)nbc);
    if (!rc) rc = fsb_dmalloc(ctx, &A->bc_vals, (size_t)nbc);
    if (rc) return rc;
    A->bc_cap = nbc;
  }
  int64_t* d_dofs = A->bc_dofs;
  double* d_vals = A->bc_vals;
  // the host arrays are pageable and caller-owned: cudaMemcpyAsync stages pageable sources before it returns
  FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(d_dofs, dofs, sizeof(int64_t) * nbc, cudaMemcpyHostToDevice, ctx->stream));
  FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(d_vals, vals, sizeof(double) * nbc, cudaMemcpyHostToDevice, ctx->stream));
  k_bc_scatter<<<fsb_grid(nbc, 256, 4096), 256, 0, ctx->stream>>>(nbc, d_dofs, d_vals, A->bc_flag, A->bc_val, x ? x->d : nullptr);
  ctx->launches++;
  const unsigned grid = fsb_grid(n, 256, (int64_t)ctx->sm_count * 32);
  if (A->bs == 1) k_dirichlet<1><<<grid, 256, 0, ctx->stream>>>(n, A->row_ptr, A->col_idx, A->vals, A->bc_flag, A->bc_val, b->d, symmetric);
  else if (A->bs == 2) k_dirichlet<2><<<grid, 256, 0, ctx->stream>>>(n, A->row_ptr, A->col_idx, A->vals, A->bc_flag, A->bc_val, b->d, symmetric);
  else k_dirichlet<3><<<grid, 256, 0, ctx->stream>>>(n, A->row_ptr, A->col_idx, A->vals, A->bc_flag, A->bc_val, b->d, symmetric);
  ctx->launches++;
  FSB_CHECK_CUDA(ctx, cudaGetLastError());
  return FSB_OK;
}
