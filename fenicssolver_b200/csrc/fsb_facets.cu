// K1: boundary detection and facet numbering on the device (what mesh.init(tdim-1) + facet.exterior() + the global facet
// numbering do for /root/reference/FenicsSolver/SolverBase.py:229,236,277-283).
//
// A facet is the sorted vertex tuple (a < b [< c]) of a cell minus one vertex; it is exterior when no other cell holds it.
// No global sort: every facet is filed under its smallest vertex a.
//   1. vertex -> cell adjacency (count, scan, fill);
//   2. classify: one thread per cell, for each of its tdim+1 facets walk the cells around a and look for the other
//      holders.  exterior <=> none; "representative" <=> this cell has the smallest index among the holders, so that
//      ucnt[a] counts every distinct facet once (dolfin's facet id = lexicographic rank among ALL facets);
//   3. scan ecnt -> bucket offsets, scan ucnt -> number of distinct facets filed under smaller vertices;
//   4. fill the exterior facets into their buckets (atomic cursor: unordered inside a bucket);
//   5. one thread per boundary vertex: sort its bucket by (b, c) — buckets are in vertex order, so the whole list is now in
//      lexicographic order — and rank each of its exterior facets among the distinct facets of the star of a that have a
//      as smallest vertex: id = ubase[a] + rank.
// Output order and ids are independent of the atomic order (keys are unique), so runs are bitwise reproducible.
#include "fsb_internal.cuh"

namespace {

constexpr int kStarCap = 384;      // distinct-candidate capacity of step 5 (a vertex star of 128 tetrahedra)

__global__ void k_f_v2c_count(const int32_t* __restrict__ cells, int64_t nent, int32_t* __restrict__ deg) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nent; i += (int64_t)gridDim.x * blockDim.x) atomicAdd(deg + cells[i], 1);
}

template <int NL>
__global__ void k_f_v2c_fill(const int32_t* __restrict__ cells, int64_t ncells, const int64_t* __restrict__ vptr, int32_t* __restrict__ cursor,
                             int32_t* __restrict__ v2c) {
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncells; c += (int64_t)gridDim.x * blockDim.x)
#pragma unroll
    for (int a = 0; a < NL; ++a) {
      const int v = cells[c * NL + a];
      v2c[vptr[v] + atomicAdd(cursor + v, 1)] = (int32_t)c;
    }
}

template <int NL>
__device__ __forceinline__ void load_cell(const int32_t* __restrict__ cells, int64_t c, int (&v)[NL]) {
  if (NL == 4) {
    const int4 q = __ldg(reinterpret_cast<const int4*>(cells) + c);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[NL - 1] = q.w;
  } else {
#pragma unroll
    for (int a = 0; a < NL; ++a) v[a] = __ldg(cells + c * NL + a);
  }
}

// facet i of sorted cell v: f[0..NL-1) ascending
template <int NL>
__device__ __forceinline__ void facet_of(const int (&v)[NL], int i, int (&f)[NL - 1]) {
#pragma unroll
  for (int j = 0, k = 0; j < NL; ++j)
    if (j != i) f[k++] = v[j];
}

template <int NL>
__global__ void __launch_bounds__(256)
k_f_classify(const int32_t* __restrict__ cells, int64_t ncells, const int64_t* __restrict__ vptr, const int32_t* __restrict__ v2c,
             uint8_t* __restrict__ mask, int32_t* __restrict__ ecnt, int32_t* __restrict__ ucnt) {
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncells; c += (int64_t)gridDim.x * blockDim.x) {
    int v[NL];
    load_cell<NL>(cells, c, v);
    unsigned m = 0;
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      int f[NL - 1];
      facet_of<NL>(v, i, f);
      const int a = f[0];
      int holders = 0;
      bool smallest = true;
      for (int64_t p = vptr[a]; p < vptr[a + 1]; ++p) {
        const int32_t c2 = v2c[p];
        if (c2 == c) continue;
        int w[NL];
        load_cell<NL>(cells, c2, w);
        bool all = true;
#pragma unroll
        for (int j = 1; j < NL - 1; ++j) {
          bool has = false;
#pragma unroll
          for (int q = 0; q < NL; ++q) has |= (w[q] == f[j]);
          all &= has;
        }
        if (all) {
          ++holders;
          smallest &= (c < c2);
          if (c2 < c) break;          // neither exterior nor the facet's representative: nothing left to learn from the rest of the star
        }
      }
      if (holders == 0) { m |= 1u << i; atomicAdd(ecnt + a, 1); }
      if (smallest) atomicAdd(ucnt + a, 1);
    }
    mask[c] = (uint8_t)m;
  }
}

// exterior facet record inside a bucket: the facet's other vertices, the opposite vertex, the cell
struct FacetRec { int32_t b, c, opp, cell; };

template <int NL>
__global__ void __launch_bounds__(256)
k_f_fill(const int32_t* __restrict__ cells, int64_t ncells, const uint8_t* __restrict__ mask, const int64_t* __restrict__ eptr,
         int32_t* __restrict__ cursor, FacetRec* __restrict__ rec) {
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncells; c += (int64_t)gridDim.x * blockDim.x) {
    const unsigned m = mask[c];
    if (!m) continue;
    int v[NL];
    load_cell<NL>(cells, c, v);
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      if (!(m & (1u << i))) continue;
      int f[NL - 1];
      facet_of<NL>(v, i, f);
      FacetRec r;
      r.b = f[1]; r.c = NL == 4 ? f[NL - 2] : -1; r.opp = v[i]; r.cell = (int32_t)c;
      rec[eptr[f[0]] + atomicAdd(cursor + f[0], 1)] = r;
    }
  }
}

__device__ __forceinline__ unsigned long long key_of(int b, int c) { return ((unsigned long long)(unsigned)b << 32) | (unsigned)(c < 0 ? 0 : c); }

// step 5a: one thread per boundary vertex sorts its bucket by (b, c) and writes the final arrays; fid[] receives the number of
// distinct facets filed under smaller vertices (the rank inside the star of a is added on demand by k_f_rank)
template <int NL>
__global__ void __launch_bounds__(128)
k_f_sort(int64_t nverts, const int64_t* __restrict__ eptr, const int64_t* __restrict__ ubase, FacetRec* __restrict__ rec,
         int32_t* __restrict__ fverts, int32_t* __restrict__ opp, int32_t* __restrict__ cell, int64_t* __restrict__ fid) {
  for (int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; a < nverts; a += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e0 = eptr[a], e1 = eptr[a + 1];
    if (e1 == e0) continue;
    for (int64_t i = e0 + 1; i < e1; ++i) {            // insertion sort of the bucket by (b, c)
      const FacetRec r = rec[i];
      const unsigned long long k = key_of(r.b, r.c);
      int64_t j = i;
      while (j > e0 && key_of(rec[j - 1].b, rec[j - 1].c) > k) { rec[j] = rec[j - 1]; --j; }
      rec[j] = r;
    }
    for (int64_t i = e0; i < e1; ++i) {
      const FacetRec r = rec[i];
      fverts[i * (NL - 1) + 0] = (int32_t)a;
      fverts[i * (NL - 1) + 1] = r.b;
      if (NL == 4) fverts[i * (NL - 1) + (NL - 2)] = r.c;
      opp[i] = r.opp;
      cell[i] = r.cell;
      fid[i] = ubase[a];
    }
  }
}

// step 5b, only when the facet ids are asked for (marker files): one thread per exterior facet (a, b, c) ranks it among the
// distinct facets of the star of a that have a as smallest vertex and adds the rank to fid.  The candidate list lives in local
// memory (3 KB per thread), which is why this is not part of the always-run path.
template <int NL>
__global__ void __launch_bounds__(128)
k_f_rank(const int32_t* __restrict__ cells, int64_t nbf, const int64_t* __restrict__ vptr, const int32_t* __restrict__ v2c,
         const int32_t* __restrict__ fverts, int64_t* __restrict__ fid, int* __restrict__ overflow) {
  for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nbf; f += (int64_t)gridDim.x * blockDim.x) {
    const int a = fverts[f * (NL - 1)];
    const unsigned long long mine = key_of(fverts[f * (NL - 1) + 1], NL == 4 ? fverts[f * (NL - 1) + (NL - 2)] : -1);
    unsigned long long cand[kStarCap];               // distinct keys smaller than mine, sorted
    int n = 0;
    bool over = false;
    for (int64_t p = vptr[a]; p < vptr[a + 1]; ++p) {
      int w[NL];
      load_cell<NL>(cells, v2c[p], w);
      int pa = 0;
#pragma unroll
      for (int q = 0; q < NL; ++q) pa = (w[q] == a) ? q : pa;
      if (pa > 1) continue;                             // a facet of this cell cannot start with a
#pragma unroll
      for (int i = 0; i < NL; ++i) {
        if ((pa == 0) != (i != 0)) continue;            // pa == 0: every facet but facet 0; pa == 1: facet 0 only
        int fc[NL - 1];
        facet_of<NL>(w, i, fc);
        const unsigned long long k = key_of(fc[1], NL == 4 ? fc[NL - 2] : -1);
        if (k >= mine) continue;
        int lo = 0, hi = n;                             // sorted insert, duplicates dropped
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (cand[mid] < k) lo = mid + 1; else hi = mid; }
        if (lo < n && cand[lo] == k) continue;
        if (n == kStarCap) { over = true; continue; }
        for (int j = n; j > lo; --j) cand[j] = cand[j - 1];
        cand[lo] = k;
        ++n;
      }
    }
    if (over) atomicExch(overflow, 1);
    fid[f] += n;
  }
}

template <int NL>
int exterior_facets_impl(fsb_mesh* mesh) {
  fsb_ctx* ctx = mesh->ctx;
  const int64_t nv = mesh->nverts, nc = mesh->ncells;
  const int cap = ctx->sm_count * 16;
  int32_t *deg = nullptr, *v2c = nullptr, *ecnt = nullptr, *ucnt = nullptr;
  int64_t *vptr = nullptr, *eptr = nullptr, *ubase = nullptr;
  uint8_t* mask = nullptr;
  FacetRec* rec = nullptr;
  int* d_over = nullptr;
  int rc = FSB_OK;
  auto cleanup = [&]() {
    if (v2c != mesh->v2c) fsb_dfree(ctx, v2c);          // an adjacency left on the mesh belongs to the mesh
    if (vptr != mesh->v2c_ptr) fsb_dfree(ctx, vptr);
    v2c = nullptr; vptr = nullptr;
    fsb_dfree(ctx, deg); fsb_dfree(ctx, ecnt); fsb_dfree(ctx, ucnt); fsb_dfree(ctx, eptr);
    fsb_dfree(ctx, ubase); fsb_dfree(ctx, mask); fsb_dfree(ctx, rec); fsb_dfree(ctx, d_over);
  };
#define TRY(x) do { rc = (x); if (rc) { cleanup(); return rc; } } while (0)
#define TRYCUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { ctx->err = std::string(#x) + ": " + cudaGetErrorString(e_); cleanup(); return FSB_ERR_CUDA; } } while (0)
  TRY(fsb_dmalloc(ctx, &deg, (size_t)nv + 1));
  // the vertex -> cell adjacency is shared with the symbolic phase of a degree-1 mesh (fsb_mat_create): whoever runs first
  // builds it and leaves it on the mesh
  const bool share_adj = mesh->degree == 1;
  const bool have_adj = share_adj && mesh->v2c != nullptr;
  if (have_adj) { vptr = mesh->v2c_ptr; v2c = mesh->v2c; }
  else {
    TRY(fsb_dmalloc(ctx, &vptr, (size_t)nv + 1));
    TRY(fsb_dmalloc(ctx, &v2c, (size_t)nc * NL));
  }
  TRY(fsb_dmalloc(ctx, &ecnt, (size_t)nv + 1));
  TRY(fsb_dmalloc(ctx, &ucnt, (size_t)nv + 1));
  TRY(fsb_dmalloc(ctx, &eptr, (size_t)nv + 1));
  TRY(fsb_dmalloc(ctx, &ubase, (size_t)nv + 1));
  TRY(fsb_dmalloc(ctx, &mask, (size_t)nc));
  TRY(fsb_dmalloc(ctx, &d_over, 1));
  TRYCUDA(cudaMemsetAsync(deg, 0, sizeof(int32_t) * (nv + 1), ctx->stream));
  TRYCUDA(cudaMemsetAsync(ecnt, 0, sizeof(int32_t) * (nv + 1), ctx->stream));
  TRYCUDA(cudaMemsetAsync(ucnt, 0, sizeof(int32_t) * (nv + 1), ctx->stream));
  TRYCUDA(cudaMemsetAsync(d_over, 0, sizeof(int), ctx->stream));
  if (!have_adj) {
    k_f_v2c_count<<<fsb_grid(nc * NL, 256, cap), 256, 0, ctx->stream>>>(mesh->cells, nc * NL, deg);
    ctx->launches++; TRYCUDA(cudaGetLastError());
    TRY(fsb_exclusive_scan(ctx, deg, vptr, nv));
    TRYCUDA(cudaMemsetAsync(deg, 0, sizeof(int32_t) * (nv + 1), ctx->stream));
    k_f_v2c_fill<NL><<<fsb_grid(nc, 256, cap), 256, 0, ctx->stream>>>(mesh->cells, nc, vptr, deg, v2c);
    ctx->launches++; TRYCUDA(cudaGetLastError());
    if (share_adj) { mesh->v2c_ptr = vptr; mesh->v2c = v2c; mesh->v2c_sorted = false; }
  }
  k_f_classify<NL><<<fsb_grid(nc, 256, cap), 256, 0, ctx->stream>>>(mesh->cells, nc, vptr, v2c, mask, ecnt, ucnt);
  ctx->launches++; TRYCUDA(cudaGetLastError());
  TRY(fsb_exclusive_scan(ctx, ecnt, eptr, nv));
  TRY(fsb_exclusive_scan(ctx, ucnt, ubase, nv));
  int64_t nbf = 0, nfacets = 0;
  TRYCUDA(cudaMemcpyAsync(&nbf, eptr + nv, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  TRYCUDA(cudaMemcpyAsync(&nfacets, ubase + nv, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  TRYCUDA(cudaStreamSynchronize(ctx->stream));
  fsb_dfree(ctx, mesh->bf_verts); fsb_dfree(ctx, mesh->bf_opp); fsb_dfree(ctx, mesh->bf_cell); fsb_dfree(ctx, mesh->bf_id);
  mesh->bf_verts = nullptr; mesh->bf_opp = nullptr; mesh->bf_cell = nullptr; mesh->bf_id = nullptr;
  mesh->nbf = -1;
  TRY(fsb_dmalloc(ctx, &rec, (size_t)nbf + 1));
  TRY(fsb_dmalloc(ctx, &mesh->bf_verts, (size_t)nbf * (NL - 1) + 1));
  TRY(fsb_dmalloc(ctx, &mesh->bf_opp, (size_t)nbf + 1));
  TRY(fsb_dmalloc(ctx, &mesh->bf_cell, (size_t)nbf + 1));
  TRY(fsb_dmalloc(ctx, &mesh->bf_id, (size_t)nbf + 1));
  if (nbf > 0) {
    TRYCUDA(cudaMemsetAsync(deg, 0, sizeof(int32_t) * (nv + 1), ctx->stream));
    k_f_fill<NL><<<fsb_grid(nc, 256, cap), 256, 0, ctx->stream>>>(mesh->cells, nc, mask, eptr, deg, rec);
    ctx->launches++; TRYCUDA(cudaGetLastError());
    k_f_sort<NL><<<fsb_grid(nv, 128, cap), 128, 0, ctx->stream>>>(nv, eptr, ubase, rec, mesh->bf_verts, mesh->bf_opp, mesh->bf_cell, mesh->bf_id);
    ctx->launches++; TRYCUDA(cudaGetLastError());
  }
  TRYCUDA(cudaStreamSynchronize(ctx->stream));
  cleanup();
  mesh->nbf = nbf;
  mesh->nfacets = nfacets;
  mesh->bf_id_ranked = false;
  return FSB_OK;
}

// dolfin facet ids of the exterior facets: bf_id holds ubase[a]; add each facet's rank inside the star of its smallest vertex
template <int NL>
int facet_ids_impl(fsb_mesh* mesh) {
  fsb_ctx* ctx = mesh->ctx;
  if (mesh->bf_id_ranked || mesh->nbf <= 0) { mesh->bf_id_ranked = true; return FSB_OK; }
  const int64_t nv = mesh->nverts, nc = mesh->ncells;
  const int cap = ctx->sm_count * 16;
  int32_t *deg = nullptr, *v2c = nullptr;
  int64_t* vptr = nullptr;
  int* d_over = nullptr;
  int rc = FSB_OK;
  auto cleanup = [&]() {
    if (v2c != mesh->v2c) fsb_dfree(ctx, v2c);
    if (vptr != mesh->v2c_ptr) fsb_dfree(ctx, vptr);
    v2c = nullptr; vptr = nullptr;
    fsb_dfree(ctx, deg); fsb_dfree(ctx, d_over);
  };
  TRY(fsb_dmalloc(ctx, &d_over, 1));
  TRYCUDA(cudaMemsetAsync(d_over, 0, sizeof(int), ctx->stream));
  if (mesh->degree == 1 && mesh->v2c) { vptr = mesh->v2c_ptr; v2c = mesh->v2c; }
  else {                                    // a degree-2 mesh keeps no vertex adjacency: rebuild it for this one pass
    TRY(fsb_dmalloc(ctx, &deg, (size_t)nv + 1));
    TRY(fsb_dmalloc(ctx, &vptr, (size_t)nv + 1));
    TRY(fsb_dmalloc(ctx, &v2c, (size_t)nc * NL));
    TRYCUDA(cudaMemsetAsync(deg, 0, sizeof(int32_t) * (nv + 1), ctx->stream));
    k_f_v2c_count<<<fsb_grid(nc * NL, 256, cap), 256, 0, ctx->stream>>>(mesh->cells, nc * NL, deg);
    ctx->launches++; TRYCUDA(cudaGetLastError());
    TRY(fsb_exclusive_scan(ctx, deg, vptr, nv));
    TRYCUDA(cudaMemsetAsync(deg, 0, sizeof(int32_t) * (nv + 1), ctx->stream));
    k_f_v2c_fill<NL><<<fsb_grid(nc, 256, cap), 256, 0, ctx->stream>>>(mesh->cells, nc, vptr, deg, v2c);
    ctx->launches++; TRYCUDA(cudaGetLastError());
  }
  k_f_rank<NL><<<fsb_grid(mesh->nbf, 128, (int64_t)ctx->sm_count * 4), 128, 0, ctx->stream>>>(mesh->cells, mesh->nbf, vptr, v2c, mesh->bf_verts, mesh->bf_id, d_over);
  ctx->launches++; TRYCUDA(cudaGetLastError());
  int over = 0;
  TRYCUDA(cudaMemcpyAsync(&over, d_over, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  TRYCUDA(cudaStreamSynchronize(ctx->stream));
  cleanup();
#undef TRY
#undef TRYCUDA
  if (over) FSB_FAIL(ctx, FSB_ERR_STATE, "facet numbering: a vertex star holds more than 384 distinct facets");
  mesh->bf_id_ranked = true;
  return FSB_OK;
}

}  // namespace

// ---- boundary geometry for SubDomain.mark: the distinct boundary vertices (ascending), each facet's vertices as indices into that
// list, the boundary vertices' coordinates and the facet midpoints ((x_a + x_b [+ x_c]) / tdim, summed in that order)
namespace {

__global__ void k_bg_flag(const int32_t* __restrict__ fverts, int64_t n, int32_t* __restrict__ flag) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) flag[fverts[i]] = 1;
}
__global__ void k_bg_compact(const int32_t* __restrict__ flag, const int64_t* __restrict__ rank, int64_t nverts, const double* __restrict__ xyz, int gdim,
                             int32_t* __restrict__ bverts, double* __restrict__ bxyz) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nverts; v += (int64_t)gridDim.x * blockDim.x)
    if (flag[v]) {
      const int64_t k = rank[v];
      bverts[k] = (int32_t)v;
      for (int i = 0; i < gdim; ++i) bxyz[k * gdim + i] = xyz[v * gdim + i];
    }
}
__global__ void k_bg_facets(const int32_t* __restrict__ fverts, int64_t nbf, int d, const int64_t* __restrict__ rank, const double* __restrict__ xyz, int gdim,
                            int32_t* __restrict__ finv, double* __restrict__ mid) {
  for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nbf; f += (int64_t)gridDim.x * blockDim.x) {
    for (int j = 0; j < d; ++j) finv[f * d + j] = (int32_t)rank[fverts[f * d + j]];
    for (int i = 0; i < gdim; ++i) {
      double s = xyz[(int64_t)fverts[f * d] * gdim + i];
      for (int j = 1; j < d; ++j) s = __dadd_rn(s, xyz[(int64_t)fverts[f * d + j] * gdim + i]);
      mid[f * gdim + i] = __ddiv_rn(s, (double)d);
    }
  }
}

}  // namespace

extern "C" int fsb_mesh_boundary_geometry(fsb_mesh* mesh, int64_t* nbv_out, int32_t* bverts, int32_t* finv, double* bxyz, double* mid) {
  if (!mesh) return FSB_ERR_ARG;
  fsb_ctx* ctx = mesh->ctx;
  int rc = fsb_mesh_exterior_facets(mesh, nullptr, nullptr);
  if (rc) return rc;
  const int64_t nbf = mesh->nbf, nv = mesh->nverts;
  const int d = mesh->tdim, g = mesh->gdim;
  const int cap = ctx->sm_count * 16;
  if (mesh->nbv < 0) {
    int32_t* flag = nullptr;
    int64_t* rank = nullptr;
    if ((rc = fsb_dmalloc(ctx, &flag, (size_t)nv + 1)) || (rc = fsb_dmalloc(ctx, &rank, (size_t)nv + 1))) { fsb_dfree(ctx, flag); return rc; }
    FSB_CHECK_CUDA(ctx, cudaMemsetAsync(flag, 0, sizeof(int32_t) * (nv + 1), ctx->stream));
    if (nbf > 0) {
      k_bg_flag<<<fsb_grid(nbf * d, 256, cap), 256, 0, ctx->stream>>>(mesh->bf_verts, nbf * d, flag);
      ctx->launches++;
    }
    rc = fsb_exclusive_scan(ctx, flag, rank, nv);
    int64_t nbv = 0;
    if (!rc) {
      FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(&nbv, rank + nv, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
      FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      if (!(rc = fsb_dmalloc(ctx, &mesh->bg_verts, (size_t)nbv + 1)) && !(rc = fsb_dmalloc(ctx, &mesh->bg_finv, (size_t)nbf * d + 1)) &&
          !(rc = fsb_dmalloc(ctx, &mesh->bg_xyz, (size_t)nbv * g + 1)) && !(rc = fsb_dmalloc(ctx, &mesh->bg_mid, (size_t)nbf * g + 1)) && nbf > 0) {
        k_bg_compact<<<fsb_grid(nv, 256, cap), 256, 0, ctx->stream>>>(flag, rank, nv, mesh->xyz, g, mesh->bg_verts, mesh->bg_xyz);
        ctx->launches++;
        k_bg_facets<<<fsb_grid(nbf, 256, cap), 256, 0, ctx->stream>>>(mesh->bf_verts, nbf, d, rank, mesh->xyz, g, mesh->bg_finv, mesh->bg_mid);
        ctx->launches++;
      }
    }
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    fsb_dfree(ctx, flag); fsb_dfree(ctx, rank);
    if (rc) return rc;
    if (e != cudaSuccess) FSB_FAIL(ctx, FSB_ERR_CUDA, cudaGetErrorString(e));
    mesh->nbv = nbv;
  }
  if (nbv_out) *nbv_out = mesh->nbv;
  const int64_t nbv = mesh->nbv;
  if (bverts && nbv) FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(bverts, mesh->bg_verts, sizeof(int32_t) * nbv, cudaMemcpyDeviceToHost, ctx->stream));
  if (finv && nbf) FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(finv, mesh->bg_finv, sizeof(int32_t) * nbf * d, cudaMemcpyDeviceToHost, ctx->stream));
  if (bxyz && nbv) FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(bxyz, mesh->bg_xyz, sizeof(double) * nbv * g, cudaMemcpyDeviceToHost, ctx->stream));
  if (mid && nbf) FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(mid, mesh->bg_mid, sizeof(double) * nbf * g, cudaMemcpyDeviceToHost, ctx->stream));
  FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FSB_OK;
}

extern "C" int fsb_mesh_exterior_facets(fsb_mesh* mesh, int64_t* nbf, int64_t* nfacets) {
  if (!mesh) return FSB_ERR_ARG;
  fsb_ctx* ctx = mesh->ctx;
  if (mesh->nbf < 0) {
    int rc;
    if (mesh->tdim == 3) rc = exterior_facets_impl<4>(mesh);
    else if (mesh->tdim == 2) rc = exterior_facets_impl<3>(mesh);
    else FSB_FAIL(ctx, FSB_ERR_ARG, "exterior facets: tdim must be 2 or 3");
    if (rc) return rc;
  }
  if (nbf) *nbf = mesh->nbf;
  if (nfacets) *nfacets = mesh->nfacets;
  return FSB_OK;
}

extern "C" int fsb_mesh_exterior_facets_get(fsb_mesh* mesh, int32_t* fverts, int32_t* opp, int32_t* cell, int64_t* facet_id) {
  if (!mesh) return FSB_ERR_ARG;
  fsb_ctx* ctx = mesh->ctx;
  if (mesh->nbf < 0) FSB_FAIL(ctx, FSB_ERR_STATE, "call fsb_mesh_exterior_facets first");
  const int64_t n = mesh->nbf;
  if (facet_id && !mesh->bf_id_ranked) {
    const int rc = mesh->tdim == 3 ? facet_ids_impl<4>(mesh) : facet_ids_impl<3>(mesh);
    if (rc) return rc;
  }
  if (n > 0) {
    if (fverts) FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(fverts, mesh->bf_verts, sizeof(int32_t) * n * mesh->tdim, cudaMemcpyDeviceToHost, ctx->stream));
    if (opp) FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(opp, mesh->bf_opp, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (cell) FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(cell, mesh->bf_cell, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (facet_id) FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(facet_id, mesh->bf_id, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
  }
  FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FSB_OK;
}
