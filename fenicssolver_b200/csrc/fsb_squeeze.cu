// Operand squeeze for the Krylov solve ("drop_zeros" option).
//
// The assembled matrix keeps every structural entry of the P1/P2 pattern, as dolfin/PETSc do (the pattern is a
// parity object: row_ptr/col_idx are compared bit for bit).  On right-angled meshes many of those entries are
// exactly 0.0 after assembly (UnitCubeMesh: 8 of the 15 entries of an interior Laplace row are sums of exact
// zeros, SURVEY 8c KAT 3), and symmetric Dirichlet elimination zeroes more.  With drop_zeros the solver
// multiplies by a compacted copy that holds only the blocks with a non-zero entry: same products, fewer bytes.
// Two passes over the matrix per solve (count, fill), about two SpMVs' worth of traffic.
#include "fsb_internal.cuh"

template <int BS>
__device__ __forceinline__ bool block_nonzero(const double* __restrict__ v) {
  bool nz = false;
#pragma unroll
  for (int e = 0; e < BS * BS; ++e) nz |= (v[e] != 0.0);
  return nz;
}

// one thread per block row; FILL = false: row lengths, FILL = true: copy the kept blocks to their new place
template <int BS, bool FILL>
__global__ void __launch_bounds__(256)
k_squeeze(int64_t nbrows, const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx, const double* __restrict__ vals,
          int32_t* __restrict__ len, const int64_t* __restrict__ new_ptr, int32_t* __restrict__ new_col, double* __restrict__ new_vals) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < nbrows; r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k0 = row_ptr[r], k1 = row_ptr[r + 1];
    int64_t o = FILL ? new_ptr[r] : 0;
    int n = 0;
    for (int64_t k = k0; k < k1; ++k) {
      const double* v = vals + k * BS * BS;
      // the diagonal block always stays (the Jacobi preconditioner and the Dirichlet rows look it up)
      if (block_nonzero<BS>(v) || col_idx[k] == r) {
        if (FILL) {
          new_col[o] = col_idx[k];
#pragma unroll
          for (int e = 0; e < BS * BS; ++e) new_vals[o * BS * BS + e] = v[e];
          ++o;
        }
        ++n;
      }
    }
    if (!FILL) len[r] = n;
  }
}

// Build (or refresh) A->sq from the current values of A.  *out = the operand the SpMV launches should use.
int fsb_mat_squeeze(fsb_mat* A, fsb_mat** out) {
  fsb_ctx* ctx = A->ctx;
  *out = A;
  if (A->nbrows <= 0 || A->nnzb <= 0) return FSB_OK;
  if (!A->sq) {
    A->sq = new fsb_mat();
    A->sq->ctx = ctx;
    A->sq->bs = A->bs;
    A->sq->nbrows = A->nbrows;
  }
  fsb_mat* S = A->sq;
  S->own0 = A->own0; S->own1 = A->own1; S->max_row_len = A->max_row_len;
  const int64_t nr = A->nbrows;
  const unsigned grid = fsb_grid(nr, 256, (int64_t)ctx->sm_count * 16);
  int32_t* len = nullptr;
  int rc = fsb_dmalloc(ctx, &len, (size_t)nr + 1);
  if (rc) return rc;
  if (!S->row_ptr && (rc = fsb_dmalloc(ctx, &S->row_ptr, (size_t)nr + 1))) { fsb_dfree(ctx, len); return rc; }
#define SQ_LAUNCH(BS, FILL) k_squeeze<BS, FILL><<<grid, 256, 0, ctx->stream>>>(nr, A->row_ptr, A->col_idx, A->vals, len, S->row_ptr, S->col_idx, S->vals)
  if (A->bs == 1) SQ_LAUNCH(1, false); else if (A->bs == 2) SQ_LAUNCH(2, false); else SQ_LAUNCH(3, false);
  ctx->launches++;
  rc = fsb_exclusive_scan(ctx, len, S->row_ptr, nr);
  fsb_dfree(ctx, len);
  if (rc) return rc;
  int64_t nnzb = 0;
  FSB_CHECK_CUDA(ctx, cudaMemcpyAsync(&nnzb, S->row_ptr + nr, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  FSB_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  // drop_zeros == 2 (auto, the default): the compacted copy pays for its two extra passes and its memory only when a good part of
  // the stored blocks is exactly zero (right-angled box meshes: 53 % of a P1 Laplace matrix); an unstructured mesh has none and the
  // solve runs on A itself after this one counting pass
  if (ctx->drop_zeros == 2 && (double)nnzb > 0.8 * (double)A->nnzb) return FSB_OK;
  if (nnzb > S->sq_cap) {          // grow (first use, or more non-zeros than last time)
    fsb_dfree(ctx, S->col_idx); fsb_dfree(ctx, S->vals);
    S->col_idx = nullptr; S->vals = nullptr; S->sq_cap = 0;
    if ((rc = fsb_dmalloc(ctx, &S->col_idx, (size_t)nnzb)) || (rc = fsb_dmalloc(ctx, &S->vals, (size_t)nnzb * A->bs * A->bs))) return rc;
    S->sq_cap = nnzb;
  }
  S->nnzb = nnzb;
  if (A->bs == 1) SQ_LAUNCH(1, true); else if (A->bs == 2) SQ_LAUNCH(2, true); else SQ_LAUNCH(3, true);
#undef SQ_LAUNCH
  FSB_LAUNCH_CHECK(ctx);
  // the TMA tiles over-read up to a few entries past the last one: keep the tail defined
  FSB_CHECK_CUDA(ctx, cudaMemsetAsync(S->col_idx + nnzb, 0, 64, ctx->stream));
  FSB_CHECK_CUDA(ctx, cudaMemsetAsync(S->vals + nnzb * A->bs * A->bs, 0, 512, ctx->stream));
  S->tile_rows = 0;
  if ((rc = fsb_mat_setup_tiles(S))) return rc;
  *out = S;
  return FSB_OK;
}
