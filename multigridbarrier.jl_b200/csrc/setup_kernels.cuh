// setup_kernels.cuh -- plan construction on the device (runs once per linear-system family, at the first
// Newton iteration that needs it).  Not on the per-iteration hot path, but on the end-to-end path of every
// mgb_solve call, so it is done on the GPU instead of in host loops:
//   device_symbolic      sparsity pattern of a sparse product (candidate keys -> radix sort -> unique), the
//                        symbolic half of the Galerkin products A T and T'(A T)
//   k_top_plan           term list of the element-block -> CSR gather (the scatter_idx of
//                        src/BlockMatrices.jl:448-491, inverted into a fixed-order gather)
// CUB (radix sort / scan / unique) is used for the sort; everything else is hand-written.
#pragma once
#include <cub/cub.cuh>

#include "solver_kernels.cuh"

namespace mgbx {

// number of candidate (row, col) pairs of row i of A*B
__global__ void k_sym_count(DevCsr A, DevCsr B, int64_t *cand) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > A.rows) return;
  if (i == A.rows) {
    cand[i] = 0;
    return;
  }
  int64_t c = 0;
  for (int64_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k) {
    const int64_t r = A.idx[k];
    c += B.ptr[r + 1] - B.ptr[r];
  }
  cand[i] = c;
}

__global__ void k_sym_fill(DevCsr A, DevCsr B, const int64_t *__restrict__ offs, unsigned long long *keys) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.rows) return;
  int64_t pos = offs[i];
  for (int64_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k) {
    const int64_t r = A.idx[k];
    for (int64_t q = B.ptr[r]; q < B.ptr[r + 1]; ++q) keys[pos++] = ((unsigned long long)i << 32) | (unsigned int)B.idx[q];
  }
}

// ptr[i] = first position with key >= (i << 32); idx = low words
__global__ void k_sym_finish(const unsigned long long *__restrict__ uniq, int64_t nnz, int64_t rows, int64_t *ptr, int32_t *idx) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nnz) idx[t] = (int32_t)(uniq[t] & 0xffffffffull);
  if (t <= rows) {
    const unsigned long long key = (unsigned long long)t << 32;
    int64_t lo = 0, hi = nnz;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (uniq[mid] < key) lo = mid + 1;
      else hi = mid;
    }
    ptr[t] = lo;
  }
}

// Element-block -> CSR gather plan.  One thread per non-zero (i, j) of the top pattern: the pair of state
// variables is fixed by the variable blocks i and j live in; the elements incident to i come from the
// transposed element incidence; inside an element the rows r (resp. c) whose R row carries column i (resp. j)
// are found by binary search.  Terms are emitted in the order (element, r, c): fixed, so the gather is
// deterministic.  FILL == 0 counts (cnt, slice widths), FILL == 1 writes src / w.
struct TopPlanParams {
  DevCsr R;                    // level->broken prolongation of the system's top level (rows nu*n)
  DevCsr EincT;                // system unknown -> incident elements (sorted)
  DevCsr pat;                  // top pattern
  const int32_t *rowof;        // row of every non-zero of pat
  int64_t n, N;
  int p, nkept, unit;
  int kept[MGBX_MAX_ND];       // state variable of kept slot q
  int64_t off[MGBX_MAX_ND + 1];    // system offsets of the kept slots
  int64_t rcol0[MGBX_MAX_ND];      // first R column of the kept slot's variable
  int pair_of[MGBX_MAX_ND * MGBX_MAX_ND];   // [qa * nkept + qb] -> pair index or -1
};

template <int FILL>
__global__ void __launch_bounds__(256) k_top_plan(TopPlanParams Q, SellPlan P, int32_t *width) {
  const int64_t nz = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int count = 0;
  if (nz < Q.pat.nnz) {
    const int64_t i = Q.rowof[nz], j = Q.pat.idx[nz];
    int qa = 0, qb = 0;
    while (qa + 1 < Q.nkept && i >= Q.off[qa + 1]) ++qa;
    while (qb + 1 < Q.nkept && j >= Q.off[qb + 1]) ++qb;
    const int pr = Q.pair_of[qa * Q.nkept + qb];
    if (pr >= 0) {
      const int va = Q.kept[qa], vb = Q.kept[qb];
      const int32_t ci = (int32_t)(Q.rcol0[qa] + (i - Q.off[qa])), cj = (int32_t)(Q.rcol0[qb] + (j - Q.off[qb]));
      const int p = Q.p;
      const int64_t off = FILL ? P.sptr[nz >> 5] + (nz & 31) : 0;
      for (int64_t t = Q.EincT.ptr[i]; t < Q.EincT.ptr[i + 1]; ++t) {
        const int64_t e = Q.EincT.idx[t];
        for (int r = 0; r < p; ++r) {
          const int64_t ka = csr_find(Q.R, (int64_t)va * Q.n + e * p + r, ci);
          if (ka < 0) continue;
          for (int c = 0; c < p; ++c) {
            const int64_t kb = csr_find(Q.R, (int64_t)vb * Q.n + e * p + c, cj);
            if (kb < 0) continue;
            if (FILL) {
              P.src[off + 32 * (int64_t)count] = (int32_t)((((int64_t)pr * Q.N + e) * p + r) * p + c);
              if (!Q.unit) P.w[off + 32 * (int64_t)count] = Q.R.val[ka] * Q.R.val[kb];
            }
            ++count;
          }
        }
      }
    }
    if (!FILL) P.cnt[nz] = count;
  }
  if (!FILL) {
    int wmax = count;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if ((threadIdx.x & 31) == 0 && nz < Q.pat.nnz) width[nz >> 5] = wmax;
  }
}

}  // namespace mgbx
