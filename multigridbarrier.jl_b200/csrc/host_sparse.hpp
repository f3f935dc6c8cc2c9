// host_sparse.hpp -- setup-time sparse index algebra on the host (runs once per mgbx_create).
//
// Nothing here is on the hot path: it builds the fixed index structures the device kernels then
// reuse in every Newton iteration -- the role of the reference's BlockAssemblyPlan
// (src/BlockMatrices.jl:281-491: col_indices / scatter_idx / output pattern, cached per R) and of the
// CUDA extension's CPU-side plan construction (ext/MultiGridBarrierCUDAExt/block_ops.jl:251-411).
#pragma once
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <stdexcept>
#include <vector>

#include "../../include/mgbx.h"

namespace mgbx {

struct HostCsr {
  int64_t rows = 0, cols = 0;
  std::vector<int64_t> ptr;   // rows + 1
  std::vector<int32_t> idx;
  std::vector<double> val;    // may be empty (pattern only)
  int64_t nnz() const { return ptr.empty() ? 0 : ptr.back(); }
};

inline HostCsr csr_from_abi(const mgbx_csr &A, bool with_values = true) {
  HostCsr H;
  H.rows = A.rows;
  H.cols = A.cols;
  if (A.rows < 0 || A.cols < 0 || A.cols > INT32_MAX || !A.rowptr) throw std::invalid_argument("bad CSR matrix");
  H.ptr.assign(A.rowptr, A.rowptr + A.rows + 1);
  const int64_t nz = H.ptr.back();
  if (H.ptr[0] != 0 || nz < 0) throw std::invalid_argument("CSR rowptr must start at 0");
  H.idx.resize(nz);
  for (int64_t i = 0; i < A.rows; ++i) {
    if (H.ptr[i + 1] < H.ptr[i]) throw std::invalid_argument("CSR rowptr not monotone");
    for (int64_t k = H.ptr[i]; k < H.ptr[i + 1]; ++k) {
      const int64_t c = A.colind[k];
      if (c < 0 || c >= A.cols) throw std::invalid_argument("CSR column index out of range");
      if (k > H.ptr[i] && c <= A.colind[k - 1]) throw std::invalid_argument("CSR columns must be sorted and unique per row");
      H.idx[k] = (int32_t)c;
    }
  }
  if (with_values) H.val.assign(A.val, A.val + nz);
  return H;
}

inline HostCsr transpose(const HostCsr &A) {
  HostCsr T;
  T.rows = A.cols;
  T.cols = A.rows;
  T.ptr.assign(A.cols + 1, 0);
  const int64_t nz = A.nnz();
  for (int64_t k = 0; k < nz; ++k) T.ptr[A.idx[k] + 1]++;
  for (int64_t c = 0; c < A.cols; ++c) T.ptr[c + 1] += T.ptr[c];
  T.idx.resize(nz);
  const bool hv = !A.val.empty();
  if (hv) T.val.resize(nz);
  std::vector<int64_t> pos(T.ptr.begin(), T.ptr.end() - 1);
  for (int64_t i = 0; i < A.rows; ++i)
    for (int64_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k) {
      const int64_t q = pos[A.idx[k]]++;
      T.idx[q] = (int32_t)i;
      if (hv) T.val[q] = A.val[k];
    }
  return T;
}

// rows [r0, r1) x cols [c0, c1), re-based to 0
inline HostCsr submatrix(const HostCsr &A, int64_t r0, int64_t r1, int64_t c0, int64_t c1) {
  HostCsr S;
  S.rows = r1 - r0;
  S.cols = c1 - c0;
  S.ptr.assign(S.rows + 1, 0);
  const bool hv = !A.val.empty();
  for (int64_t i = r0; i < r1; ++i) {
    for (int64_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
      if (A.idx[k] >= c0 && A.idx[k] < c1) {
        S.idx.push_back((int32_t)(A.idx[k] - c0));
        if (hv) S.val.push_back(A.val[k]);
      }
    S.ptr[i - r0 + 1] = (int64_t)S.idx.size();
  }
  return S;
}

// block-diagonal join of matrices (used to restrict the level transfers to the kept variables)
inline HostCsr block_diag(const std::vector<HostCsr> &B) {
  HostCsr D;
  for (auto &b : B) {
    D.rows += b.rows;
    D.cols += b.cols;
  }
  D.ptr.assign(1, 0);
  D.ptr.reserve(D.rows + 1);
  int64_t coff = 0;
  for (auto &b : B) {
    for (int64_t i = 0; i < b.rows; ++i) {
      for (int64_t k = b.ptr[i]; k < b.ptr[i + 1]; ++k) {
        D.idx.push_back((int32_t)(b.idx[k] + coff));
        D.val.push_back(b.val.empty() ? 1.0 : b.val[k]);
      }
      D.ptr.push_back((int64_t)D.idx.size());
    }
    coff += b.cols;
  }
  return D;
}

inline bool is_identity(const HostCsr &A) {
  if (A.rows != A.cols || A.nnz() != A.rows) return false;
  for (int64_t i = 0; i < A.rows; ++i) {
    if (A.ptr[i + 1] - A.ptr[i] != 1 || A.idx[A.ptr[i]] != i) return false;
    if (!A.val.empty() && A.val[A.ptr[i]] != 1.0) return false;
  }
  return true;
}

// pattern of A*B (sorted columns per row)
inline HostCsr spgemm_symbolic(const HostCsr &A, const HostCsr &B) {
  if (A.cols != B.rows) throw std::invalid_argument("spgemm: inner dimensions differ");
  HostCsr C;
  C.rows = A.rows;
  C.cols = B.cols;
  C.ptr.assign(A.rows + 1, 0);
  std::vector<int64_t> mark(B.cols, -1);
  std::vector<int32_t> rowbuf;
  for (int64_t i = 0; i < A.rows; ++i) {
    rowbuf.clear();
    for (int64_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k) {
      const int64_t a = A.idx[k];
      for (int64_t q = B.ptr[a]; q < B.ptr[a + 1]; ++q) {
        const int32_t c = B.idx[q];
        if (mark[c] != i) {
          mark[c] = i;
          rowbuf.push_back(c);
        }
      }
    }
    std::sort(rowbuf.begin(), rowbuf.end());
    C.idx.insert(C.idx.end(), rowbuf.begin(), rowbuf.end());
    C.ptr[i + 1] = (int64_t)C.idx.size();
  }
  return C;
}

// A*B with values (sorted columns per row); used once at setup to compose R_fine[l] = R_fine[l+1]*T[l]
inline HostCsr spgemm_numeric_host(const HostCsr &A, const HostCsr &B) {
  HostCsr C = spgemm_symbolic(A, B);
  C.val.assign(C.idx.size(), 0.0);
  std::vector<int64_t> posmap(B.cols, -1);
  for (int64_t i = 0; i < A.rows; ++i) {
    for (int64_t k = C.ptr[i]; k < C.ptr[i + 1]; ++k) posmap[C.idx[k]] = k;
    for (int64_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k) {
      const int64_t a = A.idx[k];
      const double av = A.val[k];
      for (int64_t q = B.ptr[a]; q < B.ptr[a + 1]; ++q) C.val[posmap[B.idx[q]]] += av * B.val[q];
    }
  }
  return C;
}

inline int64_t find_in_row(const HostCsr &A, int64_t row, int32_t col) {
  const int32_t *b = A.idx.data() + A.ptr[row], *e = A.idx.data() + A.ptr[row + 1];
  const int32_t *it = std::lower_bound(b, e, col);
  if (it == e || *it != col) return -1;
  return (int64_t)(it - A.idx.data());
}

// Element-to-unknown incidence: row e lists the sorted unique columns of R touched by the p rows
// of element e in the row blocks of the listed variables, shifted by colmap (column -> system
// unknown, -1 = not in the system).  This is the union over k of the reference plan's col_indices
// for state block k (src/BlockMatrices.jl:344-379).
inline HostCsr element_incidence(const HostCsr &R, int64_t N, int p, const std::vector<int> &vars, int64_t n,
                                 const std::vector<int64_t> &colmap, int64_t msys) {
  HostCsr E;
  E.rows = N;
  E.cols = msys;
  E.ptr.assign(N + 1, 0);
  std::vector<int32_t> buf;
  for (int64_t e = 0; e < N; ++e) {
    buf.clear();
    for (int v : vars) {
      const int64_t r0 = (int64_t)v * n + e * p;
      for (int64_t k = R.ptr[r0]; k < R.ptr[r0 + p]; ++k) {
        const int64_t c = colmap[R.idx[k]];
        if (c >= 0) buf.push_back((int32_t)c);
      }
    }
    std::sort(buf.begin(), buf.end());
    buf.erase(std::unique(buf.begin(), buf.end()), buf.end());
    E.idx.insert(E.idx.end(), buf.begin(), buf.end());
    E.ptr[e + 1] = (int64_t)E.idx.size();
  }
  return E;
}

// The reference plan's output pattern: union over elements of cols(e) x cols(e)
// (src/BlockMatrices.jl:381-446); symmetric, so CSC == CSR.
inline HostCsr plan_pattern(const HostCsr &Einc) { return spgemm_symbolic(transpose(Einc), Einc); }

}  // namespace mgbx
