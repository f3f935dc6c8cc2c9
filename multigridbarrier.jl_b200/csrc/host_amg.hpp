// host_amg.hpp -- classical Ruge-Stueben coarsening on the host (C++), behind mgbx_rs_*.
//
// The reference obtains its prolongators from the un-vendored AlgebraicMultigrid.jl (`ruge_stuben(K; max_coarse=2).levels[i].P`,
// src/amg_prolongators.jl:16-18; version unpinned).  The Python host mirror restates the published
// algorithm (hierarchy.py: classical strength theta = 0.25, first-pass RS C/F splitting with the bucket-sorted lambda measure,
// direct interpolation, Galerkin P'AP, max_levels = 10); this is the same algorithm in C++, loop for loop, so that the two can be
// compared bit for bit (tests/test_abi_cpu.py) and the hierarchy construction of SURVEY.md section 8(f) row 3 no longer needs the
// Python/numba path.  The sparse products reproduce SciPy's row-wise accumulation order (first-touch linked list, explicit zeros
// dropped, rows sorted afterwards), which is what makes the coarse operators -- and hence every later level -- bitwise equal.
// Host-only: no CUDA in this file.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <vector>

namespace mgbx {

struct AmgCsr {
  int64_t rows = 0, cols = 0;
  std::vector<int64_t> ptr, idx;
  std::vector<double> val;
};

// C = A * B exactly as scipy.sparse._sparsetools.csr_matmat computes it (then NOT sorted): per row the columns come out in reverse
// first-touch order and entries whose sum is exactly zero are dropped
inline AmgCsr amg_matmat(const AmgCsr &A, const AmgCsr &B) {
  AmgCsr C;
  C.rows = A.rows;
  C.cols = B.cols;
  C.ptr.assign(A.rows + 1, 0);
  std::vector<int64_t> next(B.cols, -1);
  std::vector<double> sums(B.cols, 0.0);
  for (int64_t i = 0; i < A.rows; ++i) {
    int64_t head = -2, length = 0;
    for (int64_t jj = A.ptr[i]; jj < A.ptr[i + 1]; ++jj) {
      const int64_t j = A.idx[jj];
      const double v = A.val[jj];
      for (int64_t kk = B.ptr[j]; kk < B.ptr[j + 1]; ++kk) {
        const int64_t k = B.idx[kk];
        sums[k] += v * B.val[kk];
        if (next[k] == -1) {
          next[k] = head;
          head = k;
          ++length;
        }
      }
    }
    for (int64_t jj = 0; jj < length; ++jj) {
      if (sums[head] != 0.0) {
        C.idx.push_back(head);
        C.val.push_back(sums[head]);
      }
      const int64_t tmp = head;
      head = next[head];
      next[tmp] = -1;
      sums[tmp] = 0.0;
    }
    C.ptr[i + 1] = (int64_t)C.idx.size();
  }
  return C;
}

inline void amg_sort_rows(AmgCsr &A) {
  std::vector<std::pair<int64_t, double>> row;
  for (int64_t i = 0; i < A.rows; ++i) {
    const int64_t b = A.ptr[i], e = A.ptr[i + 1];
    row.resize((size_t)(e - b));
    for (int64_t k = b; k < e; ++k) row[(size_t)(k - b)] = {A.idx[k], A.val[k]};
    std::sort(row.begin(), row.end(), [](const std::pair<int64_t, double> &x, const std::pair<int64_t, double> &y) { return x.first < y.first; });
    for (int64_t k = b; k < e; ++k) {
      A.idx[k] = row[(size_t)(k - b)].first;
      A.val[k] = row[(size_t)(k - b)].second;
    }
  }
}

// transpose with sorted rows (SciPy: csr -> csc reinterpretation, then tocsr)
inline AmgCsr amg_transpose(const AmgCsr &A) {
  AmgCsr T;
  T.rows = A.cols;
  T.cols = A.rows;
  T.ptr.assign(T.rows + 1, 0);
  for (int64_t k = 0; k < (int64_t)A.idx.size(); ++k) T.ptr[A.idx[k] + 1]++;
  for (int64_t i = 0; i < T.rows; ++i) T.ptr[i + 1] += T.ptr[i];
  T.idx.resize(A.idx.size());
  T.val.resize(A.val.size());
  std::vector<int64_t> fill(T.ptr.begin(), T.ptr.end() - 1);
  for (int64_t i = 0; i < A.rows; ++i)
    for (int64_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k) {
      const int64_t p = fill[A.idx[k]]++;
      T.idx[p] = i;
      T.val[p] = A.val[k];
    }
  return T;
}

// classical strength of connection: keep the diagonal and every entry with |a_ij| >= theta max_{k != i} |a_ik|
inline void amg_strength(const AmgCsr &A, double theta, std::vector<int64_t> &Sp, std::vector<int64_t> &Sj) {
  const int64_t n = A.rows;
  Sp.assign(n + 1, 0);
  std::vector<char> keep(A.idx.size(), 0);
  for (int64_t i = 0; i < n; ++i) {
    double m = 0.0;
    for (int64_t jj = A.ptr[i]; jj < A.ptr[i + 1]; ++jj)
      if (A.idx[jj] != i && std::fabs(A.val[jj]) > m) m = std::fabs(A.val[jj]);
    const double thr = theta * m;
    int64_t c = 0;
    for (int64_t jj = A.ptr[i]; jj < A.ptr[i + 1]; ++jj)
      if (A.idx[jj] == i || std::fabs(A.val[jj]) >= thr) {
        keep[jj] = 1;
        ++c;
      }
    Sp[i + 1] = Sp[i] + c;
  }
  Sj.clear();
  Sj.reserve((size_t)Sp[n]);
  for (int64_t jj = 0; jj < (int64_t)A.idx.size(); ++jj)
    if (keep[jj]) Sj.push_back(A.idx[jj]);
}

// first-pass Ruge-Stueben C/F splitting (1 = C, 0 = F) with the bucket-sorted measure lambda_i = |S^T_i|
inline std::vector<int64_t> amg_cf_splitting(int64_t n, const std::vector<int64_t> &Sp, const std::vector<int64_t> &Sj, const std::vector<int64_t> &Tp,
                                             const std::vector<int64_t> &Tj) {
  const int64_t U = 2, Cpt = 1, Fpt = 0;
  std::vector<int64_t> lam(n), interval_ptr(n + 2, 0), interval_count(n + 2, 0), index_to_node(n, 0), node_to_index(n, 0);
  for (int64_t i = 0; i < n; ++i) lam[i] = Tp[i + 1] - Tp[i];
  for (int64_t i = 0; i < n; ++i) interval_count[lam[i]]++;
  int64_t cum = 0;
  for (int64_t i = 0; i < n + 1; ++i) {
    interval_ptr[i] = cum;
    cum += interval_count[i];
    interval_count[i] = 0;
  }
  for (int64_t i = 0; i < n; ++i) {
    const int64_t li = lam[i];
    const int64_t idx = interval_ptr[li] + interval_count[li];
    index_to_node[idx] = i;
    node_to_index[i] = idx;
    interval_count[li]++;
  }
  std::vector<int64_t> splitting(n, U);
  for (int64_t i = 0; i < n; ++i)
    if (lam[i] == 0 || (lam[i] == 1 && Tj[Tp[i]] == i)) splitting[i] = Fpt;
  for (int64_t top = n - 1; top >= 0; --top) {
    const int64_t i = index_to_node[top];
    const int64_t li = lam[i];
    interval_count[li]--;
    if (splitting[i] == Fpt) continue;
    splitting[i] = Cpt;
    for (int64_t jj = Tp[i]; jj < Tp[i + 1]; ++jj) {
      const int64_t j = Tj[jj];
      if (splitting[j] == U) {
        splitting[j] = Fpt;
        for (int64_t kk = Sp[j]; kk < Sp[j + 1]; ++kk) {
          const int64_t k = Sj[kk];
          if (splitting[k] == U) {
            if (lam[k] >= n - 1) continue;
            const int64_t lk = lam[k];
            const int64_t old = node_to_index[k];
            const int64_t nw = interval_ptr[lk] + interval_count[lk] - 1;
            node_to_index[index_to_node[old]] = nw;
            node_to_index[index_to_node[nw]] = old;
            std::swap(index_to_node[old], index_to_node[nw]);
            interval_count[lk]--;
            interval_count[lk + 1]++;
            interval_ptr[lk + 1] = nw;
            lam[k]++;
          }
        }
      }
    }
    for (int64_t jj = Sp[i]; jj < Sp[i + 1]; ++jj) {
      const int64_t j = Sj[jj];
      if (splitting[j] == U) {
        if (lam[j] == 0) continue;
        const int64_t lj = lam[j];
        const int64_t old = node_to_index[j];
        const int64_t nw = interval_ptr[lj];
        node_to_index[index_to_node[old]] = nw;
        node_to_index[index_to_node[nw]] = old;
        std::swap(index_to_node[old], index_to_node[nw]);
        interval_count[lj]--;
        interval_count[lj - 1]++;
        interval_ptr[lj]++;
        interval_ptr[lj - 1] = interval_ptr[lj] - interval_count[lj - 1];
        lam[j]--;
      }
    }
  }
  return splitting;
}

// direct interpolation: C points are injected, an F point interpolates from its strong C neighbours with the row's negative /
// positive off-diagonal mass redistributed over them
inline AmgCsr amg_direct_interpolation(const AmgCsr &A, const std::vector<int64_t> &Sp, const std::vector<int64_t> &Sj, const std::vector<int64_t> &splitting) {
  const int64_t n = A.rows;
  std::vector<int64_t> cmap(n, 0);
  int64_t nc = 0;
  for (int64_t i = 0; i < n; ++i)
    if (splitting[i] == 1) cmap[i] = nc++;
  AmgCsr P;
  P.rows = n;
  P.cols = nc;
  P.ptr.assign(n + 1, 0);
  for (int64_t i = 0; i < n; ++i) {
    if (splitting[i] == 1) P.ptr[i + 1] = P.ptr[i] + 1;
    else {
      int64_t c = 0;
      for (int64_t jj = Sp[i]; jj < Sp[i + 1]; ++jj)
        if (splitting[Sj[jj]] == 1 && Sj[jj] != i) ++c;
      P.ptr[i + 1] = P.ptr[i] + c;
    }
  }
  P.idx.assign((size_t)P.ptr[n], 0);
  P.val.assign((size_t)P.ptr[n], 0.0);
  std::vector<char> strong(n, 0);
  for (int64_t i = 0; i < n; ++i) {
    if (splitting[i] == 1) {
      P.idx[P.ptr[i]] = cmap[i];
      P.val[P.ptr[i]] = 1.0;
      continue;
    }
    for (int64_t jj = Sp[i]; jj < Sp[i + 1]; ++jj) strong[Sj[jj]] = 1;
    double sum_strong_pos = 0.0, sum_strong_neg = 0.0, sum_all_pos = 0.0, sum_all_neg = 0.0, diag = 0.0;
    for (int64_t jj = A.ptr[i]; jj < A.ptr[i + 1]; ++jj) {
      const int64_t j = A.idx[jj];
      const double v = A.val[jj];
      if (j == i) diag += v;
      else {
        if (v < 0) sum_all_neg += v;
        else sum_all_pos += v;
        if (strong[j] && splitting[j] == 1) {
          if (v < 0) sum_strong_neg += v;
          else sum_strong_pos += v;
        }
      }
    }
    const double alpha = sum_strong_neg != 0.0 ? sum_all_neg / sum_strong_neg : 0.0;
    double beta = sum_strong_pos != 0.0 ? sum_all_pos / sum_strong_pos : 0.0;
    if (sum_strong_pos == 0.0) {
      diag += sum_all_pos;
      beta = 0.0;
    }
    const double neg_coeff = diag != 0.0 ? -alpha / diag : 0.0;
    const double pos_coeff = diag != 0.0 ? -beta / diag : 0.0;
    int64_t nnz = P.ptr[i];
    for (int64_t jj = A.ptr[i]; jj < A.ptr[i + 1]; ++jj) {
      const int64_t j = A.idx[jj];
      if (j != i && strong[j] && splitting[j] == 1) {
        const double v = A.val[jj];
        P.idx[nnz] = cmap[j];
        P.val[nnz] = (v < 0 ? neg_coeff : pos_coeff) * v;
        ++nnz;
      }
    }
    for (int64_t jj = Sp[i]; jj < Sp[i + 1]; ++jj) strong[Sj[jj]] = 0;
  }
  return P;
}

// prolongations finest -> coarsest (hierarchy.py ruge_stuben; call site src/amg_prolongators.jl:16-18)
inline std::vector<AmgCsr> amg_ruge_stuben(AmgCsr A, int max_coarse, int max_levels, double theta) {
  amg_sort_rows(A);
  std::vector<AmgCsr> Ps;
  while ((int)Ps.size() + 1 < max_levels && A.rows > max_coarse) {
    const int64_t n = A.rows;
    std::vector<int64_t> Sp, Sj;
    amg_strength(A, theta, Sp, Sj);
    // T = S^T (pattern only), rows sorted
    AmgCsr S;
    S.rows = S.cols = n;
    S.ptr = Sp;
    S.idx = Sj;
    S.val.assign(Sj.size(), 1.0);
    const AmgCsr T = amg_transpose(S);
    const std::vector<int64_t> splitting = amg_cf_splitting(n, Sp, Sj, T.ptr, T.idx);
    AmgCsr P = amg_direct_interpolation(A, Sp, Sj, splitting);
    if (P.cols == 0 || P.cols == n) break;
    // A <- P' A P in SciPy's evaluation order: `P.T @ A @ P` with P.T a CSC matrix is two CSC products, each computed as the
    // CSR product of the transposed operands in swapped order -- (A' P) first, then P' (A' P) -- and the result is converted back
    // to CSR (a transpose with sorted rows).  Reproducing that order is what keeps the coarse operators bitwise equal.
    const AmgCsr Pt = amg_transpose(P);
    const AmgCsr At = amg_transpose(A);
    const AmgCsr AtP = amg_matmat(At, P);
    A = amg_transpose(amg_matmat(Pt, AtP));
    Ps.push_back(std::move(P));
  }
  return Ps;
}

}  // namespace mgbx
