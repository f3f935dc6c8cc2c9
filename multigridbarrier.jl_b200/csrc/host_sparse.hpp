// host_sparse.hpp -- setup-time sparse index algebra on the host (runs once per mgbx_create).
//
// Nothing here is on the hot path: it builds the fixed index structures the device kernels then
// reuse in every Newton iteration -- the role of the reference's BlockAssemblyPlan
// (src/BlockMatrices.jl:281-491: col_indices / scatter_idx / output pattern, cached per R) and of the
// CUDA extension's CPU-side plan construction (ext/MultiGridBarrierCUDAExt/block_ops.jl:251-411).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <string>
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

// ------------------------------------------------------------------------------------------------
// Level transfers from the composed prolongations.  The reference composes R_fine[l] = R_fine[l+1] * T[l] at
// construction time and discards the level-to-level factors (src/multigrid.jl:166-170, _compose_R :192-205), so a
// backend that is handed an unmodified `AMG` sees only R_fine[1..L].  T[l] is recovered here, in this order:
//   1. selector rows (exact, O(rows)): if every column j of R_fine[l+1] owns a row i_j whose only entry is
//      (i_j, j) = a_j, then T[l][j, :] = R_fine[l][i_j, :] / a_j.  True for every FEM hierarchy of the reference:
//      0/1 gluing matrices, the identity of `:full`, the ones column of `:uniform`, the per-element linears of
//      `:broken_P1`, the corner -> broken bridge, and classical Ruge-Stueben interpolation (C points are injected);
//   2. column matching (exact): every column of R_fine[l] equals a column of R_fine[l+1] (spectral hierarchies:
//      level l = the leading Chebyshev modes, src/spectral1d.jl:63-109, kron in 2-D) -> T[l] is a 0/1 selection;
//   3. dense normal equations (R'R) T = R' R_l for small levels (<= 2048 columns), entries below 1e-13 dropped.
// The result is verified on a sample of rows (R_fine[l+1] T[l] == R_fine[l] to 1e-10 relative).
// ------------------------------------------------------------------------------------------------
namespace detail {
inline void abi_dims_ok(const mgbx_csr &A, const char *what) {
  if (!A.rowptr || (A.rowptr[A.rows] > 0 && (!A.colind || !A.val)) || A.rows < 0 || A.cols < 0 || A.cols > INT32_MAX)
    throw std::invalid_argument(std::string("level transfers are to be recovered from R_fine (T == NULL), but ") + what + " is missing");
}
}  // namespace detail

inline HostCsr recover_transfer(const mgbx_csr &Rn, const mgbx_csr &Rc) {
  detail::abi_dims_ok(Rn, "R_fine[l+1]");
  detail::abi_dims_ok(Rc, "R_fine[l]");
  if (Rn.rows != Rc.rows) throw std::invalid_argument("recover_transfer: R_fine[l] and R_fine[l+1] differ in rows");
  const int64_t mn = Rn.cols, mc = Rc.cols, rows = Rn.rows;
  HostCsr T;
  T.rows = mn;
  T.cols = mc;
  // ---- 1. selector rows
  std::vector<int64_t> sel(mn, -1);
  int64_t found = 0;
  for (int64_t i = 0; i < rows && found < mn; ++i)
    if (Rn.rowptr[i + 1] - Rn.rowptr[i] == 1) {
      const int64_t k = Rn.rowptr[i], j = Rn.colind[k];
      if (j >= 0 && j < mn && sel[j] < 0 && Rn.val[k] != 0.0) {
        sel[j] = i;
        ++found;
      }
    }
  bool done = false;
  if (found == mn) {
    T.ptr.assign(mn + 1, 0);
    for (int64_t j = 0; j < mn; ++j) T.ptr[j + 1] = T.ptr[j] + (Rc.rowptr[sel[j] + 1] - Rc.rowptr[sel[j]]);
    T.idx.resize(T.ptr[mn]);
    T.val.resize(T.ptr[mn]);
    for (int64_t j = 0; j < mn; ++j) {
      const int64_t i = sel[j];
      const double a = Rn.val[Rn.rowptr[i]];
      int64_t q = T.ptr[j];
      for (int64_t k = Rc.rowptr[i]; k < Rc.rowptr[i + 1]; ++k, ++q) {
        if (Rc.colind[k] < 0 || Rc.colind[k] >= mc) throw std::invalid_argument("recover_transfer: column index out of range");
        T.idx[q] = (int32_t)Rc.colind[k];
        T.val[q] = (a == 1.0) ? Rc.val[k] : Rc.val[k] / a;
      }
    }
    done = true;
  }
  HostCsr Hn, Hc;
  if (!done) {
    Hn = csr_from_abi(Rn);
    Hc = csr_from_abi(Rc);
  }
  // ---- 2. column matching
  if (!done && mc <= mn) {
    const HostCsr Tn = transpose(Hn), Tc = transpose(Hc);   // rows = columns of R
    auto hash_row = [](const HostCsr &A, int64_t r) {
      uint64_t hsh = 1469598103934665603ull;
      for (int64_t k = A.ptr[r]; k < A.ptr[r + 1]; ++k) {
        uint64_t bits;
        memcpy(&bits, &A.val[k], 8);
        hsh = (hsh ^ (uint64_t)(uint32_t)A.idx[k]) * 1099511628211ull;
        hsh = (hsh ^ bits) * 1099511628211ull;
      }
      return hsh;
    };
    std::vector<std::pair<uint64_t, int64_t>> hn(mn);
    for (int64_t j = 0; j < mn; ++j) hn[j] = {hash_row(Tn, j), j};
    std::sort(hn.begin(), hn.end());
    std::vector<int64_t> match(mc, -1);
    bool all = true;
    for (int64_t c = 0; c < mc && all; ++c) {
      const uint64_t hc = hash_row(Tc, c);
      auto it = std::lower_bound(hn.begin(), hn.end(), std::make_pair(hc, (int64_t)-1));
      for (; it != hn.end() && it->first == hc; ++it) {
        const int64_t j = it->second;
        const int64_t len = Tn.ptr[j + 1] - Tn.ptr[j];
        if (len != Tc.ptr[c + 1] - Tc.ptr[c]) continue;
        if (std::equal(Tn.idx.begin() + Tn.ptr[j], Tn.idx.begin() + Tn.ptr[j + 1], Tc.idx.begin() + Tc.ptr[c]) &&
            std::equal(Tn.val.begin() + Tn.ptr[j], Tn.val.begin() + Tn.ptr[j + 1], Tc.val.begin() + Tc.ptr[c])) {
          match[c] = j;
          break;
        }
      }
      all = match[c] >= 0;
    }
    if (all) {
      std::vector<std::pair<int64_t, int64_t>> ent(mc);   // (row j of T, column c)
      for (int64_t c = 0; c < mc; ++c) ent[c] = {match[c], c};
      std::sort(ent.begin(), ent.end());
      T.ptr.assign(mn + 1, 0);
      for (auto &e : ent) T.ptr[e.first + 1]++;
      for (int64_t j = 0; j < mn; ++j) T.ptr[j + 1] += T.ptr[j];
      T.idx.resize(mc);
      T.val.assign(mc, 1.0);
      for (int64_t q = 0; q < mc; ++q) T.idx[q] = (int32_t)ent[q].second;
      done = true;
    }
  }
  // ---- 3. dense normal equations
  if (!done) {
    if (mn > 2048)
      throw std::invalid_argument("recover_transfer: R_fine[l+1] has neither selector rows nor matching columns and is too large for the dense "
                                  "normal equations; pass the level transfers T explicitly");
    std::vector<double> G((size_t)mn * mn, 0.0), B((size_t)mn * mc, 0.0);
    for (int64_t i = 0; i < rows; ++i)
      for (int64_t a = Hn.ptr[i]; a < Hn.ptr[i + 1]; ++a) {
        const double va = Hn.val[a];
        double *g = G.data() + (size_t)Hn.idx[a] * mn;
        for (int64_t b = Hn.ptr[i]; b < Hn.ptr[i + 1]; ++b) g[Hn.idx[b]] += va * Hn.val[b];
        double *bb = B.data() + (size_t)Hn.idx[a] * mc;
        for (int64_t b = Hc.ptr[i]; b < Hc.ptr[i + 1]; ++b) bb[Hc.idx[b]] += va * Hc.val[b];
      }
    // Cholesky G = L L' (in place, lower), then solve for the mc right-hand sides
    for (int64_t j = 0; j < mn; ++j) {
      double d = G[(size_t)j * mn + j];
      for (int64_t k = 0; k < j; ++k) d -= G[(size_t)j * mn + k] * G[(size_t)j * mn + k];
      if (!(d > 0.0)) throw std::invalid_argument("recover_transfer: R_fine[l+1] is rank deficient; pass T explicitly");
      d = std::sqrt(d);
      G[(size_t)j * mn + j] = d;
      for (int64_t i = j + 1; i < mn; ++i) {
        double s = G[(size_t)i * mn + j];
        for (int64_t k = 0; k < j; ++k) s -= G[(size_t)i * mn + k] * G[(size_t)j * mn + k];
        G[(size_t)i * mn + j] = s / d;
      }
    }
    for (int64_t i = 0; i < mn; ++i) {   // forward
      double *bi = B.data() + (size_t)i * mc;
      for (int64_t k = 0; k < i; ++k) {
        const double l = G[(size_t)i * mn + k];
        if (l == 0.0) continue;
        const double *bk = B.data() + (size_t)k * mc;
        for (int64_t c = 0; c < mc; ++c) bi[c] -= l * bk[c];
      }
      const double d = G[(size_t)i * mn + i];
      for (int64_t c = 0; c < mc; ++c) bi[c] /= d;
    }
    for (int64_t i = mn - 1; i >= 0; --i) {   // backward with L'
      double *bi = B.data() + (size_t)i * mc;
      for (int64_t k = i + 1; k < mn; ++k) {
        const double l = G[(size_t)k * mn + i];
        if (l == 0.0) continue;
        const double *bk = B.data() + (size_t)k * mc;
        for (int64_t c = 0; c < mc; ++c) bi[c] -= l * bk[c];
      }
      const double d = G[(size_t)i * mn + i];
      for (int64_t c = 0; c < mc; ++c) bi[c] /= d;
    }
    double bmax = 0.0;
    for (double v : B) bmax = std::max(bmax, std::fabs(v));
    T.ptr.assign(1, 0);
    for (int64_t i = 0; i < mn; ++i) {
      for (int64_t c = 0; c < mc; ++c) {
        const double v = B[(size_t)i * mc + c];
        if (std::fabs(v) > 1e-13 * bmax) {
          T.idx.push_back((int32_t)c);
          T.val.push_back(v);
        }
      }
      T.ptr.push_back((int64_t)T.idx.size());
    }
  }
  // ---- verification on a sample of rows:  (Rn T)[i, :] == Rc[i, :]
  {
    std::vector<double> acc(mc, 0.0);
    const int64_t stride = std::max<int64_t>(1, rows / 512);
    double err = 0.0, ref = 0.0;
    for (int64_t i = 0; i < rows; i += stride) {
      for (int64_t a = Rn.rowptr[i]; a < Rn.rowptr[i + 1]; ++a) {
        const int64_t j = Rn.colind[a];
        for (int64_t q = T.ptr[j]; q < T.ptr[j + 1]; ++q) acc[T.idx[q]] += Rn.val[a] * T.val[q];
      }
      for (int64_t b = Rc.rowptr[i]; b < Rc.rowptr[i + 1]; ++b) {
        acc[Rc.colind[b]] -= Rc.val[b];
        ref = std::max(ref, std::fabs(Rc.val[b]));
      }
      for (int64_t a = Rn.rowptr[i]; a < Rn.rowptr[i + 1]; ++a) {
        const int64_t j = Rn.colind[a];
        for (int64_t q = T.ptr[j]; q < T.ptr[j + 1]; ++q) {
          err = std::max(err, std::fabs(acc[T.idx[q]]));
          acc[T.idx[q]] = 0.0;
        }
      }
      for (int64_t b = Rc.rowptr[i]; b < Rc.rowptr[i + 1]; ++b) {
        err = std::max(err, std::fabs(acc[Rc.colind[b]]));
        acc[Rc.colind[b]] = 0.0;
      }
    }
    if (err > 1e-10 * std::max(ref, 1e-300))
      throw std::invalid_argument("recover_transfer: R_fine[l] is not in the range of R_fine[l+1] (the hierarchy is not nested); pass T explicitly");
  }
  return T;
}

// First column of every state variable at one level, read off the block structure of R_fine[l] (rows [v n, (v+1) n)
// of the block-diagonal prolongation only touch the columns of variable v, src/multigrid.jl:474-538).
inline std::vector<int64_t> derive_var_offsets(const mgbx_csr &R, int nu, int64_t n) {
  detail::abi_dims_ok(R, "R_fine[l] (needed to derive var_offsets)");
  if (R.rows != (int64_t)nu * n) throw std::invalid_argument("derive_var_offsets: R_fine[l] must have nu*n rows");
  std::vector<int64_t> off(nu + 1, 0);
  for (int v = 0; v < nu; ++v) {
    int64_t lo = INT64_MAX, hi = -1;
    for (int64_t k = R.rowptr[(int64_t)v * n]; k < R.rowptr[(int64_t)(v + 1) * n]; ++k) {
      lo = std::min<int64_t>(lo, R.colind[k]);
      hi = std::max<int64_t>(hi, R.colind[k]);
    }
    if (hi >= 0 && lo < off[v]) throw std::invalid_argument("derive_var_offsets: R_fine[l] is not block diagonal in state-variable order");
    off[v + 1] = std::max(off[v], hi + 1);
  }
  if (off[nu] != R.cols) throw std::invalid_argument("derive_var_offsets: R_fine[l] has columns without entries (not a full-rank prolongation)");
  return off;
}

// The reference plan's output pattern: union over elements of cols(e) x cols(e)
// (src/BlockMatrices.jl:381-446); symmetric, so CSC == CSR.
inline HostCsr plan_pattern(const HostCsr &Einc) { return spgemm_symbolic(transpose(Einc), Einc); }

}  // namespace mgbx
