// node_barrier.cuh -- per-quadrature-point barrier value / gradient / Hessian.
//
// Device restatement of the reference's per-node functors (all lines under /root/reference/src):
//   utils.jl:14                         Log (-Inf outside the domain)
//   convex_linear.jl:388-390            _safe_pow = exp(a*Log(s))
//   convex_euclidian_power.jl:79-145    EuclidianPowerBarrier{,Grad,Hess}; core :387-433
//   convex_euclidian_power.jl:159-253   cobarrier (slack added to s) and slack functor
//   convex_linear.jl:119-214            linear barrier / cobarrier / slack
//   convex_piecewise.jl:15-75           sum over the selected pieces; unselected pieces are NOT evaluated
//   mgb.jl:217-287                      _feasibility_convex (phase-I wrapper)
// The functions are __host__ __device__ so that tests can check the arithmetic on the CPU through
// a test-only shim (tests/hostcheck); the product only ever calls them from kernels.
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/mgbx.h"

#ifdef __CUDACC__
#define MGBX_HD __host__ __device__ __forceinline__
#else
#define MGBX_HD inline
#endif

namespace mgbx {

struct PieceDev {
  int kind, ni, nc;
  int idx[MGBX_MAX_NI];
  const double *A, *b, *p, *mu;   // SoA grids (column stride n); A == nullptr: identity, b == nullptr: zero
  double p_uniform, mu_uniform;   // used when p == nullptr
};

struct ConvexDev {
  int npieces;
  PieceDev pc[MGBX_MAX_PIECES];
  const double *select;           // n x npieces (column stride n) or nullptr
  int feas;                       // 1: phase-I wrapper around the pieces
  int NC;                         // feasibility: user rows + 1; the slack is input NC-1
  int NF;                         // number of node inputs (= nD of the AMG this set is evaluated on)
  double fb, fR;                  // feasibility: slack bound b, box radius R
};

MGBX_HD double Log(double x) { return x <= 0.0 ? -INFINITY : log(x); }
MGBX_HD double safe_pow(double s, double a) { return exp(a * Log(s)); }

// One piece.  y: node inputs; F1 (ny) and F2 (ny x ny, row stride ny) are accumulated into.
// The Hessian of the piece in its own coordinates z = A y[idx] + b is never stored: entry (r, s) is formed on
// the fly from a handful of scalars (EP: 4 z_r z_s / rho^2 + 2 delta_rs / rho, the coupling column and the
// slack corner; LINEAR: diag(1/z^2)), so the common identity-A case touches no local array beyond z and g.
// schur: this piece's slack row is eliminated node-locally by the caller and nothing else touches that row -- the q-q block is then
// returned ALREADY CONDENSED, (2/rho) I + (4/rho^2) (B / (A + B)) z z' with H_ss = A + B, A = (al s^(al-1) / rho)^2 (the square of
// the coupling): algebraically H_qq - H_qs H_sq / H_ss, but without subtracting two terms of size ~t^2 whose difference is O(1)
// along z (the subtraction leaves no correct digit there at t ~ 1e8).  The coupling column and H_ss are still returned for the
// right-hand-side condensation and the back-substitution; the caller must skip this slack in its numerical Schur loop.
MGBX_HD double piece_eval(const PieceDev &pc, int64_t n, int64_t i, const double *y, int ny, int order,
                          bool cob, double slack, int slackpos, double *F1, double *F2, bool schur = false) {
  const int ni = pc.ni, nc = pc.nc;
  double z[MGBX_MAX_NC];
  double gz[MGBX_MAX_NC];
  double Al[MGBX_MAX_NC * MGBX_MAX_NI];
  const bool Aid = (pc.A == nullptr);
  if (!Aid)
    for (int c = 0; c < ni; ++c)
      for (int r = 0; r < nc; ++r) Al[r + c * nc] = pc.A[i + (int64_t)(c * nc + r) * n];
  for (int r = 0; r < nc; ++r) {
    double acc = 0.0;
    if (Aid) acc = y[pc.idx[r]];
    else
      for (int c = 0; c < ni; ++c) acc += Al[r + c * nc] * y[pc.idx[c]];
    z[r] = acc + (pc.b ? pc.b[i + (int64_t)r * n] : 0.0);
  }
  double f0;
  const bool ep = (pc.kind == MGBX_PIECE_EP);
  const int nq = nc - 1;
  double inv_r = 0.0, inv_r2 = 0.0, coef = 0.0, hss = 0.0, qqs = 1.0;
  if (ep) {
    if (cob) z[nq] += slack;
    const double s = z[nq];
    const double p = pc.p ? pc.p[i] : pc.p_uniform;
    const double mu = pc.mu ? pc.mu[i] : pc.mu_uniform;
    const double al = 2.0 / p;
    double qsq = 0.0;
    for (int k = 0; k < nq; ++k) qsq += z[k] * z[k];
    // one log and one exp serve every power of s on the interior (s > 0): s^(a-1) = s^a / s, s^(a-2) = s^(a-1) / s,
    // s^(2a-2) = (s^(a-1))^2; outside the domain the reference's _safe_pow values are formed as written there
    const double ls = Log(s);
    const double sa = exp(al * ls);
    const double r = sa - qsq;
    f0 = -Log(r) - mu * ls;
    if (order >= 1) {
      inv_r = 1.0 / r;
      const bool in = s > 0.0;
      const double sam1 = in ? sa / s : safe_pow(s, al - 1.0);
      for (int k = 0; k < nq; ++k) gz[k] = 2.0 * inv_r * z[k];
      gz[nq] = -al * sam1 * inv_r - mu / s;
      if (order >= 2) {
        inv_r2 = inv_r * inv_r;
        coef = -2.0 * al * sam1 * inv_r2;
        const double sam2 = in ? sam1 / s : safe_pow(s, al - 2.0);
        const double s2am2 = in ? sam1 * sam1 : safe_pow(s, 2.0 * al - 2.0);
        const double hA = al * al * s2am2 * inv_r2, hB = -al * (al - 1.0) * sam2 * inv_r + mu / (s * s);
        hss = hA + hB;
        if (schur) qqs = hB / hss;
      }
    }
  } else {
    f0 = 0.0;
    for (int r = 0; r < nc; ++r) {
      if (cob) z[r] += slack;
      f0 -= Log(z[r]);
      if (order >= 1) gz[r] = -1.0 / z[r];
    }
  }
  // Hessian entry (r, s) in piece coordinates
  auto hz = [&](int r, int s2) -> double {
    if (ep) {
      if (r < nq && s2 < nq) return 4.0 * z[r] * z[s2] * inv_r2 * qqs + (r == s2 ? 2.0 * inv_r : 0.0);
      if (r == nq && s2 == nq) return hss;
      return coef * z[r < s2 ? r : s2];
    }
    return r == s2 ? 1.0 / (z[r] * z[r]) : 0.0;
  };
  if (order >= 1) {
    for (int c = 0; c < ni; ++c) {
      double acc = 0.0;
      if (Aid) acc = gz[c];
      else
        for (int r = 0; r < nc; ++r) acc += Al[r + c * nc] * gz[r];
      F1[pc.idx[c]] += acc;
    }
    if (cob) {
      double acc = 0.0;
      if (ep) acc = gz[nq];
      else
        for (int r = 0; r < nc; ++r) acc += gz[r];
      F1[slackpos] += acc;
    }
  }
  if (order >= 2) {
    for (int a = 0; a < ni; ++a)
      for (int b = 0; b < ni; ++b) {
        double acc = 0.0;
        if (Aid) acc = hz(a, b);
        else if (ep) {
          for (int r = 0; r < nc; ++r) {
            double t = 0.0;
            for (int s2 = 0; s2 < nc; ++s2) t += hz(r, s2) * Al[s2 + b * nc];
            acc += Al[r + a * nc] * t;
          }
        } else {
          for (int r = 0; r < nc; ++r) acc += Al[r + a * nc] * hz(r, r) * Al[r + b * nc];
        }
        F2[pc.idx[a] * ny + pc.idx[b]] += acc;
      }
    if (cob) {
      // Hz * e_sl: EP -> last column of Hz, LINEAR -> the diagonal
      double corner = 0.0;
      for (int a = 0; a < ni; ++a) {
        double acc = 0.0;
        if (Aid) acc = ep ? hz(a, nq) : hz(a, a);
        else
          for (int r = 0; r < nc; ++r) acc += Al[r + a * nc] * (ep ? hz(r, nq) : hz(r, r));
        F2[pc.idx[a] * ny + slackpos] += acc;
        F2[slackpos * ny + pc.idx[a]] += acc;
      }
      if (ep) corner = hss;
      else
        for (int r = 0; r < nc; ++r) corner += hz(r, r);
      F2[slackpos * ny + slackpos] += corner;
    }
  }
  return f0;
}

// Sum of the selected pieces (+ the phase-I wrapper).  F1 / F2 are overwritten.
// schur_mask: bit k set = piece k is condensed analytically (see piece_eval)
MGBX_HD double node_eval(const ConvexDev &cd, int64_t n, int64_t i, const double *y, int order, double *F1,
                         double *F2, unsigned schur_mask = 0u) {
  const int ny = cd.NF;
  if (order >= 1)
    for (int k = 0; k < ny; ++k) F1[k] = 0.0;
  if (order >= 2)
    for (int k = 0; k < ny * ny; ++k) F2[k] = 0.0;
  double F0 = 0.0;
  const bool cob = cd.feas != 0;
  const int slackpos = cd.NC - 1;
  const double u = cob ? y[slackpos] : 0.0;
  for (int k = 0; k < cd.npieces; ++k) {
    if (cd.select && cd.select[i + (int64_t)k * n] == 0.0) continue;
    F0 += piece_eval(cd.pc[k], n, i, y, ny, order, cob, u, slackpos, F1, F2, ((schur_mask >> k) & 1u) != 0u);
  }
  if (cob) {
    const double bb = cd.fb, RR = cd.fR;
    F0 += -Log(bb - u) - Log(bb + u);
    if (order >= 1) F1[slackpos] += 1.0 / (bb - u) - 1.0 / (bb + u);
    if (order >= 2) F2[slackpos * ny + slackpos] += 1.0 / ((bb - u) * (bb - u)) + 1.0 / ((bb + u) * (bb + u));
    for (int k = cd.NC; k < ny; ++k) {
      const double v = y[k];
      F0 += -Log(RR - v) - Log(RR + v);
      if (order >= 1) F1[k] = 1.0 / (RR - v) - 1.0 / (RR + v);
      if (order >= 2) F2[k * ny + k] = 1.0 / ((RR - v) * (RR - v)) + 1.0 / ((RR + v) * (RR + v));
    }
  }
  return F0;
}

// Slack functor (convex_euclidian_power.jl:240-253, convex_linear.jl:196-214, convex_piecewise.jl:65-75):
// the smallest t such that y becomes feasible when t is added to the slack slot; max over pieces.
MGBX_HD double node_slack(const ConvexDev &cd, int64_t n, int64_t i, const double *y) {
  double out = -INFINITY;
  for (int k = 0; k < cd.npieces; ++k) {
    if (cd.select && cd.select[i + (int64_t)k * n] == 0.0) continue;
    const PieceDev &pc = cd.pc[k];
    const int ni = pc.ni, nc = pc.nc;
    double z[MGBX_MAX_NC];
    for (int r = 0; r < nc; ++r) {
      double acc = 0.0;
      if (pc.A == nullptr) acc = y[pc.idx[r]];
      else
        for (int c = 0; c < ni; ++c) acc += pc.A[i + (int64_t)(c * nc + r) * n] * y[pc.idx[c]];
      z[r] = acc + (pc.b ? pc.b[i + (int64_t)r * n] : 0.0);
    }
    double sl;
    if (pc.kind == MGBX_PIECE_EP) {
      const int nq = nc - 1;
      double qsq = 0.0;
      for (int r = 0; r < nq; ++r) qsq += z[r] * z[r];
      const double p = pc.p ? pc.p[i] : pc.p_uniform;
      const double s = z[nq];
      sl = -fmin(s - safe_pow(qsq, p / 2.0), s);
    } else {
      double m = INFINITY;
      for (int r = 0; r < nc; ++r) m = fmin(m, z[r]);
      sl = -m;
    }
    out = fmax(out, sl);
  }
  return out;
}

}  // namespace mgbx
