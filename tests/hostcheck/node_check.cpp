// TEST-ONLY shim (never part of libmgbx.so, never on a product path): evaluates the
// __host__ __device__ per-node barrier arithmetic of csrc/node_barrier.cuh on the CPU so that the
// formulas can be checked against the oracle without a GPU.  Built by tests/test_node_math.py.
#include <cstring>

#include "../../multigridbarrier.jl_b200/csrc/node_barrier.cuh"

using namespace mgbx;

static int eval_impl(const mgbx_convex *Q, int64_t n, int ny, int feas, int NC, double fb, double fR, const double *Y, int order, double *F0,
                     double *F1, double *F2, double *slack, unsigned schur_mask, int identity_A);

extern "C" int hostcheck_node_eval(const mgbx_convex *Q, int64_t n, int ny, int feas, int NC, double fb, double fR,
                                   const double *Y /* n x ny col-major */, int order, double *F0, double *F1 /* n x ny */,
                                   double *F2 /* n x ny x ny, [i + (a*ny+b)*n] */, double *slack /* n or null */) {
  return eval_impl(Q, n, ny, feas, NC, fb, fR, Y, order, F0, F1, F2, slack, 0u, 0);
}

// the same with the pieces of `schur_mask` condensed analytically (piece_eval's `schur`); identity_A: treat every piece's A as the
// identity and b as zero, as mgbx_create does after compressing constant grids away
extern "C" int hostcheck_node_eval_schur(const mgbx_convex *Q, int64_t n, int ny, const double *Y, double *F0, double *F1, double *F2,
                                         unsigned schur_mask, int identity_A) {
  return eval_impl(Q, n, ny, 0, ny + 1, 0.0, 0.0, Y, 2, F0, F1, F2, nullptr, schur_mask, identity_A);
}

static int eval_impl(const mgbx_convex *Q, int64_t n, int ny, int feas, int NC, double fb, double fR, const double *Y, int order, double *F0,
                     double *F1, double *F2, double *slack, unsigned schur_mask, int identity_A) {
  ConvexDev cd;
  memset(&cd, 0, sizeof(cd));
  cd.npieces = Q->npieces;
  for (int k = 0; k < Q->npieces; ++k) {
    const mgbx_piece &q = Q->pieces[k];
    PieceDev &d = cd.pc[k];
    d.kind = q.kind;
    d.ni = q.ni;
    d.nc = q.nc;
    for (int c = 0; c < q.ni; ++c) d.idx[c] = q.idx ? q.idx[c] : c;
    d.A = identity_A ? nullptr : q.A;
    d.b = identity_A ? nullptr : q.b;
    d.p = q.p;
    d.mu = q.mu;
    d.p_uniform = 2.0;
    d.mu_uniform = 0.0;
    if (identity_A && q.p == nullptr) d.p_uniform = 1.0;
  }
  cd.select = Q->select;
  cd.feas = feas;
  cd.NC = NC;
  cd.NF = ny;
  cd.fb = fb;
  cd.fR = fR;
  for (int64_t i = 0; i < n; ++i) {
    double y[MGBX_MAX_ND], f1[MGBX_MAX_ND], f2[MGBX_MAX_ND * MGBX_MAX_ND];
    for (int k = 0; k < ny; ++k) y[k] = Y[i + (int64_t)k * n];
    F0[i] = node_eval(cd, n, i, y, order, f1, f2, schur_mask);
    if (order >= 1)
      for (int k = 0; k < ny; ++k) F1[i + (int64_t)k * n] = f1[k];
    if (order >= 2)
      for (int k = 0; k < ny * ny; ++k) F2[i + (int64_t)k * n] = f2[k];
    if (slack) slack[i] = node_slack(cd, n, i, y);
  }
  return 0;
}
