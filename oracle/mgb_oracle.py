"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY (not product code, never shipped, never timed as product).

A NumPy/SciPy restatement of the reference's barrier-Newton hot path, used as the checker for
the CUDA library: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
arm may import it.  The product (multigridbarrier.jl_b200/) must never import this file.

Parity pin: this restatement is checked against every end-to-end golden vector the reference's
own tests hold for the path (tests/golden/reference_goldens.json, extracted from
test/runtests.jl:13-52 and test/test_algebraic.jl:38-69 by tests/golden/make_golden.py) in
tests/test_oracle_golden.py.  Third-party arithmetic behind the reference that is NOT under
/root/reference (AlgebraicMultigrid.jl ruge_stuben, CHOLMOD `\\`) is replaced by the in-repo
classical RS of hierarchy.py and SciPy SuperLU / LAPACK; the goldens are hierarchy-independent
(test/test_algebraic.jl:18-31), so they pin the converged z, not the prolongator entries.

Reference lines restated (all under /root/reference/src):
  utils.jl:14                         Log
  convex_linear.jl:388-390            _safe_pow
  convex_euclidian_power.jl:79-253    EP barrier / cobarrier / slack functors, :387-433 core grad/hess
  convex_linear.jl:119-214            linear barrier / cobarrier / slack
  convex_piecewise.jl:15-75           piecewise sums (unselected pieces are not evaluated)
  mgb.jl:217-287                      _feasibility_convex
  convex.jl:125,147-257,279-304       apply_D, barrier(Q) f0/f1/f2, masked barrier, barrier weights
  BlockMatrices.jl:322-446            assembly-plan output pattern
  newton.jl:4-27,35-50,84-103,139-154,187,222-287   illinois, line searches, stopping rules, newton
  mgb.jl:10-82,91-183,307-330,332-584               mgb_step, mgb_core, _matched_t, mgb_driver
  Parabolic.jl:126-173                parabolic_solve
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, List, Optional

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

EPS = np.finfo(np.float64).eps
KIND_EP, KIND_LINEAR = 0, 1


class MGBConvergenceFailure(Exception):
    """utils.jl:157-184; code in {'infeasible','feasibility_Rmax','stall','iteration_limit','failure'}."""

    def __init__(self, message, code="failure"):
        super().__init__(message)
        self.message = message
        self.code = code


# ------------------------------------------------------------------------------------------
# scalar conventions
# ------------------------------------------------------------------------------------------

def Log(x):
    """utils.jl:14 -- -Inf outside the domain."""
    x = np.asarray(x, dtype=np.float64)
    with np.errstate(all="ignore"):
        return np.where(x > 0, np.log(np.where(x > 0, x, 1.0)), -np.inf)


def safe_pow(s, a):
    """convex_linear.jl:388-390 -- exp(a*Log(s)); IEEE semantics (0*-Inf = NaN) kept."""
    with np.errstate(all="ignore"):
        return np.exp(a * Log(s))


# ------------------------------------------------------------------------------------------
# per-node functors, vectorised over nodes.  Y is (n, ny); with cobarrier=True the last column
# of Y is the slack and idx refers to the leading ny-1 columns.
# Returns (F0 (n,), F1 (n,ny), F2 (n,ny,ny)) up to `order`.
# ------------------------------------------------------------------------------------------

def _piece_cols(pc, ny_user):
    return list(range(pc.ni)) if pc.idx is None else list(pc.idx)


def _ep_core(q, s, p, mu, order):
    """convex_euclidian_power.jl:387-433."""
    with np.errstate(all="ignore"):
        al = 2.0 / p
        qsq = np.sum(q * q, axis=1)
        sa = safe_pow(s, al)
        r = sa - qsq
        f0 = -Log(r) - mu * Log(s)
        if order == 0:
            return f0, None, None
        inv_r = 1.0 / r
        sam1 = safe_pow(s, al - 1.0)
        gq = (2.0 * inv_r)[:, None] * q
        gs = -al * sam1 * inv_r - mu / s
        g = np.concatenate([gq, gs[:, None]], axis=1)
        if order == 1:
            return f0, g, None
        inv_r2 = inv_r * inv_r
        coef = -2.0 * al * sam1 * inv_r2
        sam2 = safe_pow(s, al - 2.0)
        s2am2 = safe_pow(s, 2.0 * al - 2.0)
        hss = -al * (al - 1.0) * sam2 * inv_r + al * al * s2am2 * inv_r2 + mu / (s * s)
        n, nq = q.shape
        H = np.zeros((n, nq + 1, nq + 1))
        H[:, :nq, :nq] = 4.0 * q[:, :, None] * q[:, None, :] * inv_r2[:, None, None]
        for i in range(nq):
            H[:, i, i] += 2.0 * inv_r
        H[:, :nq, nq] = coef[:, None] * q
        H[:, nq, :nq] = coef[:, None] * q
        H[:, nq, nq] = hss
        return f0, g, H


def piece_eval(pc, Y, order, cobarrier=False):
    n, ny = Y.shape
    ny_user = ny - 1 if cobarrier else ny
    cols = _piece_cols(pc, ny_user)
    yi = Y[:, cols]
    ni, nc = pc.ni, pc.nc
    A3 = pc.A.reshape(n, ni, nc).transpose(0, 2, 1)          # A3[i, r, c] = A_row[c*nc + r]
    with np.errstate(all="ignore"):
        z = np.einsum("irc,ic->ir", A3, yi) + pc.b
        if pc.kind == KIND_EP:
            if cobarrier:
                z = z.copy()
                z[:, -1] = z[:, -1] + Y[:, -1]
            f0, gz, Hz = _ep_core(z[:, :-1], z[:, -1], pc.p, pc.mu, order)
            e_sl = np.zeros(nc)
            e_sl[-1] = 1.0
        else:
            if cobarrier:
                z = z + Y[:, -1:]
            f0 = -np.sum(Log(z), axis=1)
            gz = Hz = None
            if order >= 1:
                gz = -1.0 / z
            if order >= 2:
                Hz = np.zeros((n, nc, nc))
                d = 1.0 / (z * z)
                for r in range(nc):
                    Hz[:, r, r] = d[:, r]
            e_sl = np.ones(nc)
        F1 = F2 = None
        if order >= 1:
            F1 = np.zeros((n, ny))
            gi = np.einsum("irc,ir->ic", A3, gz)
            for k, cidx in enumerate(cols):
                F1[:, cidx] = gi[:, k]
            if cobarrier:
                F1[:, -1] = gz @ e_sl
        if order >= 2:
            F2 = np.zeros((n, ny, ny))
            HA = np.einsum("irs,isc->irc", Hz, A3)
            Hi = np.einsum("ird,irc->idc", A3, HA)
            for a, ca in enumerate(cols):
                for b, cb in enumerate(cols):
                    F2[:, ca, cb] = Hi[:, a, b]
            if cobarrier:
                cross = np.einsum("irc,ir->ic", A3, Hz @ e_sl)
                for a, ca in enumerate(cols):
                    F2[:, ca, -1] = cross[:, a]
                    F2[:, -1, ca] = cross[:, a]
                F2[:, -1, -1] = np.einsum("r,irs,s->i", e_sl, Hz, e_sl)
    return f0, F1, F2


def piece_slack(pc, Y):
    n, ny = Y.shape
    cols = _piece_cols(pc, ny)
    A3 = pc.A.reshape(n, pc.ni, pc.nc).transpose(0, 2, 1)
    with np.errstate(all="ignore"):
        z = np.einsum("irc,ic->ir", A3, Y[:, cols]) + pc.b
        if pc.kind == KIND_EP:
            q, s = z[:, :-1], z[:, -1]
            qsq = np.sum(q * q, axis=1)
            return -np.minimum(s - safe_pow(qsq, pc.p / 2.0), s)
        return -np.min(z, axis=1)


def convex_eval(Q, Y, order, cobarrier=False):
    """Sum over the pieces selected at each node (convex_piecewise.jl:15-63)."""
    n, ny = Y.shape
    F0 = np.zeros(n)
    F1 = np.zeros((n, ny)) if order >= 1 else None
    F2 = np.zeros((n, ny, ny)) if order >= 2 else None
    for k, pc in enumerate(Q.pieces):
        f0, f1, f2 = piece_eval(pc, Y, order, cobarrier)
        if Q.select is not None:
            m = Q.select[:, k] != 0
            f0 = np.where(m, f0, 0.0)
            if f1 is not None:
                f1 = np.where(m[:, None], f1, 0.0)
            if f2 is not None:
                f2 = np.where(m[:, None, None], f2, 0.0)
        F0 = F0 + f0
        if order >= 1:
            F1 = F1 + f1
        if order >= 2:
            F2 = F2 + f2
    return F0, F1, F2


def convex_slack(Q, Y):
    """convex_piecewise.jl:65-75: max over selected pieces (typemin for unselected)."""
    out = np.full(Y.shape[0], -np.inf)
    for k, pc in enumerate(Q.pieces):
        s = piece_slack(pc, Y)
        if Q.select is not None:
            s = np.where(Q.select[:, k] != 0, s, -np.inf)
        out = np.maximum(out, s)
    return out


@dataclass
class FeasibilityConvex:
    """mgb.jl:217-287: cobarrier(yy[:NC]) - Log(b-u) - Log(b+u) - sum_i[Log(R-v_i)+Log(R+v_i)]."""
    Q: object
    b: float
    R: float
    NC: int


def feasibility_eval(FQ: FeasibilityConvex, YY, order):
    n, NF = YY.shape
    NC, bb, RR = FQ.NC, FQ.b, FQ.R
    yc = YY[:, :NC]
    u = yc[:, NC - 1]
    c0, c1, c2 = convex_eval(FQ.Q, yc, order, cobarrier=True)
    with np.errstate(all="ignore"):
        F0 = c0 - Log(bb - u) - Log(bb + u)
        for i in range(NC, NF):
            v = YY[:, i]
            F0 = F0 + (-Log(RR - v) - Log(RR + v))
        F1 = F2 = None
        if order >= 1:
            F1 = np.zeros((n, NF))
            F1[:, :NC] = c1
            F1[:, NC - 1] += 1.0 / (bb - u) - 1.0 / (bb + u)
            for i in range(NC, NF):
                v = YY[:, i]
                F1[:, i] = 1.0 / (RR - v) - 1.0 / (RR + v)
        if order >= 2:
            F2 = np.zeros((n, NF, NF))
            F2[:, :NC, :NC] = c2
            F2[:, NC - 1, NC - 1] += 1.0 / (bb - u) ** 2 + 1.0 / (bb + u) ** 2
            for i in range(NC, NF):
                v = YY[:, i]
                F2[:, i, i] = 1.0 / (RR - v) ** 2 + 1.0 / (RR + v) ** 2
    return F0, F1, F2


def node_eval(Q, Y, order):
    if isinstance(Q, FeasibilityConvex):
        return feasibility_eval(Q, Y, order)
    return convex_eval(Q, Y, order, cobarrier=False)


# ------------------------------------------------------------------------------------------
# operators
# ------------------------------------------------------------------------------------------

class Operators:
    """D_fine of one AMG as explicit (sparse or dense) matrices; D[k] is n x (nu n)
    (multigrid.jl:504-510, BlockMatrices.jl:604-640)."""

    def __init__(self, M):
        geom = M.geometry
        self.n = n = geom.n
        self.nu, self.nD = M.nu, M.nD
        self.dense = M.dense
        self.D = []
        for (var, op) in M.D:
            blocks = geom.operators[op]
            if self.dense:
                Dk = np.zeros((n, self.nu * n))
                Dk[:, var * n:(var + 1) * n] = blocks[0]
            else:
                N, V, _ = blocks.shape
                B = sp.bsr_matrix((blocks, np.arange(N), np.arange(N + 1)), shape=(n, n)).tocsr()
                B.sort_indices()
                Dk = sp.hstack([B if k == var else sp.csr_matrix((n, n)) for k in range(self.nu)],
                               format="csr")
            self.D.append(Dk)
        self.Dstack = np.vstack(self.D) if self.dense else sp.vstack(self.D, format="csr")

    def apply(self, z):
        """apply_D (convex.jl:125): (n, nD)."""
        return (self.Dstack @ z).reshape(self.nD, self.n).T

    def apply_t(self, Y):
        """sum_k D_k' Y[:, k]."""
        return self.Dstack.T @ Y.T.reshape(-1)

    def hessian(self, Hn):
        """sum_{j,k} D_j' diag(Hn[:, j, k]) D_k on the broken basis (convex.jl:185-200)."""
        n, nD = self.n, self.nD
        if self.dense:
            out = np.zeros((self.nu * n, self.nu * n))
            for j in range(nD):
                for k in range(nD):
                    if np.any(Hn[:, j, k] != 0) or True:
                        out += self.D[j].T @ (Hn[:, j, k][:, None] * self.D[k])
            return out
        rows, cols, vals = [], [], []
        ar = np.arange(n)
        for j in range(nD):
            for k in range(nD):
                rows.append(j * n + ar)
                cols.append(k * n + ar)
                vals.append(Hn[:, j, k])
        W = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                          shape=(nD * n, nD * n))
        return (self.Dstack.T @ W @ self.Dstack).tocsr()


_OPS_CACHE = {}


def operators(M) -> Operators:
    key = id(M)
    hit = _OPS_CACHE.get(key)
    if hit is None or hit[0] is not M:
        _OPS_CACHE[key] = (M, Operators(M))
    return _OPS_CACHE[key][1]


# ------------------------------------------------------------------------------------------
# barrier functional (convex.jl:147-257)
# ------------------------------------------------------------------------------------------

class Barrier:
    def __init__(self, Q, barrier_weights=None):
        self.Q = Q
        self.bw = barrier_weights

    def _weights(self, n):
        if self.bw is None:
            return None
        return self.bw

    def f0(self, s, w, c, R, ops: Operators, z0):
        Dz = ops.apply(z0 + R @ s)
        y, _, _ = node_eval(self.Q, Dz, 0)
        with np.errstate(all="ignore"):
            if self.bw is None:
                bar = (1.0 / len(w)) * np.sum(y)
            else:
                bar = np.sum(np.where(self.bw == 0, 0.0, self.bw * y))
            return float(bar + np.sum(w * np.sum(c * Dz, axis=1)))

    def f1(self, s, w, c, R, ops: Operators, z0):
        Dz = ops.apply(z0 + R @ s)
        _, g, _ = node_eval(self.Q, Dz, 1)
        with np.errstate(all="ignore"):
            if self.bw is None:
                y = (1.0 / len(w)) * g + w[:, None] * c
            else:
                y = np.where((self.bw == 0)[:, None], 0.0, self.bw[:, None] * g) + w[:, None] * c
        return R.T @ ops.apply_t(y)

    def node_hessian(self, s, w, R, ops: Operators, z0):
        Dz = ops.apply(z0 + R @ s)
        _, _, H = node_eval(self.Q, Dz, 2)
        with np.errstate(all="ignore"):
            if self.bw is None:
                return (1.0 / len(w)) * H
            return np.where((self.bw == 0)[:, None, None], 0.0, self.bw[:, None, None] * H)

    def f2(self, s, w, c, R, ops: Operators, z0):
        Hb = ops.hessian(self.node_hessian(s, w, R, ops, z0))
        if ops.dense:
            Rd = R if isinstance(R, np.ndarray) else R.toarray()
            return Rd.T @ Hb @ Rd
        return (R.T @ Hb @ R).tocsr()


def barrier_weights(w, barrier_nodes=None):
    """convex.jl:279-304; barrier_nodes None -> default mask w != 0 (mgb.jl:377)."""
    n = len(w)
    if barrier_nodes is None:
        sel = (w != 0).astype(float)
    elif isinstance(barrier_nodes, str) and barrier_nodes == ":":
        return None
    else:
        bn = np.asarray(barrier_nodes)
        if bn.dtype == bool:
            if bn.size != n:
                raise ValueError("barrier_nodes mask has length %d but the mesh has %d nodes" % (bn.size, n))
            sel = bn.astype(float)
        else:
            if bn.size == 0:
                raise ValueError("barrier_nodes must select at least one node")
            sel = np.zeros(n)
            sel[bn] = 1.0
    m = sel.sum()
    if m <= 0:
        raise ValueError("barrier_nodes selects no nodes")
    if m == n:
        return None
    return sel / m


# ------------------------------------------------------------------------------------------
# assembly-plan output pattern (BlockMatrices.jl:344-446)
# ------------------------------------------------------------------------------------------

def hessian_pattern(M, J):
    """CSR (indptr, indices) of R_J' H R_J as the reference's plan defines it: the union over
    elements e and over state-variable pairs (a, b) that occur among the D rows of
    cols_a(e) x cols_b(e), cols_k(e) = sorted unique stored columns of R in the element's rows of
    block k.  Symmetric, so CSC == CSR.  J is 0-based."""
    R = sp.csr_matrix(M.R_fine[J])
    geom = M.geometry
    N, V = geom.N, geom.V
    n = N * V
    m = R.shape[1]
    vars_used = sorted({v for (v, _) in M.D})
    rows_all, cols_all = [], []
    percol = {}
    for k in vars_used:
        lst = []
        for e in range(N):
            lo, hi = R.indptr[k * n + e * V], R.indptr[k * n + (e + 1) * V]
            lst.append(np.unique(R.indices[lo:hi]))
        percol[k] = lst
    for a in vars_used:
        for b in vars_used:
            for e in range(N):
                ca, cb = percol[a][e], percol[b][e]
                if ca.size and cb.size:
                    rows_all.append(np.repeat(ca, cb.size))
                    cols_all.append(np.tile(cb, ca.size))
    if not rows_all:
        return np.zeros(m + 1, np.int64), np.zeros(0, np.int64)
    r = np.concatenate(rows_all)
    c = np.concatenate(cols_all)
    P = sp.csr_matrix((np.ones(r.size, np.float32), (r, c)), shape=(m, m))
    P.sum_duplicates()
    P.sort_indices()
    return P.indptr.astype(np.int64), P.indices.astype(np.int64)


# ------------------------------------------------------------------------------------------
# linear solve (utils.jl:142-145)
# ------------------------------------------------------------------------------------------

def solve_sym(H, g):
    if isinstance(H, np.ndarray):
        return np.linalg.solve(H, g)
    H = sp.csc_matrix(H)
    with np.errstate(all="ignore"):
        lu = spla.splu(H, permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0,
                       options=dict(SymmetricMode=True))
        return lu.solve(g)


# ------------------------------------------------------------------------------------------
# newton.jl
# ------------------------------------------------------------------------------------------

def illinois(f, a, b, fa=None, fb=None, maxit=10000):
    """newton.jl:4-27."""
    fa = f(a) if fa is None else fa
    fb = f(b) if fb is None else fb
    assert math.isfinite(fa) and math.isfinite(fb)
    if fa == 0:
        return a
    if fa * fb >= 0:
        return b
    for _ in range(maxit):
        c = (a * fb - b * fa) / (fb - fa)
        fc = f(c)
        assert math.isfinite(fc)
        if c <= min(a, b) or c >= max(a, b) or fc * fa == 0 or fc * fb == 0:
            return c
        if fb * fc < 0:
            a, fa = b, fb
        else:
            fa /= 2
        b, fb = c, fc
    raise RuntimeError("Illinois solver failed to converge.")


def _linesearch_loop(attempt, x, y, g, beta):
    """newton.jl:35-50: any exception in a trial = rejected trial."""
    s = 1.0
    xn, yn, gn = x, y, g
    while s > 0.0:
        try:
            xn, yn, gn, done = attempt(s)
            if done:
                break
        except KeyboardInterrupt:
            raise
        except Exception:
            pass
        s = s * beta
    return xn, yn, gn


def linesearch_backtracking(beta=0.5, c1=0.1):
    """newton.jl:139-154."""
    def ls(x, y, g, n, F0, F1):
        inc = float(g @ n)

        def attempt(s):
            xn = x - s * n
            stalled = np.linalg.norm(xn - x) == 0
            yn, gn = F0(xn), F1(xn)
            if not (math.isfinite(yn) and np.all(np.isfinite(gn))):
                raise FloatingPointError("line search: non-finite step")
            return xn, yn, gn, bool(stalled or yn <= y - c1 * inc * s)
        return _linesearch_loop(attempt, x, y, g, beta)
    return ls


def linesearch_illinois(beta=0.5):
    """newton.jl:84-103."""
    def ls(x, y, g, n, F0, F1):
        inc = float(g @ n)

        def attempt(s):
            def phi(sig):
                xn = x - sig * n
                if not math.isfinite(F0(xn)):
                    raise FloatingPointError("line search: non-finite barrier value")
                return float(F1(xn) @ n)
            s2 = illinois(phi, 0.0, s, fa=inc)
            xn = x - s2 * n
            yn, gn = F0(xn), F1(xn)
            if not (math.isfinite(yn) and np.all(np.isfinite(gn))):
                raise FloatingPointError("line search: non-finite step")
            return xn, yn, gn, True
        return _linesearch_loop(attempt, x, y, g, beta)
    return ls


def stopping_exact(theta):
    """newton.jl:187."""
    return lambda ymin, yn, gmin, gn, n, ndecmin, ndec: bool(yn >= ymin and np.linalg.norm(gn) >= theta * gmin)


def stopping_inexact(lambda_tol, theta):
    """newton.jl:222-225."""
    ex = stopping_exact(theta)
    return lambda ymin, yn, gmin, gn, n, ndecmin, ndec: bool(ndec < lambda_tol or ex(ymin, yn, gmin, gn, n, ndecmin, ndec))


def newton(F0, F1, F2, x, maxit=10000, stopping_criterion=None, line_search=None, log=None):
    """newton.jl:227-287."""
    if stopping_criterion is None:
        stopping_criterion = stopping_exact(0.1)
    if line_search is None:
        line_search = linesearch_illinois()
    if not np.all(np.isfinite(x)):
        raise FloatingPointError("newton: initial point has non-finite entries")
    y = F0(x)
    if not math.isfinite(y):
        raise FloatingPointError("newton: initial objective value is not finite")
    ymin = y
    ys = [y]
    converged = False
    k = 0
    g = F1(x)
    if not np.all(np.isfinite(g)):
        raise FloatingPointError("newton: initial gradient has non-finite entries")
    gmin = float(np.linalg.norm(g))
    incmin = math.inf
    while k < maxit and not converged:
        k += 1
        H = F2(x)
        n = solve_sym(H, g)
        if not np.all(np.isfinite(n)):
            raise FloatingPointError("newton: Newton direction has non-finite entries")
        inc = float(g @ n)
        if log is not None:
            log("newton: k=%d y=%.17g |g|=%.6g lam2=%.6g" % (k, y, np.linalg.norm(g), inc))
        if inc <= 0:
            converged = abs(inc) <= EPS * max(abs(y), 1.0)
            break
        xn, yn, gn = line_search(x, y, g, n, F0, F1)
        if stopping_criterion(ymin, yn, gmin, gn, n, math.sqrt(incmin), math.sqrt(inc)):
            converged = True
        x, y, g = xn, yn, gn
        gmin = min(gmin, float(np.linalg.norm(g)))
        ymin = min(ymin, y)
        incmin = min(inc, incmin)
        ys.append(y)
    return dict(x=x, y=y, k=k, converged=converged, ys=ys)


# ------------------------------------------------------------------------------------------
# mgb.jl
# ------------------------------------------------------------------------------------------

def divide_and_conquer(eta, j, J):
    """mgb.jl:10-15."""
    if eta(j, J):
        return True
    jmid = (j + J) // 2
    if jmid == j or jmid == J:
        return False
    return divide_and_conquer(eta, j, jmid) and divide_and_conquer(eta, jmid, J)


def mgb_step(Q, M, z, c, maxit, max_newton, line_search, stopping_criterion, finalize,
             initial_step=False, barrier_weights=None, log=None, trace=None):
    """mgb.jl:16-82.  Levels are 1-based as in the reference; its[J-1] counts level J."""
    L = len(M.R_fine)
    B = Barrier(Q, barrier_weights)
    its = np.zeros(L, dtype=np.int64)
    w = M.w
    ops = operators(M)
    state = {"z": z}

    def eta(j, J, sc, mxit, ls):
        R = M.R_fine[J - 1]
        zJ = state["z"]
        s0 = np.zeros(R.shape[1])
        SOL = newton(lambda s: B.f0(s, w, c, R, ops, zJ),
                     lambda s: B.f1(s, w, c, R, ops, zJ),
                     lambda s: B.f2(s, w, c, R, ops, zJ),
                     s0, maxit=mxit, stopping_criterion=sc, line_search=ls, log=log)
        its[J - 1] += SOL["k"]
        if trace is not None:
            trace.append(dict(level=J, k=SOL["k"], converged=SOL["converged"], y=SOL["y"]))
        if SOL["converged"]:
            state["z"] = zJ + R @ SOL["x"]
        return SOL["converged"]

    def mn(j, J):
        return maxit if (initial_step and J - j == 1) else max_newton

    converged = divide_and_conquer(lambda j, J: eta(j, J, stopping_criterion, mn(j, J), line_search), 0, L)
    z_unfinalized = state["z"]
    its_finalize = 0
    if finalize is not None:
        before = int(its[L - 1])
        foo = eta(L - 1, L, finalize, maxit, line_search)
        its_finalize = int(its[L - 1]) - before        # bookkeeping only: the reference adds the finalize pass into its[L]
        converged = converged and foo
    return dict(z=state["z"], z_unfinalized=z_unfinalized, its=its, converged=converged, its_finalize=its_finalize)


def c_dot_Dz(M, c, z):
    """mgb.jl:135-136."""
    Dz = operators(M).apply(z)
    return float(sum(np.dot(M.w * c[:, j], Dz[:, j]) for j in range(M.nD)))


def mgb_core(Q, M, z, c, tol=math.sqrt(EPS), t=0.1, maxit=10000, kappa=10.0, early_stop=None,
             max_newton=None, finalize=None, barrier_weights=None, stopping_criterion=None,
             line_search=None, log=None, **_unused):
    """mgb.jl:91-183.  early_stop(z, t) -> bool."""
    if max_newton is None:
        max_newton = int(math.ceil(math.log2(-math.log2(EPS)) + 2))
    if early_stop is None:
        early_stop = lambda z, t: False
    target = 1.0 / tol
    kappa0 = kappa
    L = len(M.R_fine)
    its, ts, kappas, cdz = [], [], [], []
    kw = dict(max_newton=max_newton, maxit=maxit, barrier_weights=barrier_weights,
              stopping_criterion=stopping_criterion, line_search=line_search, log=log)
    SOL = mgb_step(Q, M, z, t * c, finalize=(finalize if t >= target else None), initial_step=True, **kw)
    if not SOL["converged"]:
        raise MGBConvergenceFailure("Initial centering failed in mgb_solve at t=%g, tol=%g, maxit=%d."
                                    % (t, tol, maxit), "stall")
    k = 1
    its.append(SOL["its"].copy())
    kappas.append(kappa)
    ts.append(t)
    z = SOL["z"]
    z_unfinalized = SOL["z_unfinalized"]
    its_finalize = SOL["its_finalize"]
    cdz.append(c_dot_Dz(M, c, z))
    while t < target and kappa > 1 and k < maxit and not early_stop(z, t):
        k += 1
        itk = np.zeros(L, dtype=np.int64)
        while kappa > 1:
            t1 = kappa * t
            SOL = mgb_step(Q, M, z, t1 * c, finalize=(finalize if t1 >= target else None), **kw)
            itk += SOL["its"]
            if SOL["converged"]:
                if SOL["its"].max() <= max_newton * 0.5:
                    kappa = min(kappa0, kappa ** 2)
                z = SOL["z"]
                z_unfinalized = SOL["z_unfinalized"]
                its_finalize = SOL["its_finalize"]
                t = t1
                break
            kappa = math.sqrt(kappa)
        its.append(itk)
        ts.append(t)
        kappas.append(kappa)
        cdz.append(c_dot_Dz(M, c, z))
    converged = (t >= target) or early_stop(z, t)
    if not converged:
        code = "stall" if kappa <= 1 else "iteration_limit"
        raise MGBConvergenceFailure("Convergence failure in mgb_solve at t=%g, k=%d, kappa=%g, tol=%g, maxit=%d."
                                    % (t, k, kappa, tol, maxit), code)
    return dict(z=z, z_unfinalized=z_unfinalized, c=c, its=np.stack(its, axis=1), ts=np.array(ts),
                kappas=np.array(kappas), c_dot_Dz=np.array(cdz), its_finalize=its_finalize)


def matched_t(Q, M, z, c, t_default, barrier_weights=None, log=None):
    """mgb.jl:307-330."""
    B = Barrier(Q, barrier_weights)
    L = len(M.R_fine)
    R = M.R_fine[L - 1]
    ops = operators(M)
    w = M.w
    s0 = np.zeros(R.shape[1])
    gphi = B.f1(s0, w, 0.0 * c, R, ops, z)
    gc = B.f1(s0, w, c, R, ops, z) - gphi
    H = B.f2(s0, w, c, R, ops, z)
    nphi = solve_sym(H, gphi)
    nc = solve_sym(H, gc)
    d = float(gc @ nc)
    b = float(gphi @ nc + gc @ nphi)
    if not d > 0:
        return t_default
    tstar = -b / (2 * d)
    if not (math.isfinite(tstar) and tstar > 0):
        return t_default
    tm = min(max(tstar, math.sqrt(EPS)), t_default)
    if log is not None:
        log("_matched_t: warm start matches t=%r, starting main ramp at t=%r" % (tstar, tm))
    return tm


def mgb_driver(Mpair, f, g, Q, t=0.1, t_feasibility=None, feasibility_Rmax=1.0 / math.sqrt(EPS),
               stopping_criterion=None, line_search=None, finalize="default", barrier_nodes=None,
               log=None, **rest):
    """mgb.jl:332-584."""
    M1, M2 = Mpair
    if t_feasibility is None:
        t_feasibility = t
    n = len(M1.w)
    if stopping_criterion is None:
        stopping_criterion = stopping_inexact(0.25 / math.sqrt(n), 0.9)
    if line_search is None:
        line_search = linesearch_backtracking()
    if isinstance(finalize, str) and finalize == "default":
        finalize = stopping_exact(0.9)
    if finalize is False:
        finalize = None
    if log is None:
        log = lambda *a: None
    bw_main = barrier_weights(M1.w, barrier_nodes)
    nD = M1.nD
    ncomp = g.shape[1]
    c0 = f
    z2 = g.T.reshape(-1).copy()                         # vcat of columns
    ops1 = operators(M1)
    Dz0 = ops1.apply(z2)
    SOL_feas = None
    F0, _, _ = convex_eval(Q, Dz0, 0)
    if not np.all(np.isfinite(F0)):
        sl = 2.0 * np.maximum(convex_slack(Q, Dz0), 1.0)
        b = 2.0 * max(1.0, float(sl.max()))
        c1 = np.zeros((n, nD + 1 + ncomp))
        c1[:, nD] = 1.0
        z1 = np.concatenate([z2, sl])
        feasible = lambda z: bool(np.max(z[ncomp * n:]) < 0)
        Rbox = max(10.0, 10.0 * float(np.max(np.abs(z2))))
        Rmax = max(float(feasibility_Rmax), Rbox)
        while True:
            log("mgb_driver: feasibility phase with bounding box R=%s" % _jlfloat(Rbox))
            Q_feas = FeasibilityConvex(Q, float(b), Rbox, nD + 1)
            failure = None
            t_first = [math.inf]

            def feas_stop(z, tt):
                if not feasible(z):
                    return False
                t_first[0] = min(t_first[0], tt)
                return tt >= 2 * t_first[0]
            try:
                kw = dict(rest)
                kw.update(early_stop=feas_stop, barrier_weights=None)
                SOL_feas = mgb_core(Q_feas, M2, z1, c1, t=t_feasibility,
                                    stopping_criterion=stopping_criterion, line_search=line_search,
                                    finalize=finalize, log=log, **kw)
            except KeyboardInterrupt:
                raise
            except Exception as e2:                    # broad on purpose (mgb.jl:510-521)
                failure = e2
            if failure is None:
                if feasible(SOL_feas["z"]):
                    break
                zf = SOL_feas["z"]
                vmax = max(float(np.max(np.abs(zf[k * n:(k + 1) * n]))) for k in range(ncomp))
                smax = float(np.max(zf[ncomp * n:]))
                if vmax <= Rbox / 2:
                    raise MGBConvergenceFailure(
                        "The problem appears to be infeasible: the feasibility subproblem converged to a "
                        "minimizer with positive constraint violation (max slack ~ %g) strictly inside the "
                        "bounding box (max |nodal value| ~ %g <= R/2 with R = %g)." % (smax, vmax, Rbox),
                        "infeasible")
                log("mgb_driver: phase-I minimizer presses the box; growing R")
            else:
                log("mgb_driver: feasibility solve failed at R=%s: %s" % (_jlfloat(Rbox), failure))
            Rnext = 10 * Rbox
            if Rnext > Rmax:
                raise MGBConvergenceFailure(
                    "Could not find a strictly feasible point with nodal values bounded by R = %g "
                    "(cap feasibility_Rmax ~ %g). The problem is infeasible, or its feasible points have "
                    "nodal values exceeding the cap (rescale the problem, or raise feasibility_Rmax)."
                    % (Rbox, Rmax), "feasibility_Rmax")
            Rbox = Rnext
        z2 = SOL_feas["z"][:z2.size]
        t = min(t, matched_t(Q, M1, z2, c0, t, barrier_weights=bw_main, log=log))
    SOL_main = mgb_core(Q, M1, z2, c0, t=t, stopping_criterion=stopping_criterion,
                        line_search=line_search, finalize=finalize, barrier_weights=bw_main,
                        log=log, **rest)
    z = SOL_main["z"].reshape(ncomp, n).T.copy()
    return dict(z=z, SOL_feasibility=SOL_feas, SOL_main=SOL_main)


def _jlfloat(x):
    """Julia prints 100.0 as '100.0'."""
    return repr(float(x))


def mgb_solve(prob, **kw):
    """mgb.jl:798-842 (CPU path: no device conversion)."""
    lines = []
    user_log = kw.pop("log", None)

    def log(*a):
        s = "".join(str(x) for x in a)
        lines.append(s)
        if user_log is not None:
            user_log(s)
    sol = mgb_driver(prob.M, prob.f, prob.g, prob.Q, log=log, **kw)
    sol["log"] = "\n".join(lines)
    sol["geometry"] = prob.geometry
    return sol


def parabolic_solve(mg, assemble, intersect, convex_Euclidian_power, prepare_amg, default_slack_space,
                    p=1.0, h=0.2, t0=0.0, t1=1.0, ts=None, f1=None, g=None, state_variables=None,
                    D=None, Q=None, **rest):
    """Parabolic.jl:126-173.  The host-side constructors are passed in so the oracle does not
    hard-wire the product package's module name."""
    geom = mg.geometry
    dim = geom.dim
    if ts is None:
        ts = np.arange(t0, t1 + 0.5 * h, h)
    x = geom.xflat()
    n = x.shape[0]
    if f1 is None:
        f1 = lambda t, xx: 0.5
    if g is None:
        g = lambda t, xx: np.array([float(np.dot(xx, xx)) if dim > 1 else float(xx[0]), 0.0, 0.0])
    if state_variables is None:
        spc = default_slack_space(geom)
        state_variables = [("u", "dirichlet"), ("s1", spc), ("s2", spc)]
    if D is None:
        D = [("u", "id")] + [("u", nm) for nm in ("dx", "dy", "dz")[:dim]] + [("s1", "id"), ("s2", "id")]
    if Q is None:
        idx1 = (0, dim + 1)
        idx2 = tuple(range(1, dim + 1)) + (dim + 2,)
        Q = intersect(mg, convex_Euclidian_power(mg, idx=idx1, p_grid=np.full(n, 2.0)),
                      convex_Euclidian_power(mg, idx=idx2, p_grid=np.full(n, float(p))))
    f1_grid = np.array([[f1(ts[j], x[i]) for j in range(len(ts))] for i in range(n)])
    U = [np.array([g(ts[k], x[i]) for i in range(n)], dtype=float) for k in range(len(ts))]
    M = prepare_amg(mg, state_variables, D)
    for k in range(len(ts) - 1):
        j = k + 1
        dt = ts[j] - ts[j - 1]
        fg = np.zeros((n, len(D)))
        fg[:, 0] = dt * f1_grid[:, j] - U[k][:, 0]
        fg[:, dim + 1] = 0.5
        fg[:, dim + 2] = dt / p
        prob = assemble(mg, M=M, g_grid=U[k + 1], f_grid=fg, Q=Q, state_variables=state_variables, D=D)
        sol = mgb_solve(prob, **rest)
        U[k + 1] = sol["z"]
    return dict(geometry=geom, ts=np.asarray(ts), u=U)
