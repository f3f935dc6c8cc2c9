"""CPU check of the per-node barrier arithmetic that the CUDA kernels execute (csrc/node_barrier.cuh,
__host__ __device__) against the oracle's vectorised restatement -- no GPU needed.  The shim under
tests/hostcheck is test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import mgb_oracle as O
from mgbx import native, problem as P

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "hostcheck", "_node_check.so")


@pytest.fixture(scope="module")
def shim():
    src = os.path.join(HERE, "hostcheck", "node_check.cpp")
    hdr = os.path.join(HERE, "..", "multigridbarrier.jl_b200", "csrc", "node_barrier.cuh")
    if (not os.path.exists(SO)) or os.path.getmtime(SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", src, "-o", SO])
    lib = C.CDLL(SO)
    lib.hostcheck_node_eval.argtypes = [C.POINTER(native.Convex), C.c_int64, C.c_int, C.c_int, C.c_int, C.c_double,
                                        C.c_double, native.c_f64p, C.c_int, native.c_f64p, native.c_f64p,
                                        native.c_f64p, native.c_f64p]
    return lib


def run(shim, Q, Y, feas=None):
    n, ny = Y.shape
    keep = native._Keep()
    Qc = native._pack_convex(keep, Q, n)
    Yc = np.ascontiguousarray(Y.T)
    F0 = np.empty(n)
    F1 = np.empty((ny, n))
    F2 = np.empty((ny * ny, n))
    sl = np.empty(n)
    fe, NC, fb, fR = (0, ny + 1, 0.0, 0.0) if feas is None else (1, feas.NC, feas.b, feas.R)
    shim.hostcheck_node_eval(C.byref(Qc), n, ny, fe, NC, fb, fR, native._ptr(Yc), 2, native._ptr(F0),
                             native._ptr(F1), native._ptr(F2), native._ptr(sl))
    return F0, F1.T, F2.T.reshape(n, ny, ny), sl


def ep(n, idx, p, rng, general=False):
    nz = len(idx)
    A = np.tile(np.eye(nz).reshape(-1), (n, 1))
    b = np.zeros((n, nz))
    if general:
        A = A + 0.1 * rng.normal(size=A.shape)
        b = 0.05 * rng.normal(size=b.shape)
    pg = np.full(n, float(p)) if np.isscalar(p) else p
    mu = np.where((pg == 2) | (pg == 1), 0.0, np.where(pg < 2, 1.0, 2.0))
    return P.Convex([P.Piece(P.KIND_EP, tuple(idx), nz, nz, A, b, pg, mu)])


def close(a, b, tol=1e-12):
    a, b = np.asarray(a), np.asarray(b)
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin)
    scale = 1.0 + np.abs(b[fin])
    assert np.all(np.abs(a[fin] - b[fin]) <= tol * scale), np.max(np.abs(a[fin] - b[fin]) / scale)


@pytest.mark.parametrize("p", [1.0, 1.5, 2.0, 3.0])
@pytest.mark.parametrize("general", [False, True])
def test_ep_interior_and_exterior(shim, p, general):
    rng = np.random.default_rng(1)
    n = 64
    Y = rng.normal(size=(n, 4))
    Y[:, 3] = np.abs(Y[:, 3]) * 3 + 0.1          # mostly interior
    Y[::7, 3] = -0.5                             # s <= 0: outside (Log -> -Inf, _safe_pow -> 0)
    Y[3::11, 3] = 1e-3                           # s^alpha - |q|^2 < 0
    Q = ep(n, (1, 2, 3), p, rng, general)
    F0, F1, F2, sl = run(shim, Q, Y)
    o0, o1, o2 = O.convex_eval(Q, Y, 2)
    close(F0, o0)
    ok = np.isfinite(o0)
    close(F1[ok], o1[ok])
    close(F2[ok], o2[ok])
    close(sl, O.convex_slack(Q, Y))


def test_linear_and_piecewise(shim):
    rng = np.random.default_rng(2)
    n, nD = 50, 5
    Y = rng.normal(size=(n, nD))
    A = rng.normal(size=(n, 3, 2))
    lin = P.Convex([P.Piece(P.KIND_LINEAR, (0, 4), 2, 3, A.transpose(0, 2, 1).reshape(n, -1), np.full((n, 3), 4.0))])
    e1 = ep(n, (1, 2, 3), 1.5, rng, True)
    sel = (rng.random((n, 2)) > 0.3).astype(float)
    Q = P.Convex(lin.pieces + e1.pieces, sel)
    Y[:, 3] = np.abs(Y[:, 3]) * 3 + 2
    F0, F1, F2, sl = run(shim, Q, Y)
    o0, o1, o2 = O.convex_eval(Q, Y, 2)
    close(F0, o0)
    ok = np.isfinite(o0)
    close(F1[ok], o1[ok])
    close(F2[ok], o2[ok])
    close(sl, O.convex_slack(Q, Y))


@pytest.mark.parametrize("kind", ["ep", "linear", "both"])
def test_feasibility_wrapper(shim, kind):
    rng = np.random.default_rng(3)
    n, nD, nu = 40, 4, 2
    NF = nD + 1 + nu
    Y = rng.normal(size=(n, NF))
    Y[:, nD] = 6.0 + rng.random(n)               # slack large enough
    e1 = ep(n, (1, 2, 3), 1.0, rng, True)
    A = rng.normal(size=(n, 2, 1))
    lin = P.Convex([P.Piece(P.KIND_LINEAR, (0,), 1, 2, A.transpose(0, 2, 1).reshape(n, -1), np.full((n, 2), 1.0))])
    pieces = {"ep": e1.pieces, "linear": lin.pieces, "both": e1.pieces + lin.pieces}[kind]
    Q = P.Convex(pieces)
    FQ = O.FeasibilityConvex(Q, 20.0, 10.0, nD + 1)
    F0, F1, F2, _ = run(shim, Q, Y, feas=FQ)
    o0, o1, o2 = O.feasibility_eval(FQ, Y, 2)
    assert np.all(np.isfinite(o0))
    close(F0, o0)
    close(F1, o1)
    close(F2, o2)


def test_analytic_schur_complement_of_the_cone_slack_keeps_its_digits(shim):
    """Node-local condensation of the slack of ONE Euclidean-power cone late in the t-ramp (rho = s^alpha - |q|^2 ~ 1e-8): the
    closed form (2/rho) I + (4/rho^2)(B/(A+B)) q q' that piece_eval returns for a `schur` piece against H_qq - H_qs H_sq / H_ss in
    50-digit arithmetic -- and the same difference formed in double precision, which has no correct digit along q."""
    import mpmath as mp
    mp.mp.dps = 50
    shim.hostcheck_node_eval_schur.argtypes = [C.POINTER(native.Convex), C.c_int64, C.c_int, native.c_f64p, native.c_f64p, native.c_f64p,
                                               native.c_f64p, C.c_uint, C.c_int]
    rng = np.random.default_rng(3)
    n = 6
    for p, al in ((1.0, 2.0), (1.5, 4.0 / 3.0)):
        mu = 0.0 if p == 1.0 else 1.0
        s = 0.5 + rng.random(n)
        d = rng.normal(size=(n, 2))
        d /= np.linalg.norm(d, axis=1)[:, None]
        rho = np.array([1e-3, 1e-5, 1e-6, 1e-7, 1e-8, 3e-9])
        qn = np.sqrt(s ** al - rho)
        Y = np.column_stack([d * qn[:, None], s])                         # inputs (q1, q2, s)
        Q = ep(n, (0, 1, 2), np.full(n, p), rng)
        keep = native._Keep()
        Qc = native._pack_convex(keep, Q, n)
        Yc = np.ascontiguousarray(Y.T)
        out = {}
        for mask in (0, 1):
            F0, F1, F2 = np.empty(n), np.empty((3, n)), np.empty((9, n))
            shim.hostcheck_node_eval_schur(C.byref(Qc), n, 3, native._ptr(Yc), native._ptr(F0), native._ptr(F1), native._ptr(F2), mask, 1)
            out[mask] = F2.T.reshape(n, 3, 3)
        for i in range(n):
            q = [mp.mpf(float(Y[i, 0])), mp.mpf(float(Y[i, 1]))]
            si = mp.mpf(float(Y[i, 2]))
            a = mp.mpf(2) / mp.mpf(p)
            r = si ** a - q[0] ** 2 - q[1] ** 2
            hss = -a * (a - 1) * si ** (a - 2) / r + a * a * si ** (2 * a - 2) / r ** 2 + mp.mpf(mu) / si ** 2
            hqs = [-2 * a * si ** (a - 1) * qq / r ** 2 for qq in q]
            S = [[4 * q[x] * q[y] / r ** 2 + (2 / r if x == y else 0) - hqs[x] * hqs[y] / hss for y in range(2)] for x in range(2)]
            scale = float(2 / r)
            H0, H1 = out[0][i], out[1][i]
            # coupling column and slack corner are the same in both modes (needed for the back-substitution)
            assert np.array_equal(H0[2, :], H1[2, :]) and np.array_equal(H0[:, 2], H1[:, 2])
            naive = H0[:2, :2] - np.outer(H0[:2, 2], H0[2, :2]) / H0[2, 2]
            err_an = max(abs(float(S[x][y]) - H1[x, y]) for x in range(2) for y in range(2)) / scale
            err_nv = max(abs(float(S[x][y]) - naive[x, y]) for x in range(2) for y in range(2)) / scale
            # component along q (the small eigenvalue of the condensed block): true value, analytic, naive
            u = np.array([float(q[0]), float(q[1])])
            u /= np.linalg.norm(u)
            lam = float(sum(S[x][y] * mp.mpf(u[x]) * mp.mpf(u[y]) for x in range(2) for y in range(2)))
            lam_an, lam_nv = float(u @ H1[:2, :2] @ u), float(u @ naive @ u)
            # rho itself is a difference of the inputs (as in the reference): its rounding, eps s^alpha / rho, is the floor for
            # both forms.  Across q both forms reach it; ALONG q the closed form stays at that floor RELATIVE TO THE O(1) ENTRY
            # while the subtraction carries an absolute error ~ eps (2/rho)^2 -- no digit left once rho <= 1e-8.
            floor = 32 * np.finfo(float).eps * float(si ** a / r)
            assert err_an < floor + 1e-14, (p, float(r), err_an, floor)
            assert abs(lam_an - lam) <= (floor + 1e-13) * abs(lam), (p, float(r), lam, lam_an)
            print("p=%g rho=%.0e  lambda_along_q: exact %.6g analytic %.6g (rel err %.1e)  subtraction %.6g (rel err %.1e)"
                  % (p, float(r), lam, lam_an, abs(lam_an - lam) / abs(lam), lam_nv, abs(lam_nv - lam) / abs(lam)))
            if p == 1.0 and float(r) <= 1e-7:     # alpha = 2 (p = 1: total variation, fem3d p = 1): the entry along q is O(1)
                assert abs(lam_nv - lam) > 0.05 * abs(lam) and abs(lam_an - lam) < 1e-6 * abs(lam), (float(r), lam, lam_an, lam_nv, err_nv)
