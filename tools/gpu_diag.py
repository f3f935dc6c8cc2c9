"""Staged GPU diagnostics: every stage of the device path against the CPU oracle, with detailed diffs.
Run on the GPU box:  python tools/gpu_diag.py [case ...]   (writes gpurun_out/diag.log as well)."""
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import scipy.sparse as sp

import mgbx  # noqa
from mgbx import native, solver, geometry as G, hierarchy as H, problem as P
import mgb_oracle as O
from helpers import GEOMS, gold, lower_bound_problem

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "diag.log"), "a")


def say(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    LOG.write(s + "\n")
    LOG.flush()


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    d = np.linalg.norm(a - b)
    return d / max(np.linalg.norm(b), 1e-300)


def stage_hooks(name, prob, which=0, t=0.1, cfgs=({}, {"condense": 0})):
    M = prob.M[which]
    L = len(M.R_fine)
    n = M.geometry.n
    rng = np.random.default_rng(0)
    bw = O.barrier_weights(M.w) if which == 0 else None
    h = native.Handle(prob, barrier_weights=bw)
    try:
        Qo = prob.Q
        if which == 1:
            need, b, zabs = h.phase1_init()
            h.set_feasibility_box(b, 100.0)
            Qo = O.FeasibilityConvex(prob.Q, b, 100.0, prob.M[0].nD + 1)
            c1 = np.zeros((n, M.nD)); c1[:, prob.M[0].nD] = 1.0
        B = O.Barrier(Qo, bw)
        ops = O.operators(M)
        z0 = prob.g.T.reshape(-1).copy()
        if which == 1:
            z0 = h.get_z(1)
        c = t * (prob.f if which == 0 else c1)
        for J in range(L):
            R = M.R_fine[J]
            m = R.shape[1]
            s = 1e-3 * rng.normal(size=m)
            try:
                y_o = B.f0(s, M.w, c, R, ops, z0)
                g_o = B.f1(s, M.w, c, R, ops, z0)
                y_d = h.barrier_eval(which, J, t, s, 0)
                g_d = h.barrier_eval(which, J, t, s, 1)
                say("  [%s] J=%d m=%d f0 dev=%.15g ora=%.15g relerr=%.2e | f1 relerr=%.2e" %
                    (name, J, m, y_d, y_o, abs(y_d - y_o) / max(1, abs(y_o)), rel(g_d, g_o)))
            except Exception as e:
                say("  [%s] J=%d f0/f1 FAILED: %r" % (name, J, e))
                traceback.print_exc()
            try:
                H_o = sp.csr_matrix(B.f2(s, M.w, c, R, ops, z0))
                H_d = h.hessian(which, J, t, s)
                optr, oind = O.hessian_pattern(M, J)
                pat_ok = np.array_equal(H_d.indptr, optr) and np.array_equal(H_d.indices, oind)
                diff = abs(H_d - H_o)
                say("  [%s] J=%d f2 pattern_exact=%s nnz=%d relerr=%.2e" %
                    (name, J, pat_ok, H_d.nnz, diff.sum() / max(abs(H_o).sum(), 1e-300)))
                rhs = g_o
                x_o = O.solve_sym(H_o, rhs)
                for cfg in cfgs:
                    h2 = native.Handle(prob, barrier_weights=bw, **cfg) if cfg else h
                    x_d, it = h2.solve_newton_system(which, J, t, s, rhs)
                    say("  [%s] J=%d solve cfg=%s iters=%d relerr=%.2e resid=%.2e" %
                        (name, J, cfg, it, rel(x_d, x_o), np.linalg.norm(H_o @ x_d - rhs) / np.linalg.norm(rhs)))
                    if cfg:
                        h2.close()
            except Exception as e:
                say("  [%s] J=%d f2/solve FAILED: %r" % (name, J, e))
                traceback.print_exc()
    finally:
        h.close()


def stage_solve(name, prob, goldname=None, **kw):
    t0 = time.time()
    try:
        sol = solver.mgb_solve(prob, **kw)
    except Exception as e:
        say("  [%s] mgb_solve FAILED: %r" % (name, e))
        traceback.print_exc()
        return None
    dt = time.time() - t0
    msg = "  [%s] mgb_solve %.2fs its=%s" % (name, dt, sol["SOL_main"]["its"].sum(axis=0).tolist())
    if goldname:
        msg += " |z-gold|=%.3e" % np.linalg.norm(sol["z"] - gold(goldname))
    st = sol["stats"]
    msg += " launches=%d pcg=%d f01=%d f2=%d ms(f01,f2,solve)=(%.1f,%.1f,%.1f)" % (
        st["gpu_launches"], st["pcg_iters"], st["f01_evals"], st["f2_evals"], st["ms_f01"], st["ms_f2"], st["ms_solve"])
    say(msg)
    return sol


def compare_oracle(name, prob, sol):
    try:
        so = O.mgb_solve(prob)
        zo = so["z"]
        say("  [%s] vs oracle: rel L2 z %.3e | obj dev %.12g ora %.12g | its dev %s ora %s" % (
            name, rel(sol["z"], zo), sol["SOL_main"]["c_dot_Dz"][-1], so["SOL_main"]["c_dot_Dz"][-1],
            sol["SOL_main"]["its"].sum(axis=0).tolist(), so["SOL_main"]["its"].sum(axis=0).tolist()))
    except Exception as e:
        say("  [%s] oracle compare FAILED: %r" % (name, e))


CASES = {
    "fem1d3": ("fem1d_3nodes", 1.0, "fem1d_3nodes_p1"),
    "p2quick": ("fem2d_P2_quickstart", 1.0, "fem2d_P2_quickstart_p1"),
    "fem1d5": ("fem1d_5nodes", 1.5, "fem1d_5nodes_p1.5"),
    "p1L2": ("fem2d_P1_L2", 1.0, "fem2d_P1_L2_p1"),
    "p2L2": ("fem2d_P2_L2", 1.5, "fem2d_P2_L2_p1.5"),
    "q1L2": ("fem3d_k1_L2", 1.0, "fem3d_k1_L2_p1"),
    "spec1d": ("spectral1d_n5", 1.0, "spectral1d_n5_p1"),
    "spec2d": ("spectral2d_n5", 1.0, "spectral2d_n5_p1"),
}


def main(argv):
    say("==== gpu_diag", time.ctime(), "devices", native.lib().mgbx_device_count())
    which = argv or ["fem1d3", "p2quick", "p1L2", "q1L2", "spec1d", "feas", "p1L5", "p1L7"]
    for c in which:
        say("== case", c)
        try:
            if c in CASES:
                gname, p, goldname = CASES[c]
                prob = P.assemble(H.amg(GEOMS[gname]()), p=p)
                stage_hooks(c, prob)
                sol = stage_solve(c, prob, goldname)
            elif c == "feas":
                prob = lower_bound_problem(50.0)
                stage_hooks(c + "/feasAMG", prob, which=1, t=0.1)
                sol = stage_solve(c, prob)
                if sol is not None:
                    say("  [feas] max|z-50| = %.3e, R=100 round: %s" % (np.max(np.abs(sol["z"] - 50.0)),
                                                                    "bounding box R=100.0" in sol["log"]))
            elif c.startswith("p1L"):
                Lr = int(c[3:])
                prob = P.assemble(H.amg(G.subdivide(G.fem2d_P1(), Lr)), p=1.5)
                if Lr <= 5:
                    stage_hooks(c, prob, cfgs=({}, {"condense": 0}, {"dense_direct_max": 64, "coarse_max": 32}))
                for cfg in (({}, {"dense_direct_max": 64, "coarse_max": 32}) if Lr <= 7 else ({},)):
                    sol = stage_solve(c + str(cfg), prob, config=cfg)
                    if sol is not None and Lr <= 7:
                        compare_oracle(c, prob, sol)
        except Exception as e:
            say("  case", c, "FAILED:", repr(e))
            traceback.print_exc()


if __name__ == "__main__":
    main(sys.argv[1:])
