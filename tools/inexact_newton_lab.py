"""How accurate must the Newton-system solve be?  CPU experiment with the oracle (test/tuning infrastructure only).

The oracle's direct solve is replaced by `x = H^-1 (g + e)` with a random residual e, |e| = rtol |g| -- the error model of a
PCG solve stopped at relative residual rtol -- and the Newton counts / t-schedule of mgb_solve are compared with the exact run.

    python tools/inexact_newton_lab.py p1l6            # fem2d_P1 level 6, p = 1.5
    python tools/inexact_newton_lab.py q1c8            # fem3d k = 1, 8^3, p = 1, t0 = 0.01
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np

import mgbx  # noqa: F401
import mgb_oracle as O
from mgbx import geometry as G, hierarchy as H, problem as P


def run(prob, kw, rtol, seed=0):
    rng = np.random.default_rng(seed)
    exact = O.solve_sym

    def noisy(Hm, g):
        if rtol == 0.0:
            return exact(Hm, g)
        e = rng.normal(size=g.shape)
        e *= rtol * np.linalg.norm(g) / np.linalg.norm(e)
        return exact(Hm, g + e)

    O.solve_sym = noisy
    try:
        sol = O.mgb_solve(prob, **kw)
    finally:
        O.solve_sym = exact
    return sol


def main():
    case = sys.argv[1] if len(sys.argv) > 1 else "p1l6"
    if case.startswith("q1c"):
        prob = P.assemble(H.amg(G.structured_box(3, int(case[3:]), k=1)), p=1.0)
        kw = dict(t=0.01)
    else:
        prob = P.assemble(H.amg(G.subdivide(G.fem2d_P1(), int(case[3:]))), p=1.5)
        kw = {}
    ref = run(prob, kw, 0.0)
    its0 = ref["SOL_main"]["its"]
    print("%s exact: %d barrier steps, Newton per level %s, total %d" % (case, its0.shape[1], its0.sum(axis=1).tolist(), its0.sum()))
    for rtol in (1e-12, 1e-9, 1e-7, 1e-6, 1e-5, 1e-4, 1e-3):
        try:
            sol = run(prob, kw, rtol)
        except Exception as e:   # noqa: BLE001
            print("rtol %.0e: FAILED %r" % (rtol, e))
            continue
        its = sol["SOL_main"]["its"]
        same = its.shape == its0.shape
        dmax = int(np.max(np.abs(its - its0))) if same else -1
        dz = np.linalg.norm(sol["z"] - ref["z"]) / np.linalg.norm(ref["z"])
        print("rtol %.0e: barrier steps %d, Newton total %d (exact %d), max |diff| per (level, step) %s, z rel diff %.1e"
              % (rtol, its.shape[1], its.sum(), its0.sum(), dmax if same else "n/a (schedule differs)", dz))


if __name__ == "__main__":
    main()
