"""CPU laboratory for the C5 family (parabolic_solve on fem2d_P2): why does the V-cycle PCG struggle late in the t-ramp?

Runs the CPU oracle's first implicit-Euler step, records the fine-level MAIN-ramp Newton systems (H, g, t), eliminates the two
node-local slacks exactly as the library does, and runs the emulated V-cycle PCG of tools/smoother_lab.py on the reduced systems
(relative residual 1e-7, as shipped) for a few preconditioner variants.  Test/tuning infrastructure: imports oracle/.

    python tools/parabolic_lab.py 5        # fem2d_P2 level 5 (3 584 nodes)
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import scipy.sparse as sp

import mgbx  # noqa: F401
import mgb_oracle as O
from mgbx import geometry as G, hierarchy as H, problem as P
from smoother_lab import VCycle, pcg, transfers


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    mg = H.amg(G.subdivide(G.fem2d_P2(), L))
    seen = []
    f1_orig, f2_orig = O.Barrier.f1, O.Barrier.f2
    state = {"M": None}

    def f2(self, s, w, c, R, ops, z0):
        Hm = f2_orig(self, s, w, c, R, ops, z0)
        M = state["M"]
        if M is not None and R.shape == M.R_fine[-1].shape:
            seen.append((sp.csr_matrix(Hm), f1_orig(self, s, w, c, R, ops, z0), float(np.max(np.abs(c)))))
        return Hm

    prep = H.prepare_amg

    def prepare(mg_, sv, D):
        Ms = prep(mg_, sv, D)
        state["M"] = Ms[0]
        return Ms

    O.Barrier.f2 = f2
    t0 = time.time()
    try:
        O.parabolic_solve(mg, P.assemble, P.intersect, P.convex_Euclidian_power, prepare, P.default_slack_space, p=1.0, h=0.2, ts=np.array([0.0, 0.2]))
    finally:
        O.Barrier.f2 = f2_orig
    M = state["M"]
    offs = M.var_offsets[-1]
    nu_ = offs[1] - offs[0]
    print("fem2d_P2 L%d: n = %d, u unknowns %d, %d fine-level main-AMG systems recorded in %.0f s" % (L, mg.geometry.n, nu_, len(seen), time.time() - t0), flush=True)
    Ts = transfers(M)
    pick = seen[::max(1, len(seen) // 14)]
    variants = [("cheb2 r8 V gershgorin (shipped)", dict(smoother="cheb", nu=2, ratio=8.0)),
                ("cheb4 r30 V gershgorin", dict(smoother="cheb", nu=4, ratio=30.0)),
                ("l1-Jacobi 2 V", dict(smoother="l1", nu=2)),
                ("sym. Gauss-Seidel V (sequential bound)", dict(smoother="sgs", nu=2))]
    print("  %-10s %-10s %-12s %s" % ("cost scale", "cond(D^-1/2 S D^-1/2)", "min eig", " | ".join(v[0] for v in variants)))
    for Hm, g, tc in pick:
        Huu, Hus, Hss = Hm[:nu_, :nu_].tocsr(), Hm[:nu_, nu_:].tocsr(), Hm[nu_:, nu_:].tocsr()
        d = Hss.diagonal()
        offd = abs(Hss - sp.diags(d)).sum()
        S = (Huu - Hus @ sp.diags(1.0 / d) @ Hus.T).tocsr()
        gs = g[:nu_] - Hus @ (g[nu_:] / d)
        dd = S.diagonal()
        Sn = (sp.diags(1 / np.sqrt(dd)) @ S @ sp.diags(1 / np.sqrt(dd))).toarray() if nu_ <= 4000 else None
        if Sn is not None:
            ev = np.linalg.eigvalsh(0.5 * (Sn + Sn.T))
            cond, emin = ev[-1] / max(ev[0], 1e-300), ev[0]
        else:
            cond, emin = float("nan"), float("nan")
        its = []
        for _, opt in variants:
            try:
                its.append(pcg(S, gs, VCycle(S, Ts, **opt), rtol=1e-7, maxit=400))
            except Exception as e:
                its.append(-1)
        print("  t*max|f| %-9.3g cond %-10.3g min eig %-10.3g slack block off-diagonal %.1e | PCG its %s" % (tc, cond, emin, offd, "  ".join("%4d" % i for i in its)), flush=True)


if __name__ == "__main__":
    main()
