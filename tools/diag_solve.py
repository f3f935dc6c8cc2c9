"""Per-Newton-system PCG diagnostics (stderr of the library at verbose=2): status, |r|/|b|, energy share of the last 4 iterations.
    python tools/diag_solve.py q1c32 t=0.01 [cfg key=value ...]     |  python tools/diag_solve.py parabolic6"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, mgbx
from mgbx import solver, geometry as G, hierarchy as H, problem as P
case = sys.argv[1]
cfg, kw = dict(verbose=2), {}
for a in sys.argv[2:]:
    k, v = a.split("=")
    val = float(v) if ("." in v or "e" in v.lower()) else int(v)
    if k == "t":
        kw["t"] = val
    else:
        cfg[k] = val
t0 = time.time()
try:
    if case.startswith("parabolic"):
        sol = solver.parabolic_solve(H.amg(G.subdivide(G.fem2d_P2(), int(case[9:]))), h=0.2, p=1.0, config=cfg)
        print(json.dumps(dict(case=case, ok=True, wall=time.time() - t0, stats=sol["stats"])))
    else:
        if case.startswith("q1c"):
            prob = P.assemble(H.amg(G.structured_box(3, int(case[3:]), k=1)), p=1.0)
        elif case.startswith("purep2_"):
            prob = P.assemble(H.amg(G.subdivide(G.fem2d_P2(bubble=False), int(case[7:]))), p=1.0)
        else:
            prob = P.assemble(H.amg(G.subdivide(G.fem2d_P1(), int(case))), p=1.5)
        sol = solver.mgb_solve(prob, config=cfg, **kw)
        st = sol["stats"]
        print(json.dumps(dict(case=case, ok=True, wall=time.time() - t0, newton=int(sol["SOL_main"]["its"].sum()), steps=int(sol["SOL_main"]["its"].shape[1]),
                              objective=float(sol["SOL_main"]["c_dot_Dz"][-1]), stats=st)))
except Exception as e:
    print(json.dumps(dict(case=case, ok=False, wall=time.time() - t0, error=repr(e)[:300])))
