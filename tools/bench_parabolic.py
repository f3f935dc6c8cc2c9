"""Configuration C5: parabolic_solve(amg(subdivide(fem2d_P2(), L)); p=1, h=0.2) -- total-variation flow, 5 implicit Euler steps,
phase I at every step, 3 state variables, piecewise intersection of two cones.  python tools/bench_parabolic.py [L]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mgbx
from mgbx import solver, geometry as G, hierarchy as H
L = int(sys.argv[1]) if len(sys.argv) > 1 else 6
t0 = time.time(); mg = H.amg(G.subdivide(G.fem2d_P2(), L)); tb = time.time() - t0
t0 = time.time(); sol = solver.parabolic_solve(mg, h=0.2, p=1.0); dt = time.time() - t0
u = np.stack(sol["u"], axis=2)
print(json.dumps({"workload": "parabolic_solve(amg(subdivide(fem2d_P2(),%d)); p=1, h=0.2)" % L, "nodes": mg.geometry.n,
                  "time_steps": len(sol["ts"]) - 1, "host_hierarchy_s": tb, "solve_s": dt, "finite": bool(np.isfinite(u).all()),
                  "u_range": [float(u[:, 0, :].min()), float(u[:, 0, :].max())],
                  "per_step": [{k: (round(v, 1) if isinstance(v, float) else v) for k, v in st.items()} for st in sol["stats"]]}))
