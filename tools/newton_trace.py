"""Per-Newton-iteration trace (level, PCG iterations) of one solve.  python tools/newton_trace.py [L] [key=value ...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mgbx
from mgbx import solver, geometry as G, hierarchy as H, problem as P
case = sys.argv[1] if len(sys.argv) > 1 else "10"
cfg = dict(verbose=2)
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    cfg[k] = float(v) if "." in v or "e" in v else int(v)
pp = cfg.pop("p", None)
if case.startswith("q1L"):
    prob = P.assemble(H.amg(G.subdivide(G.fem3d(k=1), int(case[3:]))), p=float(pp or 1.0))
elif case.startswith("q1c"):
    prob = P.assemble(H.amg(G.structured_box(3, int(case[3:]), k=1)), p=float(pp or 1.0))
else:
    prob = P.assemble(H.amg(G.subdivide(G.fem2d_P1(), int(case))), p=float(pp or 1.5))
t0 = time.time()
sol = solver.mgb_solve(prob, config=cfg)
print("total %.3fs its %s" % (time.time() - t0, sol["SOL_main"]["its"].sum(axis=1).tolist()), sol["stats"])
