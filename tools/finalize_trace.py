import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mgbx
from mgbx import solver, geometry as G, hierarchy as H, problem as P
prob = P.assemble(H.amg(G.subdivide(G.fem3d(k=1), 3)), p=1.0)
cfg = dict(dense_direct_max=64, coarse_max=32, verbose=2)
for kv in sys.argv[1:]:
    k, v = kv.split("="); cfg[k] = float(v) if ("." in v or "e" in v) else int(v)
sol = solver.mgb_solve(prob, config=cfg)
print(sol["SOL_main"]["its"].sum(axis=0).tolist())
