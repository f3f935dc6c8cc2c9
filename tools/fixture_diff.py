"""GPU solve of a size-fixture case next to the oracle fixture: t-schedule, Newton counts per barrier step and level, z / objective errors.
    python tools/fixture_diff.py fem3d_k1_c24_t0.1 [cfg key=value ...]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mgbx
from mgbx import solver, geometry as G, hierarchy as H, problem as P
name = sys.argv[1]
cfg = {}
for a in sys.argv[2:]:
    k, v = a.split("=")
    cfg[k] = float(v) if ("." in v or "e" in v.lower()) else int(v)
fx = np.load(os.path.join(ROOT, "tests", "golden", "size_%s.npz" % name))
meta = json.loads(str(fx["meta"]))
kw = dict(meta.get("solve_kwargs", {}))
if name.startswith("fem3d_k1_c"):
    prob = P.assemble(H.amg(G.structured_box(3, int(name.split("_c")[1].split("_")[0]), k=1)), p=1.0)
elif name.startswith("fem3d_k"):
    k = int(name[7]); L = int(name.split("_L")[1].split("_")[0])
    prob = P.assemble(H.amg(G.subdivide(G.fem3d(k=k), L)), p=1.0)
elif name.startswith("fem2d_P1_L"):
    prob = P.assemble(H.amg(G.subdivide(G.fem2d_P1(), int(name.split("_L")[1].split("_")[0]))), p=1.5)
else:
    raise SystemExit("unknown case")
sol = solver.mgb_solve(prob, config=cfg, **kw)
S = sol["SOL_main"]
st = meta["stride"]
rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
print(json.dumps(dict(case=name, cfg=cfg, z_rel=rel(sol["z"][::st], fx["z"]), obj_rel=float(abs(S["c_dot_Dz"][-1] - fx["c_dot_Dz"][-1]) / abs(fx["c_dot_Dz"][-1])),
                      ts_gpu=[float(t) for t in S["ts"]], ts_oracle=[float(t) for t in fx["ts"]],
                      its_gpu=S["its"].sum(axis=0).tolist(), its_oracle=fx["its"].sum(axis=0).tolist(),
                      fine_gpu=S["its"][-1].tolist(), fine_oracle=fx["its"][-1].tolist(), max_gpu=S["its"].max(axis=0).tolist(), max_oracle=fx["its"].max(axis=0).tolist(),
                      fin_gpu=int(S["its_finalize"]), fin_oracle=int(fx["its_finalize"]), stats={k: v for k, v in sol["stats"].items() if k != "device_memory_report"})))
