import sys, time
sys.path.insert(0, ".")
import numpy as np, mgbx, torch
from mgbx import solver, geometry as G, hierarchy as H, problem as P
c = int(sys.argv[1]); ts = [float(a) for a in sys.argv[2:] if "=" not in a] or [0.1]
cfg = {}
for a in sys.argv[2:]:
    if "=" in a:
        k, v = a.split("="); cfg[k] = float(v) if ("." in v or "e" in v) else int(v)
t0 = time.time(); prob = P.assemble(H.amg(G.structured_box(3, c, k=1)), p=1.0); print("host build %.1fs n=%d levels %s" % (time.time() - t0, prob.geometry.n, [R.shape[1] for R in prob.M[0].R_fine]), flush=True)
for t in ts:
    t0 = time.time()
    try:
        sol = solver.mgb_solve(prob, t=t, config=cfg)
        st = sol["stats"]
        print("t0=%g: solve %.1fs (create %.1fs) its/level %s t-steps %d; stages f01 %.0f f2 %.0f solve %.0f ms; pcg %d; obj %.10g" % (t, time.time() - t0, st["create_s"], sol["SOL_main"]["its"].sum(axis=1).tolist(), sol["SOL_main"]["its"].shape[1], st["ms_f01"], st["ms_f2"], st["ms_solve"], st["pcg_iters"], sol["SOL_main"]["c_dot_Dz"][-1]), flush=True)
    except Exception as e:
        print("t0=%g FAILED after %.1fs: %r" % (t, time.time() - t0, e), flush=True)
