"""Per-kernel-class achieved GB/s on a 3-D configuration (fem3d k=1 on c^3 hexahedra, p = 1, initial t = 0.01): the HBM-bound
regime the C2 bench cannot show.  python tools/roofline_3d.py [c]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mgbx
from mgbx import native, solver, geometry as G, hierarchy as H, problem as P
import bench
c = int(sys.argv[1]) if len(sys.argv) > 1 else 64
prob = P.assemble(H.amg(G.structured_box(3, c, k=1)), p=1.0)
M = prob.M[0]
h = native.Handle(prob, barrier_weights=solver.barrier_weights(M.w))
sol = solver.mgb_solve(prob, handle=h, t=0.01)                    # warm-up: plans, module load
h.set_grids(None, prob.g)
h.set_profile(1); h.kernel_stats(reset=True)
t0 = time.time(); sol = solver.mgb_solve(prob, handle=h, t=0.01); dt = time.time() - t0
ks = h.kernel_stats(reset=True); h.set_profile(0)
st = sol["stats"]
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6545.0) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6545.0
ab = bench.algorithmic_bytes(prob, h, {k: v[0] for k, v in ks.items()})
info = h.solver_info()
rows = {}
for cls in ("elem_f01", "elem_f2", "elem_generic_f2", "csr_gather", "spgemm", "spmv", "pcg_persistent"):
    nl, ms = ks.get(cls, (0, 0.0))
    if not nl:
        continue
    if cls == "pcg_persistent":
        bpl = ab["pcg_iteration"] * st["pcg_iters"] / nl
    elif cls == "spmv":
        R = M.R_fine[-1]; bpl = 12 * R.nnz + 20 * R.shape[0]              # the fine-level R / R' products dominate this class
    else:
        bpl = ab.get(cls)
    ach = bpl / (ms * 1e-3 / nl) / 1e9 if bpl else None
    rows[cls] = dict(launches=nl, avg_us=round(1e3 * ms / nl, 1), algorithmic_mb_per_launch=round(bpl / 1e6, 1) if bpl else None,
                     achieved_gbs=round(ach) if ach else None, frac_of_measured_hbm_peak=round(ach / peak, 3) if ach else None)
print(json.dumps(dict(workload="fem3d k=1, %d^3 hexahedra, p=1, t0=0.01" % c, nodes=M.geometry.n, elements=M.geometry.N,
                      fine_unknowns_condensed=info["m"][0] if info["nlev"] else None, nnz_top=info["nnz"][0] if info["nlev"] else None,
                      solve_wall_s=round(dt, 2), newton_steps=int(sol["SOL_main"]["its"].sum()), pcg_iters=st["pcg_iters"],
                      hbm_peak_gbs=peak, kernels=rows)))
h.close()
