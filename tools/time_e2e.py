"""Where does the end-to-end time of mgb_solve(prob) go?  python tools/time_e2e.py [L]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mgbx
from mgbx import native, solver, geometry as G, hierarchy as H, problem as P
case = sys.argv[1] if len(sys.argv) > 1 else "10"
t0 = time.time()
if case.startswith("q1c"):
    prob = P.assemble(H.amg(G.structured_box(3, int(case[3:]), k=1)), p=1.0)
else:
    prob = P.assemble(H.amg(G.subdivide(G.fem2d_P1(), int(case))), p=1.5)
print("host problem build %.1fs: n=%d N=%d levels %s" % (time.time() - t0, prob.geometry.n, prob.geometry.N,
                                                        [R.shape[1] for R in prob.M[0].R_fine]), flush=True)
import torch; torch.cuda.init(); torch.zeros(1, device="cuda")     # CUDA context exists, as in bench.py
for rep in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    t0 = time.time()
    sol = solver.mgb_solve(prob, config=dict(verbose=1))
    t1 = time.time()
    st = sol["stats"]
    print("rep %d: total %.3fs create %.3fs  device stages f01 %.0f f2 %.0f solve %.0f ms; newton %d; t-steps %d"
          % (rep, t1 - t0, st["create_s"], st["ms_f01"], st["ms_f2"], st["ms_solve"], int(sol["SOL_main"]["its"].sum()),
             sol["SOL_main"]["its"].shape[1]), "pcg", st["pcg_iters"], "its/level", sol["SOL_main"]["its"].sum(axis=1).tolist(), flush=True)
    print(torch.cuda.max_memory_allocated() / 1e9, "GB torch;", "free/total GB", [x / 1e9 for x in torch.cuda.mem_get_info()], flush=True)
