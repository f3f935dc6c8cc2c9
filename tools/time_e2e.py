"""Where does the end-to-end time of mgb_solve(prob) go?  python tools/time_e2e.py [L]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mgbx
from mgbx import native, solver, geometry as G, hierarchy as H, problem as P
L = int(sys.argv[1]) if len(sys.argv) > 1 else 10
prob = P.assemble(H.amg(G.subdivide(G.fem2d_P1(), L)), p=1.5)
import torch; torch.cuda.init(); torch.zeros(1, device="cuda")     # CUDA context exists, as in bench.py
for rep in range(2):
    t0 = time.time()
    sol = solver.mgb_solve(prob, config=dict(verbose=1))
    t1 = time.time()
    st = sol["stats"]
    print("rep %d: total %.3fs create %.3fs  device stages f01 %.0f f2 %.0f solve %.0f ms; newton %d; t-steps %d"
          % (rep, t1 - t0, st["create_s"], st["ms_f01"], st["ms_f2"], st["ms_solve"], int(sol["SOL_main"]["its"].sum()),
             sol["SOL_main"]["its"].shape[1]), flush=True)
