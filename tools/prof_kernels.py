"""Small driver for ncu: a few fine-level evaluations / assemblies / solves of the bench problem.
python tools/prof_kernels.py [L] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mgbx
from mgbx import native, solver, geometry as G, hierarchy as H, problem as P
L = int(sys.argv[1]) if len(sys.argv) > 1 else 10
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
prob = P.assemble(H.amg(G.subdivide(G.fem2d_P1(), L)), p=1.5)
h = native.Handle(prob, barrier_weights=solver.barrier_weights(prob.M[0].w))
M = prob.M[0]; J = len(M.R_fine) - 1; m = M.R_fine[J].shape[1]
rng = np.random.default_rng(0)
s = 1e-4 * rng.normal(size=m); g = rng.normal(size=m)
for k in range(reps):
    y = h.barrier_eval(0, J, 1.0, s, 0)
    x, it = h.solve_newton_system(0, J, 1.0, s, g)
print("f0 %.12g, pcg iterations %d" % (y, it), h.solver_info())
h.close()
