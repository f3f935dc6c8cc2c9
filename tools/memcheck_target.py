"""Small solves that touch every kernel family, for `compute-sanitizer --tool memcheck python tools/memcheck_target.py`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mgbx
from mgbx import native, solver, geometry as G, hierarchy as H, problem as P
# persistent PCG + specialised element kernels + Galerkin plans (PCG forced on a small mesh)
prob = P.assemble(H.amg(G.subdivide(G.fem2d_P1(), 4)), p=1.5)
sol = solver.mgb_solve(prob, config=dict(dense_direct_max=0, coarse_max=16, tail_max=40))
print("p1 pcg ok", sol["stats"]["pcg_iters"], flush=True)
# generic fused kernels + direct small solves + phase I (parabolic: 3 variables, piecewise set)
sol = solver.parabolic_solve(H.amg(G.fem2d_P2()), h=0.5, p=1.0)
print("parabolic ok", flush=True)
# dense DMMA path + blocked Cholesky
prob = P.assemble(H.amg(G.spectral2d(n=12)), p=1.0)
sol = solver.mgb_solve(prob, tol=1e-3)
print("spectral ok", flush=True)
# 3-D, generic + specialised DIM=3
prob = P.assemble(H.amg(G.subdivide(G.fem3d(k=1), 3)), p=1.0)
sol = solver.mgb_solve(prob, tol=1e-3, config=dict(dense_direct_max=64, coarse_max=32))
print("fem3d ok", sol["stats"]["pcg_iters"], flush=True)
