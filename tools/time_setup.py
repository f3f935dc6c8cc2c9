"""Where does handle creation go?  python tools/time_setup.py [L]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mgbx
from mgbx import native, solver, geometry as G, hierarchy as H, problem as P
L = int(sys.argv[1]) if len(sys.argv) > 1 else 10
t = time.time(); prob = P.assemble(H.amg(G.subdivide(G.fem2d_P1(), L)), p=1.5); print("problem build %.2fs" % (time.time() - t))
t = time.time(); keep = native._Keep(); a0 = native._pack_amg(keep, prob.M[0]); print("pack main amg %.2fs" % (time.time() - t))
t = time.time(); a1 = native._pack_amg(keep, prob.M[1]); print("pack feas amg %.2fs" % (time.time() - t))
t = time.time(); h = native.Handle(prob, verbose=1); print("Handle() total %.2fs" % (time.time() - t))
M = prob.M[0]; J = len(M.R_fine) - 1; m = M.R_fine[J].shape[1]
s = np.zeros(m); g = np.ones(m)
t = time.time(); x, it = h.solve_newton_system(0, J, 1.0, s, g); print("first solve (build_system + graph capture) %.2fs, pcg %d" % (time.time() - t, it))
t = time.time(); x, it = h.solve_newton_system(0, J, 1.0, s, g); print("second solve %.3fs" % (time.time() - t))
h.close()
