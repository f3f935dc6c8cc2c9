"""Spectral (dense) configuration C3: spectral2d(n), p = 1.  Solve time, DMMA GEMM throughput against the measured
cuBLAS DGEMM rate of the box.  python tools/bench_spectral.py [n ...]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import mgbx
from mgbx import native, solver, geometry as G, hierarchy as H, problem as P

def dgemm_peak(n=4096, reps=5):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    torch.matmul(a, b); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 0.0
    for _ in range(reps):
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best

peak = dgemm_peak()
print(json.dumps({"cublas_dgemm_tflops_4096": peak}), flush=True)
cfg = {}
for a in sys.argv[1:]:
    if "=" in a:                       # mgbx_config override, e.g. spectral_kron=0
        k, v = a.split("=")
        cfg[k] = float(v) if ("." in v or "e" in v.lower()) else int(v)
for n1 in [int(a) for a in sys.argv[1:] if "=" not in a] or [16, 24, 32]:
    prob = P.assemble(H.amg(G.spectral2d(n=n1)), p=1.0)
    M = prob.M[0]
    h = native.Handle(prob, barrier_weights=solver.barrier_weights(M.w), **cfg)
    sol = solver.mgb_solve(prob, handle=h)                       # warm-up (plans, module load)
    h.set_grids(None, prob.g)
    h.set_profile(1); h.kernel_stats(reset=True)
    f0 = h.solver_info()["dgemm_flops"]
    torch.cuda.synchronize(); t0 = time.time()
    sol = solver.mgb_solve(prob, handle=h)
    torch.cuda.synchronize(); dt = time.time() - t0
    ks = h.kernel_stats(reset=True); fl = h.solver_info()["dgemm_flops"] - f0
    h.set_profile(0); h.close()
    its = int(sol["SOL_main"]["its"].sum())
    g = ks["dgemm_dmma"]
    print(json.dumps({"workload": "spectral2d(n=%d), p=1" % n1, "config": cfg, "objective": float(sol["SOL_main"]["c_dot_Dz"][-1]), "nodes": M.geometry.n, "fine_unknowns": M.R_fine[-1].shape[1],
                      "newton_steps": its, "solve_s": dt, "dof_newton_steps_per_s": M.geometry.n * its / dt,
                      "dgemm_launches": g[0], "dgemm_ms": g[1], "dgemm_tflops": fl / (g[1] * 1e-3) / 1e12 if g[1] else None,
                      "dgemm_frac_of_cublas": (fl / (g[1] * 1e-3) / 1e12 / peak) if g[1] else None,
                      "kernel_ms": {k: round(v[1], 2) for k, v in ks.items() if v[0]}}), flush=True)
