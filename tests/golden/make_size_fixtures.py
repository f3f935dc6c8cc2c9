#!/usr/bin/env python
"""Oracle fixtures at (or towards) the BASELINE.json sizes -- generated ONCE on CPU, committed as small .npz files.

    python tests/golden/make_size_fixtures.py [case ...]        # no argument: every case, sequentially
    python tests/golden/make_size_fixtures.py --list

Each case runs the CPU oracle (oracle/mgb_oracle.py: the line-cited restatement of the reference, pinned by the
reference's own golden vectors in tests/test_oracle_golden.py) through the FULL default solve -- t-ramp to
t >= 1/sqrt(eps), finalize pass included -- on a problem built by the same host calls the GPU tests make, and stores what
the north-star's gates need:

    z           final state (n x nu), possibly strided (every `stride`-th node) to keep the file small
    z_unfin     the state before the finalize pass (SOL.z_unfinalized, src/mgb.jl:76-80), same stride
    znorm, znorm_unfin   full-vector L2 norms
    its         Newton iterations per level and barrier step (SOL_main.its); its_finalize = the part of its[L, end] spent in
                the finalize pass (bookkeeping added to the oracle; the reference adds it into its[L])
    ts, kappas, c_dot_Dz  the t-schedule and the objective history (final objective = c_dot_Dz[-1])
    meta        JSON: case description, sizes, oracle wall time, outcome ("ok" or the MGBConvergenceFailure code)

The GPU tests (tests/test_gpu_parity.py::test_size_fixture_*) assert z 1e-6 rel. L2, objective 1e-8 rel., Newton counts
+-1 per barrier step against these files; bench.py reads the C2 fixtures for its `parity` field.  Failing cases are
fixtures too: fem3d 32^3 at the default t = 0.1 records the reference algorithm's own "Initial centering failed".
"""
import json
import os
import sys
import time

os.environ.setdefault("OMP_NUM_THREADS", "1")
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
os.environ.setdefault("MKL_NUM_THREADS", "1")

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import numpy as np

import mgbx  # noqa: F401
from mgbx import geometry as G, hierarchy as H, problem as P
import mgb_oracle as O


def _solve_case(build, stride=1, **kw):
    t0 = time.time()
    prob = build()
    t_build = time.time() - t0
    n = prob.geometry.n
    t0 = time.time()
    out = dict(n=n)
    try:
        sol = O.mgb_solve(prob, **kw)
        S = sol["SOL_main"]
        nu = sol["z"].shape[1]
        zu = S["z_unfinalized"].reshape(nu, n).T
        out.update(z=sol["z"][::stride].copy(), z_unfin=zu[::stride].copy(), znorm=float(np.linalg.norm(sol["z"])),
                   znorm_unfin=float(np.linalg.norm(zu)), its=S["its"], ts=S["ts"], kappas=S["kappas"], c_dot_Dz=S["c_dot_Dz"],
                   its_finalize=S["its_finalize"])
        if sol["SOL_feasibility"] is not None:
            F = sol["SOL_feasibility"]
            out.update(feas_its=F["its"], feas_ts=F["ts"])
        outcome = "ok"
    except O.MGBConvergenceFailure as e:
        outcome = e.code
        out["message"] = e.message
    meta = dict(outcome=outcome, n=int(n), stride=stride, oracle_wall_s=round(time.time() - t0, 1), host_build_s=round(t_build, 1),
                levels=[int(R.shape[1]) for R in prob.M[0].R_fine], solve_kwargs={k: v for k, v in kw.items()})
    return out, meta


def case_fem2d_P1(L, p, stride):
    return lambda: _solve_case(lambda: P.assemble(H.amg(G.subdivide(G.fem2d_P1(), L)), p=p), stride=stride)


def case_fem3d(c, t, p=1.0, maxit=10000, stride=1):
    kw = dict(t=t)
    if maxit != 10000:
        kw["maxit"] = maxit
    return lambda: _solve_case(lambda: P.assemble(H.amg(G.structured_box(3, c, k=1)), p=p), stride=stride, **kw)


def case_fem3d_k(k, L, t=0.1, stride=1):
    """fem3d(k=k) refined L times (src/TensorFEM.jl: the reference's default fem3d is k = 3): Q_k hexahedra, (k+1)^3 nodes each."""
    return lambda: _solve_case(lambda: P.assemble(H.amg(G.subdivide(G.fem3d(k=k), L)), p=1.0), stride=stride, t=t)


def case_spectral2d(n1):
    return lambda: _solve_case(lambda: P.assemble(H.amg(G.spectral2d(n=n1)), p=1.0))


def case_pure_p2(L, stride=1):
    """test/test_pure_p2.jl:53-63 at size: bubble-free P2, slack in :broken_P1 (the un-condensable family)."""
    def build():
        mg = H.amg(G.subdivide(G.fem2d_P2(bubble=False), L))
        return P.assemble(mg, p=1.0)
    return lambda: _solve_case(build, stride=stride)


def case_parabolic(L, h=0.2, p=1.0, stride=1):
    def run():
        t0 = time.time()
        mg = H.amg(G.subdivide(G.fem2d_P2(), L))
        t_build = time.time() - t0
        t0 = time.time()
        sol = O.parabolic_solve(mg, P.assemble, P.intersect, P.convex_Euclidian_power, H.prepare_amg, P.default_slack_space, p=p, h=h)
        U = np.stack(sol["u"], axis=0)[:, ::stride]           # (steps+1, n/stride, 3)
        out = dict(n=mg.geometry.n, u=U, ts_time=sol["ts"], unorm=np.array([np.linalg.norm(u) for u in sol["u"]]))
        meta = dict(outcome="ok", n=int(mg.geometry.n), stride=stride, oracle_wall_s=round(time.time() - t0, 1), host_build_s=round(t_build, 1),
                    solve_kwargs=dict(h=h, p=p))
        return out, meta
    return run


CASES = {
    # config C2 family (BASELINE.json configs[1]): fem2d_P1, p = 1.5
    "fem2d_P1_L7_p1.5": ("mgb_solve(assemble(amg(subdivide(fem2d_P1(),7)); p=1.5))", case_fem2d_P1(7, 1.5, 1)),
    "fem2d_P1_L8_p1.5": ("mgb_solve(assemble(amg(subdivide(fem2d_P1(),8)); p=1.5))", case_fem2d_P1(8, 1.5, 1)),
    "fem2d_P1_L9_p1.5": ("mgb_solve(assemble(amg(subdivide(fem2d_P1(),9)); p=1.5))", case_fem2d_P1(9, 1.5, 4)),
    "fem2d_P1_L10_p1.5": ("mgb_solve(assemble(amg(subdivide(fem2d_P1(),10)); p=1.5))  [= bench.py's workload, config C2]", case_fem2d_P1(10, 1.5, 16)),
    # config C4 family: fem3d k=1, p = 1 on c^3 hexahedra, default t = 0.1 and the t = 0.01 start
    "fem3d_k1_c16_t0.1": ("mgb_solve(assemble(amg(fem3d(k=1, K=16^3 box)); p=1.0); t=0.1)", case_fem3d(16, 0.1)),
    "fem3d_k1_c16_t0.01": ("mgb_solve(assemble(amg(fem3d(k=1, K=16^3 box)); p=1.0); t=0.01)", case_fem3d(16, 0.01)),
    "fem3d_k1_c24_t0.1": ("mgb_solve(assemble(amg(fem3d(k=1, K=24^3 box)); p=1.0); t=0.1)", case_fem3d(24, 0.1)),
    "fem3d_k1_c24_t0.01": ("mgb_solve(assemble(amg(fem3d(k=1, K=24^3 box)); p=1.0); t=0.01)", case_fem3d(24, 0.01)),
    "fem3d_k1_c32_t0.01": ("mgb_solve(assemble(amg(fem3d(k=1, K=32^3 box)); p=1.0); t=0.01)", case_fem3d(32, 0.01, stride=4)),
    # config C4 variants (ii)/(iii): Q_2 and Q_3 hexahedra
    "fem3d_k2_L3_t0.1": ("mgb_solve(assemble(amg(subdivide(fem3d(k=2),3)); p=1.0); t=0.1)", case_fem3d_k(2, 3)),
    "fem3d_k2_L4_t0.1": ("mgb_solve(assemble(amg(subdivide(fem3d(k=2),4)); p=1.0); t=0.1)", case_fem3d_k(2, 4)),
    "fem3d_k2_L5_t0.1": ("mgb_solve(assemble(amg(subdivide(fem3d(k=2),5)); p=1.0); t=0.1)", case_fem3d_k(2, 5)),
    "fem3d_k3_L2_t0.1": ("mgb_solve(assemble(amg(subdivide(fem3d(k=3),2)); p=1.0); t=0.1)", case_fem3d_k(3, 2)),
    "fem3d_k3_L3_t0.1": ("mgb_solve(assemble(amg(subdivide(fem3d(k=3),3)); p=1.0); t=0.1)", case_fem3d_k(3, 3)),
    "fem3d_k3_L4_t0.1": ("mgb_solve(assemble(amg(subdivide(fem3d(k=3),4)); p=1.0); t=0.1)", case_fem3d_k(3, 4)),
    # config C3 family
    "spectral2d_n32_p1": ("mgb_solve(assemble(amg(spectral2d(n=32)); p=1.0))", case_spectral2d(32)),
    # config C5 family
    "parabolic_fem2d_P2_L4": ("parabolic_solve(amg(subdivide(fem2d_P2(),4)); p=1, h=0.2)", case_parabolic(4)),
    "parabolic_fem2d_P2_L5": ("parabolic_solve(amg(subdivide(fem2d_P2(),5)); p=1, h=0.2)", case_parabolic(5)),
    "parabolic_fem2d_P2_L6": ("parabolic_solve(amg(subdivide(fem2d_P2(),6)); p=1, h=0.2)", case_parabolic(6)),
    "parabolic_fem2d_P2_L7": ("parabolic_solve(amg(subdivide(fem2d_P2(),7)); p=1, h=0.2)", case_parabolic(7, stride=4)),
    "parabolic_fem2d_P2_L8": ("parabolic_solve(amg(subdivide(fem2d_P2(),8)); p=1, h=0.2)", case_parabolic(8, stride=16)),
    # un-condensable family (test/test_pure_p2.jl) at size
    "pure_p2_L4_p1": ("mgb_solve(assemble(amg(subdivide(fem2d_P2(bubble=false),4)); p=1.0))", case_pure_p2(4)),
    "pure_p2_L6_p1": ("mgb_solve(assemble(amg(subdivide(fem2d_P2(bubble=false),6)); p=1.0))", case_pure_p2(6)),
    "pure_p2_L7_p1": ("mgb_solve(assemble(amg(subdivide(fem2d_P2(bubble=false),7)); p=1.0))", case_pure_p2(7, stride=2)),
    "pure_p2_L8_p1": ("mgb_solve(assemble(amg(subdivide(fem2d_P2(bubble=false),8)); p=1.0))", case_pure_p2(8, stride=8)),
}


def main():
    args = sys.argv[1:]
    if args and args[0] == "--list":
        for k, (call, _) in CASES.items():
            print(k, "--", call)
        return
    names = args or list(CASES)
    for name in names:
        call, fn = CASES[name]
        print("[fixture] %s: %s" % (name, call), flush=True)
        out, meta = fn()
        meta.update(case=name, call=call, generator="tests/golden/make_size_fixtures.py",
                    oracle="oracle/mgb_oracle.py (NumPy/SciPy SuperLU restatement, 1 thread)")
        path = os.path.join(HERE, "size_%s.npz" % name)
        np.savez_compressed(path, meta=json.dumps(meta), **{k: np.asarray(v) for k, v in out.items()})
        print("[fixture] %s -> %s (%.1f kB) %s" % (name, os.path.basename(path), os.path.getsize(path) / 1e3, json.dumps(meta)), flush=True)


if __name__ == "__main__":
    main()
