"""Host logic on the CPU: the Python mirror of mgb_core / mgb_driver / mgb_solve / parabolic_solve
(multigridbarrier.jl_b200/solver.py -- what a Julia shim does above the C ABI) driven through a stand-in handle whose
`step` / `scalars` / phase-I entry points are answered by the CPU oracle instead of libmgbx.  The control flow under test is
the product's; the numbers behind the handle are the oracle's, so the result must reproduce the oracle's own drivers exactly:
same t-ramp, same kappa schedule, same Newton counts, same z, same failure codes."""
import math
import types

import numpy as np
import pytest

import mgb_oracle as O
from helpers import default_problem, lower_bound_problem
from mgbx import geometry as G, hierarchy as H, native, problem as P, solver


class OracleHandle:
    """Implements the slice of native.Handle that solver.py uses."""

    def __init__(self, prob, barrier_weights=None):
        self.prob = prob
        self.M = prob.M
        self.n = len(prob.M[0].w)
        self.nu = [prob.M[0].nu, prob.M[1].nu]
        self.bw = barrier_weights
        self.f = prob.f.copy()
        self.z = [prob.g.T.reshape(-1).copy(), None]
        self.zunfin = [self.z[0].copy(), None]
        self.zinit_feas = None
        self.box = (1.0, 10.0)
        self.steps = 0

    # ---- helpers
    def _setup(self, which):
        M = self.M[which]
        if which == native.MAIN:
            return M, self.prob.Q, self.f, self.bw
        nD, ncomp = self.M[0].nD, self.M[0].nu
        c1 = np.zeros((self.n, nD + 1 + ncomp))
        c1[:, nD] = 1.0
        return M, O.FeasibilityConvex(self.prob.Q, float(self.box[0]), float(self.box[1]), nD + 1), c1, None

    def step_opts(self, **kw):
        o = types.SimpleNamespace(maxit=10000, max_newton=8, initial_step=0, stop_kind=1, stop_lambda_tol=0.25 / math.sqrt(self.n),
                                  stop_theta=0.9, finalize=0, finalize_theta=0.9, line_search=0, ls_beta=0.5, ls_c1=0.1)
        for k, v in kw.items():
            setattr(o, k, v)
        return o

    def step(self, which, t, o):
        M, Q, cost, bw = self._setup(which)
        sc = O.stopping_inexact(o.stop_lambda_tol, o.stop_theta) if o.stop_kind == 1 else O.stopping_exact(o.stop_theta)
        ls = O.linesearch_backtracking(o.ls_beta, o.ls_c1) if o.line_search == 0 else O.linesearch_illinois(o.ls_beta)
        fin = O.stopping_exact(o.finalize_theta) if o.finalize else None
        self.steps += 1
        r = types.SimpleNamespace(its=[0] * 32, f01_evals=0, f2_evals=0, linear_solves=0, pcg_iters=0, ms_f01=0.0, ms_f2=0.0, ms_solve=0.0,
                                  converged=0, solve_failures=0, its_finalize=0, direct_fallbacks=0)
        try:
            SOL = O.mgb_step(Q, M, self.z[which], t * cost, o.maxit, o.max_newton, ls, sc, fin, initial_step=bool(o.initial_step),
                             barrier_weights=bw)
        except (FloatingPointError, RuntimeError, ValueError, ArithmeticError):
            return native.NON_FINITE, r
        for k, v in enumerate(SOL["its"]):
            r.its[k] = int(v)
        if SOL["converged"]:
            self.z[which] = SOL["z"]
            self.zunfin[which] = SOL["z_unfinalized"]
            r.its_finalize = int(SOL["its_finalize"])
            r.converged = 1
            return native.OK, r
        return native.NOT_CONVERGED, r

    def scalars(self, which=native.MAIN):
        M, Q, cost, bw = self._setup(which)
        z = self.z[which]
        nu = M.nu
        out = types.SimpleNamespace(c_dot_Dz=O.c_dot_Dz(M, cost, z), var_max=[0.0] * 12, var_absmax=[0.0] * 12, all_finite=1)
        for k in range(nu):
            seg = z[k * self.n:(k + 1) * self.n]
            out.var_max[k] = float(seg.max())
            out.var_absmax[k] = float(np.abs(seg).max())
        return out

    def phase1_init(self):
        M1 = self.M[0]
        z2 = self.z[0]
        Dz0 = O.operators(M1).apply(z2)
        F0, _, _ = O.convex_eval(self.prob.Q, Dz0, 0)
        zabs = float(np.max(np.abs(z2)))
        if np.all(np.isfinite(F0)):
            return False, 0.0, zabs
        sl = 2.0 * np.maximum(O.convex_slack(self.prob.Q, Dz0), 1.0)
        self.zinit_feas = np.concatenate([z2, sl])
        self.z[1] = self.zinit_feas.copy()
        return True, 2.0 * max(1.0, float(sl.max())), zabs

    def set_feasibility_box(self, b, R):
        self.box = (b, R)

    def reset_feasibility_state(self):
        self.z[1] = self.zinit_feas.copy()

    def handoff(self):
        self.z[0] = self.z[1][:self.z[0].size].copy()

    def matched_t(self, t_default):
        return O.matched_t(self.prob.Q, self.M[0], self.z[0], self.f, t_default, barrier_weights=self.bw), float("nan")

    def get_z(self, which=native.MAIN):
        return self.z[which].copy()

    def memory_report(self):
        return "", 0

    def get_z_unfinalized(self, which=native.MAIN):
        return self.zunfin[which].copy()

    def set_grids(self, f_grid=None, g_grid=None):
        if f_grid is not None:
            self.f = np.asarray(f_grid, float).copy()
        if g_grid is not None:
            self.z[0] = np.asarray(g_grid, float).T.reshape(-1).copy()

    def launch_count(self):
        return 0

    def close(self):
        pass


def _same_history(sd, so):
    a, b = sd["SOL_main"], so["SOL_main"]
    assert a["its"].shape == b["its"].shape and np.array_equal(a["its"], b["its"])
    assert np.allclose(a["ts"], b["ts"], rtol=0, atol=0) and np.allclose(a["kappas"], b["kappas"], rtol=0, atol=0)
    assert np.allclose(a["c_dot_Dz"], b["c_dot_Dz"], rtol=1e-14, atol=0)
    assert np.allclose(sd["z"], so["z"], rtol=0, atol=1e-13)


@pytest.mark.parametrize("geom,p", [("fem1d_5nodes", 1.0), ("fem2d_P2_quickstart", 1.0), ("fem2d_P1_L2", 1.5), ("spectral1d_n5", 1.0)])
def test_mgb_core_and_driver_reproduce_the_oracle_drivers(geom, p):
    prob = default_problem(geom, p)
    bw = solver.barrier_weights(prob.M[0].w)
    sd = solver.mgb_solve(prob, handle=OracleHandle(prob, bw))
    so = O.mgb_solve(prob)
    _same_history(sd, so)


def test_illinois_and_no_finalize_options_are_passed_through():
    prob = default_problem("fem1d_5nodes", 1.5)
    bw = solver.barrier_weights(prob.M[0].w)
    sd = solver.mgb_solve(prob, handle=OracleHandle(prob, bw), line_search=1, finalize=False)
    so = O.mgb_solve(prob, line_search=O.linesearch_illinois(), finalize=False)
    _same_history(sd, so)


def test_phase1_escalation_handoff_and_failure_codes():
    prob = lower_bound_problem(50.0)
    sd = solver.mgb_solve(prob, handle=OracleHandle(prob))
    so = O.mgb_solve(prob)
    assert sd["SOL_feasibility"] is not None and "bounding box R=100.0" in sd["log"]
    assert np.array_equal(sd["SOL_feasibility"]["its"], so["SOL_feasibility"]["its"])
    _same_history(sd, so)
    with pytest.raises(solver.MGBConvergenceFailure) as e:
        p2 = lower_bound_problem(0.0, infeasible_pair=True)
        solver.mgb_solve(p2, handle=OracleHandle(p2))
    assert e.value.code == "infeasible"
    with pytest.raises(solver.MGBConvergenceFailure) as e:
        p3 = lower_bound_problem(1.0e6)
        solver.mgb_solve(p3, handle=OracleHandle(p3), feasibility_Rmax=1000.0)
    assert e.value.code == "feasibility_Rmax"
    with pytest.raises(solver.MGBConvergenceFailure) as e:
        p4 = default_problem("fem1d_3nodes", 1.0)
        solver.mgb_solve(p4, handle=OracleHandle(p4, solver.barrier_weights(p4.M[0].w)), tol=1e-50, maxit=25)
    assert e.value.code in ("stall", "iteration_limit")


def test_barrier_weights_follow_the_reference_rules():
    """convex.jl:279-304."""
    w = np.array([1.0, 0.0, 2.0, 0.0])
    assert np.allclose(solver.barrier_weights(w), [0.5, 0, 0.5, 0])
    assert solver.barrier_weights(np.ones(3)) is None                      # everything selected: the 1/n path
    assert solver.barrier_weights(w, ":") is None
    assert np.allclose(solver.barrier_weights(w, [0, 1]), [0.5, 0.5, 0, 0])
    assert np.allclose(solver.barrier_weights(w, np.array([True, False, False, True])), [0.5, 0, 0, 0.5])
    with pytest.raises(ValueError):
        solver.barrier_weights(w, np.array([], dtype=int))
    with pytest.raises(ValueError):
        solver.barrier_weights(w, np.array([True, False]))


def test_fixture_schedule_gate_accepts_only_the_documented_deviations():
    """tests/test_gpu_fixtures.py::compare_schedules is the gate the GPU runs are held to against the oracle fixtures.  It must pass
    identical and +-1 schedules, the two documented deviations (a fork explained by the kappa-growth threshold; a kappa refinement in
    the LAST barrier step), and reject everything else -- checked here on synthetic schedules, no GPU needed."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("gpu_fixtures_gate", os.path.join(os.path.dirname(__file__), "test_gpu_fixtures.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    gate = mod.compare_schedules
    ts = np.array([0.1, 1.0, 10.0, 100.0, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8])
    its = np.zeros((3, len(ts)), dtype=np.int64)
    its[2] = [12, 7, 6, 6, 6, 5, 5, 5, 5, 7]          # fine level; last step includes a finalize pass of 2
    its[0, 0], its[1, 0] = 8, 8
    fin = 2
    assert gate(ts, its, fin, ts, its, fin)["schedule"] == "identical"
    a = its.copy()
    a[2, 3] += 1
    assert gate(ts, a, fin, ts, its, fin)["schedule"] == "identical"                      # +-1 on a barrier step
    a[2, 3] += 1
    with pytest.raises(AssertionError):
        gate(ts, a, fin, ts, its, fin)                                                        # +2 is not
    a = its.copy()
    a[2, -1] += 5                                                                             # finalize pass longer: excluded ...
    assert gate(ts, a, fin + 5, ts, its, fin)["schedule"] == "identical"
    with pytest.raises(AssertionError):
        gate(ts, a, fin + 20, ts, its, fin)                                                   # ... but held to its halving tail
    # a crawl of 1 361 iterations may differ by 0.5 %
    c = its.copy()
    c[1, 0] = 1361
    d = c.copy()
    d[1, 0] = 1359
    assert gate(ts, d, fin, ts, c, fin)["schedule"] == "identical"
    d[1, 0] = 1340
    with pytest.raises(AssertionError):
        gate(ts, d, fin, ts, c, fin)
    # (1) fork explained by the kappa threshold: 4 vs 5 Newton iterations at step 5 -> the next t differs
    o = its.copy()
    o[2, 5] = 4
    to = np.array([0.1, 1.0, 10.0, 100.0, 1e3, 1e4, 1e5 * 3.1622776601683795, 1e6 * 3.1622776601683795, 1e7 * 3.1622776601683795, 1e8 * 3.1622776601683795])
    g = its.copy()
    g[2, 5] = 5
    assert gate(ts, g, fin, to, o, fin)["schedule"].startswith("forked at step 5")
    g[2, 5] = 6                                                                               # 4 vs 6: not explained
    with pytest.raises(AssertionError):
        gate(ts, g, fin, to, o, fin)
    g[2, 5] = 5
    g[2, 2] += 2                                                                              # an earlier step off by 2: rejected even with the fork
    with pytest.raises(AssertionError):
        gate(ts, g, fin, to, o, fin)
    # (2) the last barrier step: one side refined kappa (more than max_newton iterations on a level, different final t)
    r = its.copy()
    r[2, -1] = 9 + 6 + fin
    tr = ts.copy()
    tr[-1] = 7e7                                                                              # kappa halved once more in the last step
    out = gate(tr, r, fin, ts, its, fin)
    assert out["schedule"].startswith("identical up to the last barrier step")
    tr2 = ts.copy()
    tr2[-1] = 1.5e8                                                                           # different final t WITHOUT a refinement: rejected
    with pytest.raises(AssertionError):
        gate(tr2, its, fin, ts, its, fin)
    tr3 = tr.copy()
    tr3[-1] = 5e7                                                                             # final t below 1/tol: rejected
    with pytest.raises(AssertionError):
        gate(tr3, r, fin, ts, its, fin)
