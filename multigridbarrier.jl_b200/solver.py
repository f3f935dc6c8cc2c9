"""Host mirror of the reference's outer drivers over the C ABI (what the Julia `B200Device` shim does).

  mgb_solve        src/mgb.jl:798-842   native_to_device -> mgb_driver -> device_to_native
  mgb_driver       src/mgb.jl:332-584   feasibility probe, phase I with box escalation, _matched_t, main ramp
  mgb_core         src/mgb.jl:91-183    t-ramp with kappa adaptation (host scalars only)
  parabolic_solve  src/Parabolic.jl:126-173

Everything below `mgb_core`'s loop body -- mgb_step, newton, the line search, f0/f1/f2, the
Hessian assembly and the linear solve -- runs on the device inside `mgbx_step`; only scalars come
back.  There is no CPU path: a missing library or GPU raises.
"""
from __future__ import annotations

import math
import time
from typing import Optional

import numpy as np

from . import native
from .problem import (assemble, convex_Euclidian_power, default_slack_space, intersect)
from .hierarchy import prepare_amg

EPS = np.finfo(np.float64).eps


class MGBConvergenceFailure(Exception):
    """src/utils.jl:157-184; code in {'infeasible','feasibility_Rmax','stall','iteration_limit','failure'}."""

    def __init__(self, message, code="failure"):
        super().__init__(message)
        self.message = message
        self.code = code


def barrier_weights(w, barrier_nodes=None):
    """src/convex.jl:279-304 with the default mask w != 0 (src/mgb.jl:377)."""
    n = len(w)
    if barrier_nodes is None:
        sel = (w != 0).astype(float)
    elif isinstance(barrier_nodes, str) and barrier_nodes == ":":
        return None
    else:
        bn = np.asarray(barrier_nodes)
        if bn.dtype == bool:
            if bn.size != n:
                raise ValueError("barrier_nodes mask has length %d but the mesh has %d nodes" % (bn.size, n))
            sel = bn.astype(float)
        else:
            if bn.size == 0:
                raise ValueError("barrier_nodes must select at least one node")
            sel = np.zeros(n)
            sel[bn] = 1.0
    m = sel.sum()
    if m <= 0:
        raise ValueError("barrier_nodes selects no nodes")
    if m == n:
        return None
    return sel / m


def _jlfloat(x):
    return repr(float(x))


def mgb_core(h: native.Handle, which, L, tol=math.sqrt(EPS), t=0.1, maxit=10000, kappa=10.0,
             early_stop=None, max_newton=None, finalize=True, finalize_theta=0.9, stop_kind=1,
             stop_lambda_tol=None, stop_theta=0.9, ls_beta=0.5, ls_c1=0.1, line_search=0, log=None,
             stats=None):
    """src/mgb.jl:91-183 -- the t-ramp.  State z lives on the device; early_stop(t) -> bool."""
    if max_newton is None:
        max_newton = int(math.ceil(math.log2(-math.log2(EPS)) + 2))
    if early_stop is None:
        early_stop = lambda t: False
    target = 1.0 / tol
    kappa0 = kappa
    its, ts, kappas, cdz, times = [], [], [], [], []
    fin = [0]                 # Newton iterations of the finalize pass of the last accepted step
    t_begin = time.time()

    def opts(tt, initial):
        o = h.step_opts(maxit=maxit, max_newton=max_newton, initial_step=int(initial), stop_kind=stop_kind,
                        stop_theta=stop_theta, finalize=int(bool(finalize) and tt >= target),
                        finalize_theta=finalize_theta, line_search=line_search, ls_beta=ls_beta, ls_c1=ls_c1)
        if stop_lambda_tol is not None:
            o.stop_lambda_tol = stop_lambda_tol
        return o

    def run(tt, initial):
        rc, r = h.step(which, tt, opts(tt, initial))
        if stats is not None:
            stats["f01_evals"] += r.f01_evals
            stats["f2_evals"] += r.f2_evals
            stats["linear_solves"] += r.linear_solves
            stats["pcg_iters"] += r.pcg_iters
            stats["solve_failures"] = stats.get("solve_failures", 0) + r.solve_failures
            stats["direct_fallbacks"] = stats.get("direct_fallbacks", 0) + r.direct_fallbacks
            stats["ms_f01"] += r.ms_f01
            stats["ms_f2"] += r.ms_f2
            stats["ms_solve"] += r.ms_solve
        if rc == native.NON_FINITE:
            raise FloatingPointError("newton: non-finite objective, gradient or direction at t=%g" % tt)
        if rc == native.OK:
            fin[0] = int(r.its_finalize)
        return rc == native.OK, np.array(r.its[:L], dtype=np.int64)

    ok, it0 = run(t, True)
    if not ok:
        raise MGBConvergenceFailure("Initial centering failed in mgb_solve at t=%g, tol=%g, maxit=%d."
                                    % (t, tol, maxit), "stall")
    k = 1
    its.append(it0)
    kappas.append(kappa)
    ts.append(t)
    times.append(time.time())
    cdz.append(h.scalars(which).c_dot_Dz)
    while t < target and kappa > 1 and k < maxit and not early_stop(t):
        k += 1
        itk = np.zeros(L, dtype=np.int64)
        while kappa > 1:
            t1 = kappa * t
            ok, iti = run(t1, False)
            itk += iti
            if log is not None:
                log("mgb_core: t=%g kappa=%g its=%s converged=%s" % (t1, kappa, iti.tolist(), ok))
            if ok:
                if iti.max() <= max_newton * 0.5:
                    kappa = min(kappa0, kappa ** 2)
                t = t1
                break
            kappa = math.sqrt(kappa)
        its.append(itk)
        ts.append(t)
        kappas.append(kappa)
        times.append(time.time())
        cdz.append(h.scalars(which).c_dot_Dz)
    converged = (t >= target) or early_stop(t)
    if not converged:
        code = "stall" if kappa <= 1 else "iteration_limit"
        raise MGBConvergenceFailure("Convergence failure in mgb_solve at t=%g, k=%d, kappa=%g, tol=%g, maxit=%d."
                                    % (t, k, kappa, tol, maxit), code)
    return dict(its=np.stack(its, axis=1), ts=np.array(ts), kappas=np.array(kappas),
                c_dot_Dz=np.array(cdz), times=np.array(times), t_begin=t_begin,
                t_elapsed=time.time() - t_begin, its_finalize=fin[0])


def mgb_driver(h: native.Handle, M, t=0.1, t_feasibility=None, feasibility_Rmax=1.0 / math.sqrt(EPS),
               finalize=True, log=None, stats=None, **rest):
    """src/mgb.jl:332-584 on a device-resident problem."""
    if t_feasibility is None:
        t_feasibility = t
    if log is None:
        log = lambda *a: None
    M1, M2 = M
    n = int(getattr(M1, "n_global", 0) or len(M1.w))     # the WHOLE mesh's node count (a multi-GPU rank holds a slice)
    n_loc = len(M1.w)
    ncomp = M1.nu
    L1 = len(M1.R_fine)
    rest.setdefault("stop_lambda_tol", 0.25 / math.sqrt(n))
    SOL_feas = None
    need, b, zabs = h.phase1_init()
    if need:
        L2 = len(M2.R_fine)
        Rbox = max(10.0, 10.0 * zabs)
        Rmax = max(float(feasibility_Rmax), Rbox)

        def feasible():
            return h.scalars(native.FEAS).var_max[ncomp] < 0

        first = True
        while True:
            log("mgb_driver: feasibility phase with bounding box R=%s" % _jlfloat(Rbox))
            h.set_feasibility_box(b, Rbox)
            if not first:
                h.reset_feasibility_state()      # no warm start between box rounds (src/mgb.jl:538-543)
            first = False
            failure = None
            t_first = [math.inf]

            def feas_stop(tt):
                if not feasible():
                    return False
                t_first[0] = min(t_first[0], tt)
                return tt >= 2 * t_first[0]
            try:
                SOL_feas = mgb_core(h, native.FEAS, L2, t=t_feasibility, early_stop=feas_stop,
                                    finalize=finalize, log=log, stats=stats, **rest)
            except KeyboardInterrupt:
                raise
            except native.MgbxError:
                raise
            except Exception as e2:               # broad on purpose (src/mgb.jl:510-521)
                failure = e2
            if failure is None:
                if feasible():
                    break
                sc = h.scalars(native.FEAS)
                vmax = max(sc.var_absmax[k] for k in range(ncomp))
                smax = sc.var_max[ncomp]
                if vmax <= Rbox / 2:
                    raise MGBConvergenceFailure(
                        "The problem appears to be infeasible: the feasibility subproblem converged to a "
                        "minimizer with positive constraint violation (max slack ~ %g) strictly inside the "
                        "bounding box (max |nodal value| ~ %g <= R/2 with R = %g)." % (smax, vmax, Rbox),
                        "infeasible")
                log("mgb_driver: phase-I minimizer presses the box; growing R")
            else:
                log("mgb_driver: feasibility solve failed at R=%s: %s" % (_jlfloat(Rbox), failure))
            Rnext = 10 * Rbox
            if Rnext > Rmax:
                raise MGBConvergenceFailure(
                    "Could not find a strictly feasible point with nodal values bounded by R = %g "
                    "(cap feasibility_Rmax ~ %g). The problem is infeasible, or its feasible points have "
                    "nodal values exceeding the cap (rescale the problem, or raise feasibility_Rmax)."
                    % (Rbox, Rmax), "feasibility_Rmax")
            Rbox = Rnext
        h.handoff()
        tm, tstar = h.matched_t(t)
        if tm < t or tm == t:
            if math.isfinite(tstar) and tstar > 0 and tm != t:
                log("_matched_t: warm start matches t=%r, starting main ramp at t=%r" % (tstar, tm))
        t = min(t, tm)
    SOL_main = mgb_core(h, native.MAIN, L1, t=t, finalize=finalize, log=log, stats=stats, **rest)
    z = h.get_z(native.MAIN).reshape(ncomp, n_loc).T.copy()
    SOL_main["z_unfinalized"] = h.get_z_unfinalized(native.MAIN).reshape(ncomp, n_loc).T.copy()   # src/mgb.jl:76-80
    return dict(z=z, SOL_feasibility=SOL_feas, SOL_main=SOL_main)


def mgb_solve(prob, barrier_nodes=None, config=None, log=None, handle=None, comm=None, **kw):
    """src/mgb.jl:798-842.  Returns dict(z, SOL_main, SOL_feasibility, log, geometry, stats).

    comm = (rank, world, nccl_id): multi-GPU run, one process per GPU.  `prob` is the WHOLE problem on every rank; this
    rank keeps the element block partition.element_range(N, rank, world) and returns its rows of z in sol["z"]
    (sol["node_range"] = (i0, i1)); the scalar histories (its, ts, c_dot_Dz) are identical on all ranks."""
    lines = []
    node_range = None
    if comm is not None and comm[1] > 1:
        from . import partition
        bw_full = barrier_weights(prob.M[0].w, barrier_nodes)
        prob = partition.shard_problem(prob, comm[0], comm[1])
        node_range = prob.node_range
        bw_shard = partition.shard_barrier_weights(bw_full, *node_range)
    else:
        comm = None

    def _log(*a):
        s = "".join(str(x) for x in a)
        lines.append(s)
        if log is not None:
            log(s)
    bw = bw_shard if comm is not None else barrier_weights(prob.M[0].w, barrier_nodes)
    own = handle is None
    t0 = time.time()
    h = handle if handle is not None else native.Handle(prob, barrier_weights=bw, comm=comm, **(config or {}))
    t_create = time.time() - t0
    stats = dict(f01_evals=0, f2_evals=0, linear_solves=0, pcg_iters=0, solve_failures=0, direct_fallbacks=0, ms_f01=0.0, ms_f2=0.0, ms_solve=0.0)
    try:
        l0 = h.launch_count()
        sol = mgb_driver(h, prob.M, log=_log, stats=stats, **kw)
        stats["gpu_launches"] = h.launch_count() - l0
        stats["device_memory_report"], stats["device_bytes"] = h.memory_report()
    finally:
        if own:
            h.close()
    stats["create_s"] = t_create
    sol["log"] = "\n".join(lines)
    sol["geometry"] = prob.geometry
    sol["node_range"] = node_range
    sol["stats"] = stats
    return sol


def parabolic_solve(mg, p=1.0, h=0.2, t0=0.0, t1=1.0, ts=None, f1=None, g=None, state_variables=None,
                    D=None, Q=None, config=None, **rest):
    """src/Parabolic.jl:126-173: implicit Euler; one device solve per time step on a handle that is
    created once (the hierarchy, plans and operators stay resident; only f_grid / g_grid change)."""
    geom = mg.geometry
    dim = geom.dim
    if ts is None:
        ts = np.arange(t0, t1 + 0.5 * h, h)
    x = geom.xflat()
    n = x.shape[0]
    default_f1, default_g = f1 is None, g is None      # the defaults are evaluated vectorised (no per-node Python call)
    if f1 is None:
        f1 = lambda t, xx: 0.5
    if g is None:
        g = lambda t, xx: np.array([float(np.dot(xx, xx)) if dim > 1 else float(xx[0]), 0.0, 0.0])
    if state_variables is None:
        spc = default_slack_space(geom)
        state_variables = [("u", "dirichlet"), ("s1", spc), ("s2", spc)]
    if D is None:
        D = [("u", "id")] + [("u", nm) for nm in ("dx", "dy", "dz")[:dim]] + [("s1", "id"), ("s2", "id")]
    if Q is None:
        idx1 = (0, dim + 1)
        idx2 = tuple(range(1, dim + 1)) + (dim + 2,)
        Q = intersect(mg, convex_Euclidian_power(mg, idx=idx1, p_grid=np.full(n, 2.0)),
                      convex_Euclidian_power(mg, idx=idx2, p_grid=np.full(n, float(p))))
    if default_f1:
        f1_grid = np.full((n, len(ts)), 0.5)
    else:
        f1_grid = np.array([[f1(ts[j], x[i]) for j in range(len(ts))] for i in range(n)])
    if default_g:
        g0 = x[:, 0] * x[:, 0] if dim > 1 else x[:, 0].astype(float)
        for d in range(1, dim):
            g0 = g0 + x[:, d] * x[:, d]
        U = [np.stack([g0, np.zeros(n), np.zeros(n)], axis=1) for _ in range(len(ts))]
    else:
        U = [np.array([g(ts[k], x[i]) for i in range(n)], dtype=float) for k in range(len(ts))]
    M = prepare_amg(mg, state_variables, D)
    handle = None
    step_stats = []
    try:
        for k in range(len(ts) - 1):
            j = k + 1
            dt = ts[j] - ts[j - 1]
            fg = np.zeros((n, len(D)))
            fg[:, 0] = dt * f1_grid[:, j] - U[k][:, 0]
            fg[:, dim + 1] = 0.5
            fg[:, dim + 2] = dt / p
            prob = assemble(mg, M=M, g_grid=U[k + 1], f_grid=fg, Q=Q, state_variables=state_variables, D=D)
            if handle is None:
                bw = barrier_weights(prob.M[0].w, rest.get("barrier_nodes"))
                handle = native.Handle(prob, barrier_weights=bw, **(config or {}))
            else:
                handle.set_grids(fg, U[k + 1])
            sol = mgb_solve(prob, handle=handle, **rest)
            U[k + 1] = sol["z"]
            step_stats.append(dict(sol["stats"], newton_main=int(sol["SOL_main"]["its"].sum()),
                                   newton_feas=int(sol["SOL_feasibility"]["its"].sum()) if sol["SOL_feasibility"] else 0))
    finally:
        if handle is not None:
            handle.close()
    return dict(geometry=geom, ts=np.asarray(ts), u=U, stats=step_stats)
