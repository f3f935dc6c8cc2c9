"""Pins the CPU oracle to the reference's own golden vectors (tol 1e-6 as in
test/runtests.jl:13-52, test/test_algebraic.jl:13) and failure semantics
(test/test_feasibility.jl:24-87)."""
import numpy as np
import pytest

import mgb_oracle as O
from helpers import (PARABOLIC_CASES, SOLVE_CASES, GEOMS, default_problem, gold, lower_bound_problem)
from mgbx import hierarchy as H, problem as P

TOL = 1e-6


@pytest.mark.parametrize("name,geom,p", SOLVE_CASES)
def test_mgb_solve_golden(name, geom, p):
    sol = O.mgb_solve(default_problem(geom, p))
    assert np.linalg.norm(sol["z"] - gold(name)) < TOL


@pytest.mark.parametrize("name,geom", PARABOLIC_CASES)
def test_parabolic_golden(name, geom):
    sol = O.parabolic_solve(H.amg(GEOMS[geom]()), P.assemble, P.intersect, P.convex_Euclidian_power,
                            H.prepare_amg, P.default_slack_space, h=0.5, p=1.0)
    assert np.linalg.norm(np.stack(sol["u"], axis=2) - gold(name)) < TOL


def test_feasibility_escalation():
    sol = O.mgb_solve(lower_bound_problem(50.0))
    assert sol["SOL_feasibility"] is not None
    assert np.max(np.abs(sol["z"] - 50.0)) < 1e-3
    assert "bounding box R=100.0" in sol["log"]


def test_infeasible_certified():
    with pytest.raises(O.MGBConvergenceFailure) as e:
        O.mgb_solve(lower_bound_problem(0.0, infeasible_pair=True))
    assert e.value.code == "infeasible"
    assert "appears to be infeasible" in e.value.message


def test_feasibility_rmax():
    with pytest.raises(O.MGBConvergenceFailure) as e:
        O.mgb_solve(lower_bound_problem(1.0e6), feasibility_Rmax=1000.0)
    assert e.value.code == "feasibility_Rmax"


def test_feasible_start_skips_phase1():
    sol = O.mgb_solve(lower_bound_problem(-50.0))
    assert sol["SOL_feasibility"] is None
    assert np.max(np.abs(sol["z"] + 50.0)) < 1e-3


def test_linear_cobarrier_hessian_identity():
    """test/test_algebraic_coverage.jl:48-59: rectangular cobarrier Hessian == B' diag(1/F^2) B."""
    rng = np.random.default_rng(0)
    n, nc, ni = 5, 3, 2
    A = rng.normal(size=(n, nc, ni))
    pc = P.Piece(P.KIND_LINEAR, None, ni, nc, A.transpose(0, 2, 1).reshape(n, -1), np.full((n, nc), 5.0))
    Y = np.concatenate([0.1 * rng.normal(size=(n, ni)), np.full((n, 1), 1.0)], axis=1)
    _, g, Hm = O.piece_eval(pc, Y, 2, cobarrier=True)
    for i in range(n):
        B = np.concatenate([A[i], np.ones((nc, 1))], axis=1)
        F = B @ Y[i] + 5.0
        assert np.allclose(Hm[i], B.T @ np.diag(1 / F ** 2) @ B, rtol=1e-13)
        assert np.allclose(g[i], -B.T @ (1 / F), rtol=1e-13)


def test_newton_exact_quadratic():
    """test/test_algebraic_coverage.jl:85-97: k == 1 and converged on an exact quadratic."""
    A = np.array([[2.0, 0.0], [0.0, 4.0]])
    sol = O.newton(lambda x: 0.5 * x @ A @ x, lambda x: A @ x, lambda x: A, np.array([1.0, 1.0]),
                   stopping_criterion=O.stopping_exact(0.1), line_search=O.linesearch_backtracking())
    assert sol["converged"] and sol["k"] <= 2


def test_illinois_edge():
    assert O.illinois(lambda x: x - 1.0, 0.0, 2.0) == pytest.approx(1.0)
    assert O.illinois(lambda x: x, 0.0, 2.0) == 0.0
