"""Shared problem builders for the tests (the same calls the reference's tests make)."""
import json
import os

import numpy as np

import mgbx
from mgbx import geometry as G, hierarchy as H, problem as P

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                                   "reference_goldens.json")))


def gold(name):
    e = GOLD[name]
    return np.array(e["data"]).reshape(e["shape"], order="F")


GEOMS = {
    "fem1d_3nodes": lambda: G.fem1d(nodes=np.linspace(-1, 1, 3)),
    "fem1d_5nodes": lambda: G.fem1d(nodes=np.linspace(-1, 1, 5)),
    "fem2d_P2_quickstart": lambda: G.fem2d_P2(),
    "fem2d_P1_L2": lambda: G.subdivide(G.fem2d_P1(), 2),
    "fem2d_P2_L2": lambda: G.subdivide(G.fem2d_P2(), 2),
    "fem3d_k1_L2": lambda: G.subdivide(G.fem3d(k=1), 2),
    "spectral1d_n5": lambda: G.spectral1d(n=5),
    "spectral2d_n5": lambda: G.spectral2d(n=5),
    "spectral1d_n4": lambda: G.spectral1d(n=4),
    "spectral2d_n4": lambda: G.spectral2d(n=4),
}

SOLVE_CASES = [
    ("fem1d_3nodes_p1", "fem1d_3nodes", 1.0),
    ("fem2d_P2_quickstart_p1", "fem2d_P2_quickstart", 1.0),
    ("spectral1d_n5_p1", "spectral1d_n5", 1.0),
    ("spectral2d_n5_p1", "spectral2d_n5", 1.0),
    ("fem1d_5nodes_p1", "fem1d_5nodes", 1.0),
    ("fem1d_5nodes_p1.5", "fem1d_5nodes", 1.5),
    ("fem2d_P1_L2_p1", "fem2d_P1_L2", 1.0),
    ("fem2d_P1_L2_p1.5", "fem2d_P1_L2", 1.5),
    ("fem2d_P2_L2_p1", "fem2d_P2_L2", 1.0),
    ("fem2d_P2_L2_p1.5", "fem2d_P2_L2", 1.5),
    ("fem3d_k1_L2_p1", "fem3d_k1_L2", 1.0),
    ("fem3d_k1_L2_p1.5", "fem3d_k1_L2", 1.5),
]

PARABOLIC_CASES = [
    ("parabolic_fem1d_3nodes_h0.5_p1", "fem1d_3nodes"),
    ("parabolic_fem2d_P2_h0.5_p1", "fem2d_P2_quickstart"),
    ("parabolic_spectral1d_n4_h0.5_p1", "spectral1d_n4"),
    ("parabolic_spectral2d_n4_h0.5_p1", "spectral2d_n4"),
]


def default_problem(geom_name, p):
    return P.assemble(H.amg(GEOMS[geom_name]()), p=p)


def lower_bound_problem(lower, nodes=5, infeasible_pair=False):
    """test/test_feasibility.jl:13-22,44-51."""
    mg = H.amg(G.fem1d(nodes=np.linspace(-1, 1, nodes)))
    n = mg.geometry.n
    if infeasible_pair:
        Q = P.convex_linear(mg, idx=(0,), A_grid=np.tile([1.0, -1.0], (n, 1)),
                            b_grid=np.tile([-1.0, 0.0], (n, 1)))
    else:
        Q = P.convex_linear(mg, idx=(0,), A_grid=np.ones((n, 1)), b_grid=np.full((n, 1), -lower))
    return P.assemble(mg, state_variables=[("u", "full")], D=[("u", "id")],
                      f=lambda x: np.array([1.0]), g=lambda x: np.array([0.0]), Q=Q)
