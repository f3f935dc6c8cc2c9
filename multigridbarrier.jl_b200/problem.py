"""Convex-set descriptors and `assemble` (host side, pure data).

Reference interface mirrored (files under /root/reference/src):
  convex.jl:80-122                  Convex, intersect
  convex_euclidian_power.jl:352-453 convex_Euclidian_power (args = A, b, p, mu grids)
  convex_linear.jl:87-222           convex_linear (args = A, b grids)
  convex_piecewise.jl:110-182       convex_piecewise (args = select grid + piece args)
  mgb.jl:587-613, 711-727           defaults and assemble -> MGBProblem

In the reference a Convex carries Julia functors; its *data* are the per-node grids in `args`.
Here a Convex is only the data plus a kind tag: that is exactly what crosses the C ABI
("Convex sets cross the ABI as descriptors").  idx is 0-based.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence

import numpy as np

from .geometry import Geometry
from .hierarchy import AMG, MultiGrid, prepare_amg

__all__ = ["Piece", "Convex", "convex_Euclidian_power", "convex_linear", "convex_piecewise",
           "intersect", "MGBProblem", "assemble", "default_D", "default_f", "default_g",
           "default_idx", "map_rows"]

KIND_EP, KIND_LINEAR = 0, 1


def map_rows(f: Callable, *grids) -> np.ndarray:
    """Row-wise map (utils.jl:122-126): f receives one row of each grid."""
    n = grids[0].shape[0]
    out = [np.atleast_1d(np.asarray(f(*[g[i] for g in grids]), dtype=float)) for i in range(n)]
    return np.stack(out, axis=0)


@dataclass
class Piece:
    kind: int                 # KIND_EP | KIND_LINEAR
    idx: Optional[tuple]      # 0-based D-row indices, or None for Colon
    ni: int                   # number of indexed inputs (EP: nz)
    nc: int                   # LINEAR: constraint rows; EP: nz
    A: np.ndarray             # (n, nc*ni), per-node A flattened column-major
    b: np.ndarray             # (n, nc)
    p: Optional[np.ndarray] = None    # EP only, (n,)
    mu: Optional[np.ndarray] = None   # EP only, (n,)

    def min_inputs(self):
        return self.ni if self.idx is None else max(self.idx) + 1


@dataclass
class Convex:
    """pieces + optional select grid (n, K); a plain EP/linear set is K = 1 without select."""
    pieces: List[Piece]
    select: Optional[np.ndarray] = None

    def validate(self, nD: int):
        """convex.jl:54-97 input-spec validation."""
        for pc in self.pieces:
            if pc.idx is None:
                if pc.ni != nD:
                    raise ValueError("convex constraint with idx = Colon() expects exactly %d D row(s), "
                                     "but D has %d row(s)" % (pc.ni, nD))
            elif max(pc.idx) + 1 > nD:
                raise ValueError("convex constraint indexes input row %d, but D has only %d row(s)"
                                 % (max(pc.idx) + 1, nD))


def _norm_idx(idx):
    if idx is None:
        return None
    idx = tuple(int(i) for i in idx)
    if len(idx) == 0:
        raise ValueError("idx must contain at least one input row")
    if min(idx) < 0:
        raise ValueError("idx entries must be non-negative (0-based)")
    return idx


def _grid(fn, x, shape_cols=None):
    g = map_rows(fn, x)
    return g


def convex_Euclidian_power(mg: MultiGrid, idx=None, A=None, b=None, p=None,
                           A_grid=None, b_grid=None, p_grid=None) -> Convex:
    """Power cone {y : s >= ||q||^p, [q; s] = A y[idx] + b} (convex_euclidian_power.jl:352-453).
    A(x) -> (nz, nz) matrix (default identity), b(x) -> (nz,) or scalar (placed in the last
    slot), p(x) -> exponent (default 2)."""
    x = mg.geometry.xflat()
    n = x.shape[0]
    idx = _norm_idx(idx)
    if A_grid is None:
        if A is None:
            if idx is None:
                raise ValueError("a default A with idx = Colon() cannot determine the constraint "
                                 "dimension; pass an explicit idx, or a matrix-valued A.")
            nz = len(idx)
            A_grid = np.asfortranarray(np.tile(np.eye(nz).reshape(-1, order="F"), (n, 1)))   # column-major like the reference's grids
        else:
            A_grid = map_rows(lambda xi: np.asarray(A(xi), float).reshape(-1, order="F"), x)
    nz = len(idx) if idx is not None else int(round(np.sqrt(A_grid.shape[1])))
    if nz * nz != A_grid.shape[1]:
        raise ValueError("A_grid has %d columns per node; expected nz^2 = %d" % (A_grid.shape[1], nz * nz))
    if b_grid is None:
        if b is None:
            b_grid = np.zeros((n, nz), order="F")
        else:
            def bf(xi):
                bx = b(xi)
                if np.isscalar(bx):
                    v = np.zeros(nz)
                    v[-1] = bx
                    return v
                return np.asarray(bx, float)
            b_grid = map_rows(bf, x)
    b_grid = np.asarray(b_grid, float).reshape(n, -1)
    if b_grid.shape[1] != nz:
        raise ValueError("b_grid has %d value(s) per node but nz = %d" % (b_grid.shape[1], nz))
    if p_grid is None:
        p_grid = np.full(n, 2.0) if p is None else map_rows(lambda xi: float(p(xi)), x).reshape(n)
    p_grid = np.asarray(p_grid, float).reshape(n)
    mu = np.where((p_grid == 2) | (p_grid == 1), 0.0, np.where(p_grid < 2, 1.0, 2.0))
    return Convex([Piece(KIND_EP, idx, nz, nz, np.asarray(A_grid, float), b_grid, p_grid, mu)])


def convex_linear(mg: MultiGrid, idx=None, A=None, b=None, A_grid=None, b_grid=None) -> Convex:
    """{y : A y[idx] + b > 0} with a log barrier per row (convex_linear.jl:87-222)."""
    x = mg.geometry.xflat()
    n = x.shape[0]
    idx = _norm_idx(idx)
    if A_grid is None:
        if A is None:
            if idx is None:
                raise ValueError("a default A with idx = Colon() cannot determine the constraint size")
            m = len(idx)
            A_grid = np.tile(np.eye(m).reshape(-1, order="F"), (n, 1))
        else:
            A_grid = map_rows(lambda xi: np.atleast_2d(np.asarray(A(xi), float)).reshape(-1, order="F"), x)
    A_grid = np.asarray(A_grid, float).reshape(n, -1)
    if b_grid is None:
        if b is None:
            bs = 0.0
        else:
            bs = b(x[0])
        if b is None or np.isscalar(bs):
            if idx is not None:
                nc = A_grid.shape[1] // len(idx)
            else:
                nc = np.atleast_2d(np.asarray(A(x[0]))).shape[0]
            b_grid = map_rows(lambda xi: np.full(nc, 0.0 if b is None else float(b(xi))), x)
        else:
            b_grid = map_rows(lambda xi: np.asarray(b(xi), float), x)
    b_grid = np.asarray(b_grid, float).reshape(n, -1)
    nc = b_grid.shape[1]
    if A_grid.shape[1] % nc != 0:
        raise ValueError("A_grid has %d columns per node, not a multiple of the %d constraint row(s)"
                         % (A_grid.shape[1], nc))
    ni = A_grid.shape[1] // nc
    if idx is not None and len(idx) != ni:
        raise ValueError("A_grid implies ni = %d but len(idx) = %d" % (ni, len(idx)))
    return Convex([Piece(KIND_LINEAR, idx, ni, nc, A_grid, b_grid)])


def convex_piecewise(mg: MultiGrid, Q: Sequence[Convex], select=None, select_grid=None) -> Convex:
    """Sum of the pieces active at each node (convex_piecewise.jl:110-182).  Nested piecewise
    sets are flattened (their selects multiply)."""
    x = mg.geometry.xflat()
    n = x.shape[0]
    K = len(Q)
    if select_grid is None:
        if select is None:
            select_grid = np.ones((n, K))
        else:
            select_grid = map_rows(lambda xi: np.asarray(select(xi), float), x)
    select_grid = np.asarray(select_grid, float).reshape(n, K)
    pieces, cols = [], []
    for k, q in enumerate(Q):
        for j, pc in enumerate(q.pieces):
            pieces.append(pc)
            col = (select_grid[:, k] != 0).astype(float)
            if q.select is not None:
                col = col * (q.select[:, j] != 0)
            cols.append(col)
    return Convex(pieces, np.stack(cols, axis=1))


def intersect(mg: MultiGrid, *Q: Convex) -> Convex:
    """All pieces active everywhere (convex.jl:116-122)."""
    return convex_piecewise(mg, Q)


# ---- defaults (mgb.jl:587-613) ---------------------------------------------------------------

def default_f(dim):
    return lambda x: np.array([0.5] + [0.0] * dim + [1.0])


def default_g(dim):
    if dim == 1:
        return lambda x: np.array([x[0], 2.0])
    return lambda x: np.array([float(np.dot(x, x)), 100.0])


def default_f_grid(dim, x):
    """map_rows(default_f(dim), x) without the per-node Python call."""
    out = np.zeros((x.shape[0], dim + 2))
    out[:, 0] = 0.5
    out[:, dim + 1] = 1.0
    return out


def default_g_grid(dim, x):
    """map_rows(default_g(dim), x) without the per-node Python call; the sum of squares is formed left to right exactly as
    the reference writes it (x[1]^2+x[2]^2+x[3]^2, src/mgb.jl:592-593)."""
    n = x.shape[0]
    if dim == 1:
        return np.stack([x[:, 0].astype(float), np.full(n, 2.0)], axis=1)
    s = x[:, 0] * x[:, 0]
    for d in range(1, dim):
        s = s + x[:, d] * x[:, d]
    return np.stack([s, np.full(n, 100.0)], axis=1)


def default_D(dim):
    return [("u", "id")] + [("u", n) for n in ("dx", "dy", "dz")[:dim]] + [("s", "id")]


def default_idx(dim):
    return tuple(range(1, dim + 2))


def default_slack_space(geom: Geometry):
    """multigrid.jl:420, fem2d_P2.jl:70."""
    return "broken_P1" if (geom.kind == "p2" and not geom.bubble) else "full"


@dataclass
class MGBProblem:
    """Closure-free problem: pure arrays (mgb.jl:666-674)."""
    M: tuple            # (main AMG, feasibility AMG)
    f: np.ndarray       # (n, nD)
    g: np.ndarray       # (n, nu)
    Q: Convex
    geometry: Geometry


def assemble(mg: MultiGrid, dim=None, state_variables=None, D=None, x=None, p=1.0,
             g=None, f=None, g_grid=None, f_grid=None, Q: Optional[Convex] = None, M=None,
             **_ignored) -> MGBProblem:
    """Lower a problem specification to pure data (mgb.jl:711-727)."""
    geom = mg.geometry
    dim = geom.dim if dim is None else dim
    if state_variables is None:
        state_variables = [("u", "dirichlet"), ("s", default_slack_space(geom))]
    if D is None:
        D = default_D(dim)
    if x is None:
        x = geom.xflat()
    if g_grid is None:
        g_grid = default_g_grid(dim, x) if g is None else map_rows(g, x)
    if f_grid is None:
        f_grid = default_f_grid(dim, x) if f is None else map_rows(f, x)
    if Q is None:
        Q = convex_Euclidian_power(mg, idx=default_idx(dim), p_grid=np.full(x.shape[0], float(p)))
    if M is None:
        M = prepare_amg(mg, state_variables, D)
    Q.validate(M[0].nD)
    # grids are kept column-major, the layout the reference's (Julia) arrays have and the C ABI takes -- no conversion at the boundary
    return MGBProblem(M, np.asfortranarray(f_grid, dtype=float), np.asfortranarray(g_grid, dtype=float), Q, geom)
