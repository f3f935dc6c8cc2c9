"""Element partition of an assembled problem for multi-GPU runs.

One rank owns a contiguous block of elements, hence the broken nodes [e0*p, e1*p): every per-node
grid (w, f, g, convex grids, barrier weights), the operator blocks and the rows of the prolongations
are sliced to that block.  State variables whose fine-level space is the broken space itself (`:full`:
R block == identity) become *node-local* unknowns, stored only on the owning rank; all other unknowns
(continuous spaces, every coarser level) are *shared*: each rank holds the whole vector and partial sums
over elements (R' g, R'HR) are combined by an all-reduce inside the library.

What the library needs beyond an ordinary single-rank problem is carried in `AMG.n_global` and
`AMG.var_local`; everything else is just a smaller problem.
"""
from __future__ import annotations

import copy
from dataclasses import replace
from typing import List, Tuple

import numpy as np
import scipy.sparse as sp

from .geometry import Geometry
from .hierarchy import AMG
from .problem import Convex, MGBProblem, Piece


def element_range(N: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced split of N elements."""
    base, rem = divmod(N, world)
    e0 = rank * base + min(rank, rem)
    return e0, e0 + base + (1 if rank < rem else 0)


def node_local_variables(M: AMG) -> List[bool]:
    """A variable is node-local if its fine-level block of R is the n x n identity (a `:full` space)."""
    n = M.geometry.n
    R = sp.csr_matrix(M.R_fine[-1])
    offs = M.var_offsets[-1]
    out = []
    for v in range(M.nu):
        c0, c1 = offs[v], offs[v + 1]
        if c1 - c0 != n:
            out.append(False)
            continue
        blk = R[v * n:(v + 1) * n, c0:c1]
        out.append(blk.nnz == n and bool((blk.diagonal() == 1.0).all()))
    return out


def _shard_geometry(g: Geometry, e0: int, e1: int) -> Geometry:
    V = g.V
    ops = {k: np.ascontiguousarray(a[e0:e1]) for k, a in g.operators.items()}
    return replace(g, x=g.x[e0:e1], w=g.w[e0 * V:e1 * V], t=g.t[e0:e1], operators=ops)


def shard_amg(M: AMG, e0: int, e1: int) -> AMG:
    g = M.geometry
    if g.dense:
        raise ValueError("spectral (dense) discretisations do not partition by element: replicas only")
    n, V = g.n, g.V
    i0, i1 = e0 * V, e1 * V
    nl = i1 - i0
    local = node_local_variables(M)
    L = M.L
    offs = M.var_offsets[-1]
    R = sp.csr_matrix(M.R_fine[-1])
    new_offs = [0]
    for v in range(M.nu):
        new_offs.append(new_offs[-1] + (nl if local[v] else offs[v + 1] - offs[v]))
    blocks = []
    for v in range(M.nu):
        rows = R[v * n + i0:v * n + i1]
        row_blocks = []
        for u in range(M.nu):
            blk = rows[:, offs[u]:offs[u + 1]]
            if local[u]:
                blk = blk[:, i0:i1]
            row_blocks.append(blk)
        blocks.append(row_blocks)
    R_loc = sp.bmat(blocks, format="csr")
    R_loc.sort_indices()
    T = list(M.T)
    if L > 1:
        Tl = sp.csr_matrix(M.T[-1])
        parts = []
        for v in range(M.nu):
            rows = Tl[offs[v]:offs[v + 1]]
            parts.append(rows[i0:i1] if local[v] else rows)
        T[-1] = sp.vstack(parts, format="csr")
        T[-1].sort_indices()
    R_fine = list(M.R_fine[:-1]) + [R_loc]
    var_offsets = [list(o) for o in M.var_offsets[:-1]] + [new_offs]
    out = AMG(geometry=_shard_geometry(g, e0, e1), w=M.w[i0:i1], R_fine=R_fine, T=T, D=M.D,
              state_variables=M.state_variables, var_offsets=var_offsets, dense=False)
    out.n_global = n
    out.var_local = [int(b) for b in local]
    out.node_range = (i0, i1)
    return out


def _shard_convex(Q: Convex, i0: int, i1: int) -> Convex:
    pieces = []
    for pc in Q.pieces:
        q = copy.copy(pc)
        q.A = pc.A[i0:i1]
        q.b = pc.b[i0:i1]
        q.p = pc.p[i0:i1] if pc.p is not None else None
        q.mu = pc.mu[i0:i1] if pc.mu is not None else None
        pieces.append(q)
    out = copy.copy(Q)
    out.pieces = pieces
    out.select = Q.select[i0:i1] if Q.select is not None else None
    return out


def shard_problem(prob: MGBProblem, rank: int, world: int) -> MGBProblem:
    """The sub-problem of `rank`: its elements, the shared unknowns, its node-local unknowns."""
    g = prob.geometry
    e0, e1 = element_range(g.N, rank, world)
    i0, i1 = e0 * g.V, e1 * g.V
    M1 = shard_amg(prob.M[0], e0, e1)
    M2 = shard_amg(prob.M[1], e0, e1) if prob.M[1] is not None else None
    out = MGBProblem((M1, M2), prob.f[i0:i1], prob.g[i0:i1], _shard_convex(prob.Q, i0, i1), M1.geometry)
    out.node_range = (i0, i1)
    out.n_global = g.n
    return out


def shard_barrier_weights(bw, i0, i1):
    return None if bw is None else np.ascontiguousarray(bw[i0:i1])
