"""Host-side (CPU, setup-time) multigrid hierarchy construction: amg(), MultiGrid, AMG pair.

Reference behaviour restated (files under /root/reference/src):
  amg_prolongators.jl:16-78   prolongator factories, ladder assembly
  multigrid.jl:192-265        _compose_R, _stretch_per_subspace
  multigrid.jl:372-412        _assemble_amg_dicts (:full / :uniform / riders / dirichlet classes)
  multigrid.jl:474-538        amg_helper / _prepare_amg (main + feasibility AMG pair)
  fem2d_P1.jl:72-126, fem2d_P2.jl:388-455, TensorFEM.jl:686-796   per-discretization amg()
  spectral1d.jl:63-109, spectral2d.jl:15-42                         spectral hierarchies

The classical Ruge-Stueben coarsening lives in the third-party AlgebraicMultigrid.jl (not under
/root/reference; version unpinned).  `ruge_stuben` below restates the
published algorithm (classical strength theta=0.25, first-pass RS C/F splitting, direct
interpolation, max_levels=10); the converged solution is hierarchy independent
(test/test_algebraic.jl:18-31).  Prolongator index structure: parity unpinned.

In addition to the level->fine prolongations R[X][l] that the reference keeps, the level->level
transfers T[X][l] (R[X][l] = R[X][l+1] @ T[X][l]) are retained: the B200 solver needs them for
the Galerkin coarse operators and the V-cycle (hard part 2).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .geometry import Geometry, _TFRef, _spectral1d_levels, boundary_node_ids, find_boundary

try:                                    # numba only accelerates the sequential C/F splitting
    import numba
    _njit = numba.njit(cache=True)
except Exception:                       # pragma: no cover
    numba = None
    _njit = lambda f: f

__all__ = ["MultiGrid", "AMG", "amg", "prepare_amg", "ruge_stuben", "amg_ruge_stuben", "amg_ruge_stuben_native"]


# --------------------------------------------------------------------------
# classical Ruge-Stueben
# --------------------------------------------------------------------------

@_njit
def _strength(n, Ap, Aj, Ax, theta):
    Sp = np.zeros(n + 1, np.int64)
    keep = np.zeros(Aj.shape[0], np.bool_)
    for i in range(n):
        m = 0.0
        for jj in range(Ap[i], Ap[i + 1]):
            if Aj[jj] != i and abs(Ax[jj]) > m:
                m = abs(Ax[jj])
        thr = theta * m
        c = 0
        for jj in range(Ap[i], Ap[i + 1]):
            if Aj[jj] == i or abs(Ax[jj]) >= thr:
                keep[jj] = True
                c += 1
        Sp[i + 1] = Sp[i] + c
    Sj = np.empty(Sp[n], np.int64)
    k = 0
    for jj in range(Aj.shape[0]):
        if keep[jj]:
            Sj[k] = Aj[jj]
            k += 1
    return Sp, Sj, keep


@_njit
def _rs_cf_splitting(n, Sp, Sj, Tp, Tj):
    U, Cpt, Fpt = 2, 1, 0
    lam = np.zeros(n, np.int64)
    for i in range(n):
        lam[i] = Tp[i + 1] - Tp[i]
    interval_ptr = np.zeros(n + 2, np.int64)
    interval_count = np.zeros(n + 2, np.int64)
    index_to_node = np.zeros(n, np.int64)
    node_to_index = np.zeros(n, np.int64)
    for i in range(n):
        interval_count[lam[i]] += 1
    cum = 0
    for i in range(n + 1):
        interval_ptr[i] = cum
        cum += interval_count[i]
        interval_count[i] = 0
    for i in range(n):
        li = lam[i]
        idx = interval_ptr[li] + interval_count[li]
        index_to_node[idx] = i
        node_to_index[i] = idx
        interval_count[li] += 1
    splitting = np.full(n, U, np.int64)
    for i in range(n):
        if lam[i] == 0 or (lam[i] == 1 and Tj[Tp[i]] == i):
            splitting[i] = Fpt
    for top in range(n - 1, -1, -1):
        i = index_to_node[top]
        li = lam[i]
        interval_count[li] -= 1
        if splitting[i] == Fpt:
            continue
        splitting[i] = Cpt
        for jj in range(Tp[i], Tp[i + 1]):
            j = Tj[jj]
            if splitting[j] == U:
                splitting[j] = Fpt
                for kk in range(Sp[j], Sp[j + 1]):
                    k = Sj[kk]
                    if splitting[k] == U:
                        if lam[k] >= n - 1:
                            continue
                        lk = lam[k]
                        old = node_to_index[k]
                        new = interval_ptr[lk] + interval_count[lk] - 1
                        node_to_index[index_to_node[old]] = new
                        node_to_index[index_to_node[new]] = old
                        tmp = index_to_node[old]
                        index_to_node[old] = index_to_node[new]
                        index_to_node[new] = tmp
                        interval_count[lk] -= 1
                        interval_count[lk + 1] += 1
                        interval_ptr[lk + 1] = new
                        lam[k] += 1
        for jj in range(Sp[i], Sp[i + 1]):
            j = Sj[jj]
            if splitting[j] == U:
                if lam[j] == 0:
                    continue
                lj = lam[j]
                old = node_to_index[j]
                new = interval_ptr[lj]
                node_to_index[index_to_node[old]] = new
                node_to_index[index_to_node[new]] = old
                tmp = index_to_node[old]
                index_to_node[old] = index_to_node[new]
                index_to_node[new] = tmp
                interval_count[lj] -= 1
                interval_count[lj - 1] += 1
                interval_ptr[lj] += 1
                interval_ptr[lj - 1] = interval_ptr[lj] - interval_count[lj - 1]
                lam[j] -= 1
    return splitting


@_njit
def _direct_interp(n, Ap, Aj, Ax, Sp, Sj, splitting):
    cmap = np.zeros(n, np.int64)
    nc = 0
    for i in range(n):
        if splitting[i] == 1:
            cmap[i] = nc
            nc += 1
    Pp = np.zeros(n + 1, np.int64)
    for i in range(n):
        if splitting[i] == 1:
            Pp[i + 1] = Pp[i] + 1
        else:
            c = 0
            for jj in range(Sp[i], Sp[i + 1]):
                if splitting[Sj[jj]] == 1 and Sj[jj] != i:
                    c += 1
            Pp[i + 1] = Pp[i] + c
    Pj = np.zeros(Pp[n], np.int64)
    Px = np.zeros(Pp[n], np.float64)
    strong = np.zeros(n, np.bool_)
    for i in range(n):
        if splitting[i] == 1:
            Pj[Pp[i]] = cmap[i]
            Px[Pp[i]] = 1.0
            continue
        for jj in range(Sp[i], Sp[i + 1]):
            strong[Sj[jj]] = True
        sum_strong_pos = 0.0
        sum_strong_neg = 0.0
        sum_all_pos = 0.0
        sum_all_neg = 0.0
        diag = 0.0
        for jj in range(Ap[i], Ap[i + 1]):
            j = Aj[jj]
            v = Ax[jj]
            if j == i:
                diag += v
            else:
                if v < 0:
                    sum_all_neg += v
                else:
                    sum_all_pos += v
                if strong[j] and splitting[j] == 1:
                    if v < 0:
                        sum_strong_neg += v
                    else:
                        sum_strong_pos += v
        alpha = sum_all_neg / sum_strong_neg if sum_strong_neg != 0.0 else 0.0
        beta = sum_all_pos / sum_strong_pos if sum_strong_pos != 0.0 else 0.0
        if sum_strong_pos == 0.0:
            diag += sum_all_pos
            beta = 0.0
        neg_coeff = -alpha / diag if diag != 0.0 else 0.0
        pos_coeff = -beta / diag if diag != 0.0 else 0.0
        nnz = Pp[i]
        for jj in range(Ap[i], Ap[i + 1]):
            j = Aj[jj]
            if j != i and strong[j] and splitting[j] == 1:
                v = Ax[jj]
                Pj[nnz] = cmap[j]
                Px[nnz] = (neg_coeff if v < 0 else pos_coeff) * v
                nnz += 1
        for jj in range(Sp[i], Sp[i + 1]):
            strong[Sj[jj]] = False
    return Pp, Pj, Px, nc


def ruge_stuben(A: sp.spmatrix, max_coarse=2, max_levels=10, theta=0.25) -> List[sp.csr_matrix]:
    """Prolongations finest -> coarsest of a classical RS hierarchy (amg_prolongators.jl:16-18
    call site; algorithm per the published AlgebraicMultigrid.jl / PyAMG description)."""
    A = sp.csr_matrix(A, dtype=np.float64)
    A.sort_indices()
    Ps = []
    while len(Ps) + 1 < max_levels and A.shape[0] > max_coarse:
        n = A.shape[0]
        Ap, Aj, Ax = A.indptr.astype(np.int64), A.indices.astype(np.int64), A.data
        Sp_, Sj_, _ = _strength(n, Ap, Aj, Ax, theta)
        S = sp.csr_matrix((np.ones(len(Sj_)), Sj_, Sp_), shape=(n, n))
        T = S.T.tocsr()
        T.sort_indices()
        splitting = _rs_cf_splitting(n, Sp_, Sj_, T.indptr.astype(np.int64), T.indices.astype(np.int64))
        Pp, Pj, Px, nc = _direct_interp(n, Ap, Aj, Ax, Sp_, Sj_, splitting)
        if nc == 0 or nc == n:
            break
        P = sp.csr_matrix((Px, Pj, Pp), shape=(n, nc))
        Ps.append(P)
        A = (P.T @ A @ P).tocsr()
        A.sort_indices()
    return Ps


def amg_ruge_stuben(**kw):
    """Prolongator factory (amg_prolongators.jl:16-18)."""
    kw.setdefault("max_coarse", 2)
    return lambda K: ruge_stuben(K, **kw)


def amg_ruge_stuben_native(**kw):
    """The same factory on the C++ implementation inside libmgbx (csrc/host_amg.hpp, mgbx_rs_*): bitwise the same prolongations
    (tests/test_abi_cpu.py), without numba."""
    kw.setdefault("max_coarse", 2)

    def f(K):
        from . import native
        return native.ruge_stuben(K, **kw)
    return f


# --------------------------------------------------------------------------
# MultiGrid container
# --------------------------------------------------------------------------

@dataclass
class MultiGrid:
    """Geometry + per-subspace level->fine prolongations R[X][l] (multigrid.jl:185-188) and the
    retained level->level transfers T[X][l] (l -> l+1)."""
    geometry: Geometry
    R: Dict[str, list]
    T: Dict[str, list]

    @property
    def L(self):
        return len(next(iter(self.R.values())))


def _is_dense(M):
    return isinstance(M, np.ndarray)


def _mm(A, B):
    return A @ B


def _block_diagonal_inverse(G, sizes=(2, 3, 4, 6, 7, 8)):
    """Inverse of a sparse matrix that is block diagonal with uniform blocks of one of `sizes`; None if it is not."""
    C = sp.coo_matrix(G)
    m = C.shape[0]
    for b in sizes:
        if m % b or not np.array_equal(C.row // b, C.col // b):
            continue
        blocks = np.zeros((m // b, b, b))
        np.add.at(blocks, (C.row // b, C.row % b, C.col % b), C.data)
        inv = np.linalg.inv(blocks)
        base = np.repeat(np.arange(m // b) * b, b * b)
        rows = base + np.tile(np.repeat(np.arange(b), b), m // b)
        cols = base + np.tile(np.tile(np.arange(b), b), m // b)
        return sp.csr_matrix((inv.reshape(-1), (rows, cols)), shape=(m, m))
    return None


def _transfer(s_next, r, s):
    """T with s_next @ T = r @ s (least squares; exact because the spaces are nested)."""
    rhs = r @ s
    if _is_dense(s_next):
        return np.linalg.lstsq(s_next, rhs, rcond=None)[0]
    s_next = sp.csc_matrix(s_next)
    G = (s_next.T @ s_next).tocsc()
    B = sp.csc_matrix(s_next.T @ rhs)
    d = G.diagonal()
    if (G - sp.diags(d)).nnz == 0 or abs(G - sp.diags(d)).sum() == 0:
        Tm = sp.diags(1.0 / d) @ B
    else:
        Ginv = _block_diagonal_inverse(G)
        # spsolve with a sparse right-hand side solves column by column (quadratic in the mesh size: 220 s of host set-up on a
        # 200 k-node P2 mesh); the Gram matrix of an element-local embedding (:broken_P1) is block diagonal and is inverted blockwise
        Tm = (Ginv @ B) if Ginv is not None else sp.csc_matrix(spla.spsolve(G, B))
    Tm = sp.csr_matrix(Tm)
    Tm.data[np.abs(Tm.data) < 1e-14] = 0.0
    Tm.eliminate_zeros()
    return Tm


def _stretch(refine: Dict[str, list], subspaces: Dict[str, list]):
    """multigrid.jl:226-265."""
    LX = {X: len(refine[X]) for X in refine}
    Lmax = max(LX.values())
    if all(v == Lmax for v in LX.values()):
        return refine, subspaces
    rs, ss = {}, {}
    for X in refine:
        Lx = LX[X]
        if Lx == Lmax:
            rs[X], ss[X] = refine[X], subspaces[X]
            continue
        s2n = [int(np.ceil(Lx * i / Lmax)) for i in range(1, Lmax + 1)]     # 1-based natural level
        rf, sb = [None] * Lmax, [None] * Lmax
        for i in range(Lmax):
            ni = s2n[i]
            sb[i] = subspaces[X][ni - 1]
            if i == Lmax - 1:
                rf[i] = refine[X][Lx - 1]
            elif s2n[i + 1] > ni:
                rf[i] = refine[X][ni - 1]
            else:
                m = sb[i].shape[0]
                rf[i] = np.eye(m) if _is_dense(sb[i]) else sp.identity(m, format="csr")
        rs[X], ss[X] = rf, sb
    return rs, ss


def _compose(geometry, subspaces, refine) -> MultiGrid:
    """multigrid.jl:192-217: R[X][l] = refine[L]...refine[l] @ sub[l]; also the transfers."""
    refine, subspaces = _stretch(refine, subspaces)
    R, T = {}, {}
    for X in subspaces:
        rX, sX = refine[X], subspaces[X]
        L = len(rX)
        rfp = [None] * L
        rfp[L - 1] = rX[L - 1]
        for l in range(L - 2, -1, -1):
            rfp[l] = rfp[l + 1] @ rX[l]
        Rl = [rfp[l] @ sX[l] for l in range(L)]
        if not _is_dense(Rl[0]):
            Rl = [sp.csr_matrix(M) for M in Rl]
        R[X] = Rl
        T[X] = [_transfer(sX[l + 1], rX[l], sX[l]) for l in range(L - 1)]
        for M in R[X] + T[X]:          # canonical CSR once, here: every consumer (block joins for the two AMGs) wants sorted rows
            if sp.issparse(M) and not M.has_canonical_format:
                M.sum_duplicates()
    return MultiGrid(geometry, R, T)


# --------------------------------------------------------------------------
# amg() for FEM discretizations
# --------------------------------------------------------------------------

def _corner_labels_from_t(t, corner_local):
    """multigrid.jl:137-151: compact corner ids by first occurrence, (corner, element) order."""
    flat = t[:, corner_local].reshape(-1)
    _, first, inv = np.unique(flat, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    return rank[inv.reshape(-1)].reshape(t.shape[0], len(corner_local)), order.size


def _p1_stiffness(corners, tri):
    """fem2d_P2.jl:646-669."""
    x1, y1 = corners[tri[:, 0], 0], corners[tri[:, 0], 1]
    x2, y2 = corners[tri[:, 1], 0], corners[tri[:, 1], 1]
    x3, y3 = corners[tri[:, 2], 0], corners[tri[:, 2], 1]
    det2 = (x2 - x1) * (y3 - y1) - (x3 - x1) * (y2 - y1)
    b = np.stack([y2 - y3, y3 - y1, y1 - y2], 1)
    c = np.stack([x3 - x2, x1 - x3, x2 - x1], 1)
    s = 1.0 / (2 * np.abs(det2))
    vals = (b[:, :, None] * b[:, None, :] + c[:, :, None] * c[:, None, :]) * s[:, None, None]
    rows = np.repeat(tri[:, :, None], 3, axis=2)
    cols = np.repeat(tri[:, None, :], 3, axis=1)
    nv = corners.shape[0]
    return sp.csr_matrix((vals.reshape(-1), (rows.reshape(-1), cols.reshape(-1))), shape=(nv, nv))


def _ladder(P_amg, bridge, n_doubled):
    """amg_prolongators.jl:48-66."""
    K_amg = len(P_amg) + 1
    L = K_amg + 1
    refine = [None] * L
    for i, P in enumerate(P_amg):
        refine[K_amg - 2 - i] = sp.csr_matrix(P)
    refine[K_amg - 1] = sp.csr_matrix(bridge)
    refine[L - 1] = sp.identity(n_doubled, format="csr")
    sizes = [0] * L
    sizes[K_amg - 1] = bridge.shape[1]
    for kk in range(K_amg - 2, -1, -1):
        sizes[kk] = refine[kk].shape[1]
    sizes[L - 1] = n_doubled
    return refine, sizes, L, K_amg


def _continuous_subspace(labels, n_unique, dirichlet_ids):
    """fem2d_P2.jl:331-346: 0/1 gluing matrix onto non-Dirichlet ids."""
    keep = np.ones(n_unique, bool)
    keep[np.asarray(list(dirichlet_ids), dtype=np.int64)] = False
    pos = np.full(n_unique, -1, np.int64)
    pos[keep] = np.arange(keep.sum())
    p = pos[labels]
    rows = np.nonzero(p >= 0)[0]
    return sp.csr_matrix((np.ones(rows.size), (rows, p[rows])), shape=(labels.size, int(keep.sum())))


def _corner_lift(geom: Geometry):
    """Per-element lift of corner values to the broken basis: (local weights V x nc, corner slots)."""
    if geom.kind == "p1":
        return np.eye(3), np.array([0, 1, 2])
    if geom.kind == "p2":
        Lw = np.zeros((geom.V, 3))
        Lw[0, 0] = Lw[2, 1] = Lw[4, 2] = 1.0
        Lw[1, [0, 1]] = Lw[3, [1, 2]] = Lw[5, [2, 0]] = 0.5
        if geom.V == 7:
            Lw[6, :] = 1.0 / 3
        return Lw, np.array([0, 2, 4])
    ref = _TFRef(geom.dim, geom.k)
    return ref.q1_lift(), ref.corner_local()


def _bridge(geom, corner_conn, n_v, interior, Lw):
    """Interior corners -> broken basis (fem2d_P1.jl:211-231, fem2d_P2.jl:675-708,
    TensorFEM.jl:686-712)."""
    N, V = geom.N, geom.V
    nc = corner_conn.shape[1]
    idx = np.full(n_v, -1, np.int64)
    idx[interior] = np.arange(len(interior))
    cui = idx[corner_conn]                                   # (N, nc)
    rows = (np.arange(N)[:, None, None] * V + np.arange(V)[None, :, None]) + np.zeros((1, 1, nc), np.int64)
    cols = np.broadcast_to(cui[:, None, :], (N, V, nc))
    vals = np.broadcast_to(Lw[None, :, :], (N, V, nc))
    m = (cols >= 0) & (vals != 0)
    B = sp.csr_matrix((vals[m], (rows[m], cols[m])), shape=(N * V, len(interior)))
    B.sum_duplicates()
    return B


def _broken_op_sparse(blocks):
    N, V, _ = blocks.shape
    return sp.bsr_matrix((blocks, np.arange(N), np.arange(N + 1)), shape=(N * V, N * V)).tocsr()


def _broken_p1_embedding(N, V):
    """fem2d_P2.jl:355-380."""
    slot = np.array([[1, -1, 1], [1, 0, 0], [1, 1, -1], [0, 1, 0], [-1, 1, 1], [0, 0, 1]], float)
    if V == 7:
        slot = np.vstack([slot, [1 / 3, 1 / 3, 1 / 3]])
    return sp.kron(sp.identity(N), sp.csr_matrix(slot), format="csr")


def amg(geom: Geometry, prolongator=None, dirichlet_nodes: Optional[Dict[str, Sequence]] = None
        ) -> MultiGrid:
    """Attach an algebraic-multigrid hierarchy (multigrid.jl:290-356 and the per-discretization
    methods).  dirichlet_nodes: name -> list of (v, e) pairs (0-based)."""
    if geom.dense:
        return _amg_spectral(geom)
    if prolongator is None:
        prolongator = amg_ruge_stuben(max_coarse=2)
    N, V = geom.N, geom.V
    n_doubled = N * V
    labels = geom.t.reshape(-1)
    n_unique = int(labels.max()) + 1
    Lw, cslots = _corner_lift(geom)
    cconn, n_v = _corner_labels_from_t(geom.t, cslots)
    xf = geom.xflat()
    if geom.kind in ("p1", "p2"):
        corners = np.zeros((n_v, 2))
        flatc = cconn.reshape(-1)
        src = (np.arange(N)[:, None] * V + cslots[None, :]).reshape(-1)
        _, first = np.unique(flatc, return_index=True)
        corners[flatc[first]] = xf[src[first]]
        K_full = _p1_stiffness(corners, cconn)
    else:
        W = sp.diags(geom.w)
        A = None
        for name in ("dx", "dy", "dz")[: geom.x.shape[2]]:
            Da = _broken_op_sparse(geom.operators[name])
            A = Da.T @ W @ Da if A is None else A + Da.T @ W @ Da
        S = _bridge(geom, cconn, n_v, np.arange(n_v), Lw)
        K_full = (S.T @ A @ S).tocsr()
    full_to_corner = np.full(n_unique, -1, np.int64)
    full_to_corner[geom.t[:, cslots].reshape(-1)] = cconn.reshape(-1)

    def hierarchy(interior):
        interior = np.asarray(interior, np.int64)
        if interior.size == 0:
            Ps = []
        else:
            Ps = prolongator(K_full[interior][:, interior].tocsr())
        return _ladder(Ps, _bridge(geom, cconn, n_v, interior, Lw), n_doubled)

    refine_full, sizes_full, L_full, K_amg_full = hierarchy(np.arange(n_v))
    ident = lambda m: sp.identity(m, format="csr")
    ones = lambda m: sp.csr_matrix(np.ones((m, 1)))
    sub_full = [ident(sizes_full[k]) for k in range(K_amg_full)] + [ident(n_doubled)]
    sub_uni = [ones(sizes_full[k]) for k in range(K_amg_full)] + [ones(n_doubled)]
    subspaces = {"full": sub_full, "uniform": sub_uni}
    refine = {"full": refine_full, "uniform": refine_full}
    if geom.kind == "p2":
        subspaces["broken_P1"] = [ident(sizes_full[k]) for k in range(K_amg_full)] + \
            [_broken_p1_embedding(N, V)]
        refine["broken_P1"] = refine_full
    if dirichlet_nodes is None:
        dirichlet_nodes = {"dirichlet": find_boundary(geom)}
    for sym, nodes in dirichlet_nodes.items():
        if sym in subspaces:
            raise ValueError("dirichlet_nodes key :%s is reserved" % sym)
        lin = np.array([v + e * V for (v, e) in nodes], dtype=np.int64)
        dd = np.unique(labels[lin]) if lin.size else np.zeros(0, np.int64)
        dc = np.unique(full_to_corner[dd][full_to_corner[dd] >= 0]) if dd.size else dd
        interior = np.setdiff1d(np.arange(n_v), dc)
        r, sizes, Ld, Kd = hierarchy(interior)
        if geom.kind != "p1":                       # mask Dirichlet rows of the bridge
            keep = np.ones(n_unique)
            keep[dd] = 0.0
            Bm = sp.diags(keep[labels]) @ r[Kd - 1]
            Bm = sp.csr_matrix(Bm)
            Bm.eliminate_zeros()
            r[Kd - 1] = Bm
            sub_fine = _continuous_subspace(labels, n_unique, dd)
        else:
            sub_fine = r[Kd - 1]
        subspaces[sym] = [ident(sizes[k]) for k in range(Kd)] + [sub_fine]
        refine[sym] = r
    return _compose(geom, subspaces, refine)


def _amg_spectral(geom: Geometry) -> MultiGrid:
    lv = _spectral1d_levels(geom.k)
    mg1 = _compose(geom, lv["subspaces"], {X: lv["refine"] for X in lv["subspaces"]})
    if geom.kind == "spectral1d":
        return mg1
    R = {X: [np.kron(M, M) for M in mg1.R[X]] for X in mg1.R}
    T = {X: [np.kron(M, M) for M in mg1.T[X]] for X in mg1.T}
    return MultiGrid(geom, R, T)


# --------------------------------------------------------------------------
# AMG pair (main, feasibility)
# --------------------------------------------------------------------------

@dataclass
class AMG:
    """Pure-data hierarchy consumed by the solver (multigrid.jl:278-288).

    R_fine[l]: (nu*n, m_l) block-diagonal join of the per-variable prolongations;
    T[l]: (m_{l+1}, m_l) level transfers; D: list of (var index, operator name);
    var_offsets[l][k]: first column of state variable k at level l."""
    geometry: Geometry
    w: np.ndarray
    R_fine: list
    T: list
    D: list
    state_variables: list
    var_offsets: list
    dense: bool

    @property
    def nu(self):
        return len(self.state_variables)

    @property
    def nD(self):
        return len(self.D)

    @property
    def L(self):
        return len(self.R_fine)


def _blockdiag(mats):
    if _is_dense(mats[0]):
        import scipy.linalg as sl
        return sl.block_diag(*mats)
    # direct CSR concatenation (what sp.block_diag(mats, format="csr") returns, without its COO round trip and sort:
    # 3.6 of the 17 s of host set-up on a 40^3 hexahedral mesh)
    indptr, indices, data = [np.zeros(1, np.int64)], [], []
    rows = cols = nnz = 0
    for M in mats:
        M = sp.csr_matrix(M)
        if not M.has_canonical_format:
            M = M.copy()
            M.sum_duplicates()
        indptr.append(M.indptr[1:].astype(np.int64) + nnz)
        indices.append(M.indices.astype(np.int64) + cols)
        data.append(np.asarray(M.data, dtype=float))
        rows += M.shape[0]
        cols += M.shape[1]
        nnz += M.nnz
    idt = np.int32 if max(rows, cols, nnz) < 2 ** 31 - 1 else np.int64
    out = sp.csr_matrix((np.concatenate(data), np.concatenate(indices).astype(idt), np.concatenate(indptr).astype(idt)), shape=(rows, cols))
    out.has_canonical_format = True
    return out


def amg_helper(mg: MultiGrid, state_variables, D) -> AMG:
    """multigrid.jl:474-512."""
    geom = mg.geometry
    sv = [tuple(r) for r in state_variables]
    nu = len(sv)
    L = len(mg.R[sv[0][1]])
    R_fine = [_blockdiag([mg.R[sv[k][1]][l] for k in range(nu)]) for l in range(L)]
    T = [_blockdiag([mg.T[sv[k][1]][l] for k in range(nu)]) for l in range(L - 1)]
    offs = []
    for l in range(L):
        o = [0]
        for k in range(nu):
            o.append(o[-1] + mg.R[sv[k][1]][l].shape[1])
        offs.append(o)
    bar = {sv[k][0]: k for k in range(nu)}
    Dl = []
    for (var, op) in D:
        if var not in bar:
            raise ValueError("D references state variable :%s, not in state_variables" % var)
        if op not in geom.operators:
            raise ValueError("D references operator :%s; available: %s" % (op, list(geom.operators)))
        Dl.append((bar[var], op))
    return AMG(geometry=geom, w=geom.w, R_fine=R_fine, T=T, D=Dl, state_variables=sv,
               var_offsets=offs, dense=geom.dense)


def prepare_amg(mg: MultiGrid, state_variables, D):
    """(main, feasibility) AMG pair (multigrid.jl:515-538)."""
    sv = [tuple(r) for r in state_variables]
    Dm = [tuple(r) for r in D]
    M1 = amg_helper(mg, sv, Dm)
    s1 = sv + [("feasibility_slack", "full")]
    D1 = Dm + [("feasibility_slack", "id")] + [(v[0], "id") for v in sv]
    M2 = amg_helper(mg, s1, D1)
    return M1, M2
