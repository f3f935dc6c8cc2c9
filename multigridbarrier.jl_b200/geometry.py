"""Host-side (CPU, setup-time) geometry builders.

These produce the *inputs* of the hot path (operators as p x p x N element blocks,
quadrature weights, connectivity) with the same layouts and numerical conventions as
the reference; they are not on the accelerated path ("setup OUT OF
SCOPE as code to accelerate, must be restated to generate inputs").

Reference behaviour restated here (files under /root/reference/src):
  TensorFEM.jl:117-263   reference element (Chebyshev-Lobatto nodes, Clenshaw-Curtis weights)
  TensorFEM.jl:338-383   tensor_dofmap (topological numbering)
  TensorFEM.jl:428-490   isoparametric operator build
  TensorFEM.jl:643-678   find_boundary (face use count)
  TensorFEM.jl:821-954   geometric subdivision
  fem2d_P1.jl:131-308    P1 triangles, refinement, operators
  fem2d_P2.jl:74-154     P2(+bubble) reference triangle
  fem2d_P2.jl:169-207    P2 connectivity refinement
  fem2d_P2.jl:468-596    isoparametric P2 operators
  spectral1d.jl:63-109   Chebyshev hierarchy, spectral2d.jl:15-42 Kronecker lift

Array conventions (numpy, 0-based):
  x[e, v, d]      node coordinates  (reference: x[v, e, d])
  t[e, v]         global node id of local node v of element e (0-based)
  ops[name][e, r, c]  element block, row r / column c  (reference data[r, c, e])
  flat node index i = e * V + v  (reference v + (e-1)V)
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np

__all__ = [
    "Geometry", "fem1d", "fem2d", "fem3d", "fem2d_P1", "fem2d_P2", "spectral1d",
    "spectral2d", "subdivide", "find_boundary", "tensor_dofmap", "structured_box",
    "structured_triangles",
]


@dataclass
class Geometry:
    """Single-level discretization (reference: src/multigrid.jl:37-43)."""
    kind: str                      # 'tensor' | 'p1' | 'p2' | 'spectral1d' | 'spectral2d'
    x: np.ndarray                  # (N, V, D)
    w: np.ndarray                  # (N*V,)
    t: np.ndarray                  # (N, V) int64, 0-based
    operators: Dict[str, np.ndarray]   # name -> (N, V, V) blocks; spectral: N == 1 (dense)
    dim: int                       # intrinsic dimension
    k: int = 1                     # polynomial order (tensor) / spectral n
    bubble: bool = True            # P2 only
    dense: bool = False            # spectral geometries (single dense "element")
    extra: dict = field(default_factory=dict)

    @property
    def N(self):
        return self.x.shape[0]

    @property
    def V(self):
        return self.x.shape[1]

    @property
    def n(self):
        return self.x.shape[0] * self.x.shape[1]

    def xflat(self):
        return self.x.reshape(self.n, self.x.shape[2])


# --------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------

def _dedupe(xf: np.ndarray):
    """Coordinate dedup -> labels by first occurrence (TensorFEM.jl:74-110 gives the same
    *partition*; its label order depends on Julia's hash/RNG and is not reproducible
    outside Julia)."""
    tol = max(np.abs(xf).max(), 1.0) * 100 * np.finfo(xf.dtype).eps
    key = np.round(xf / (16 * tol)).astype(np.int64)
    _, first, inv = np.unique(key, axis=0, return_index=True, return_inverse=True)
    inv = inv.reshape(-1)
    order = np.argsort(first, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    return rank[inv]


def _first_occurrence_ids(keys: np.ndarray, start: int):
    """Number distinct rows of `keys` (in traversal order) by first occurrence, from `start`."""
    if keys.ndim == 1:
        _, first, inv = np.unique(keys, return_index=True, return_inverse=True)
    else:
        _, first, inv = np.unique(keys, axis=0, return_index=True, return_inverse=True)
    inv = inv.reshape(-1)
    order = np.argsort(first, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    return start + rank[inv], order.size


# --------------------------------------------------------------------------
# Tensor-product Q_k elements (fem1d / fem2d / fem3d)
# --------------------------------------------------------------------------

def _tf_nodes(k):
    return np.array([-np.cos(np.pi * i / k) for i in range(k + 1)]) if k > 0 else np.array([0.0])


def _tf_weights(k):
    if k == 0:
        return np.array([2.0])
    N = k
    w = np.zeros(N + 1)
    for i in range(N + 1):
        val = 1.0
        for j in range(1, N // 2 + 1):
            c = 1.0 if 2 * j == N else 2.0
            val += c / (1 - 4.0 * j * j) * np.cos(np.pi * (2 * j * i) / N)
        w[i] = val / N if (i == 0 or i == N) else 2 * val / N
    return w


def _tf_dmat(nodes):
    s = len(nodes)
    D = np.zeros((s, s))
    for i in range(s):
        for j in range(s):
            if i == j:
                D[i, j] = sum(1.0 / (nodes[i] - nodes[m]) for m in range(s) if m != i)
            else:
                num = 1.0
                for m in range(s):
                    if m != j and m != i:
                        num *= nodes[i] - nodes[m]
                den = 1.0
                for m in range(s):
                    if m != j:
                        den *= nodes[j] - nodes[m]
                D[i, j] = num / den
    return D


def _tf_lagrange(nodes, xv):
    s = len(nodes)
    vals = np.ones(s)
    for i in range(s):
        for j in range(s):
            if i != j:
                vals[i] *= (xv - nodes[j]) / (nodes[i] - nodes[j])
    return vals


class _TFRef:
    def __init__(self, d, k):
        s = k + 1
        self.d, self.k, self.s, self.n = d, k, s, s ** d
        self.nodes1 = _tf_nodes(k)
        self.w1 = _tf_weights(k)
        self.D1 = _tf_dmat(self.nodes1)
        I1 = np.eye(s)
        self.Daxis = []
        for axis in range(1, d + 1):
            M = np.ones((1, 1))
            for j in range(1, d + 1):           # factors for axes d, d-1, ..., 1
                M = np.kron(M, self.D1 if (d - j + 1) == axis else I1)
            self.Daxis.append(M)
        idx = np.indices((s,) * d).reshape(d, -1)      # C-order: last axis fastest
        # want axis 1 fastest: multi-index of lin = i1 + s*i2 + ...
        lin = np.arange(self.n)
        mi = np.zeros((self.n, d), dtype=np.int64)
        r = lin.copy()
        for a in range(d):
            mi[:, a] = r % s
            r //= s
        self.mi = mi
        self.nodesref = self.nodes1[mi]
        self.wref = np.prod(self.w1[mi], axis=1)

    def q1_lift(self):
        d, n = self.d, self.n
        nc = 1 << d
        L = np.ones((n, nc))
        for c in range(nc):
            for a in range(d):
                xi = self.nodesref[:, a]
                bit = (c >> a) & 1
                L[:, c] *= (1 - xi) / 2 if bit == 0 else (1 + xi) / 2
        return L

    def corner_local(self):
        d, s = self.d, self.s
        out = []
        for c in range(1 << d):
            lin, stride = 0, 1
            for a in range(d):
                ia = 0 if ((c >> a) & 1) == 0 else s - 1
                lin += ia * stride
                stride *= s
            out.append(lin)
        return np.array(out, dtype=np.int64)


def _tf_face_pos(ids, pi, pj, k):
    g = lambda i, j: ids[i + 2 * j]
    i0 = j0 = 0
    best = g(0, 0)
    for j in (0, 1):
        for i in (0, 1):
            if g(i, j) < best:
                best, i0, j0 = g(i, j), i, j
    ri = pi if i0 == 0 else k - pi
    rj = pj if j0 == 0 else k - pj
    if g(1 - i0, j0) > g(i0, 1 - j0):
        ri, rj = rj, ri
    return ri + rj * (k + 1)


def _entity_corner_ids(cor, mi, inter, s, d):
    nint = len(inter)
    out = []
    for combo in range(1 << nint):
        cbits = 0
        for a in range(d):
            if a in inter:
                bit = (combo >> inter.index(a)) & 1
            else:
                bit = 1 if mi[a] == s - 1 else 0
            cbits |= bit << a
        out.append(int(cor[cbits]))
    return out


def tensor_dofmap(t_corner: np.ndarray, k: int, d: int) -> np.ndarray:
    """Full-node connectivity from corner connectivity (TensorFEM.jl:338-383).
    t_corner: (N, 2^d), 0-based ids. Returns (N, (k+1)^d), 0-based."""
    ref = _TFRef(d, k)
    s, n = ref.s, ref.n
    N = t_corner.shape[0]
    if k == 1:
        return np.asarray(t_corner, dtype=np.int64)[:, :].copy()  # nodes == corners (same order)
    t = np.empty((N, n), dtype=np.int64)
    next_id = int(t_corner.max()) + 1 if t_corner.size else 0
    reg = {}
    mis = [tuple(int(v) for v in ref.mi[v_]) for v_ in range(n)]
    inters = [[a for a in range(d) if mi[a] != 0 and mi[a] != s - 1] for mi in mis]
    for e in range(N):
        cor = t_corner[e]
        for v in range(n):
            mi, inter = mis[v], inters[v]
            nint = len(inter)
            if nint == d:
                t[e, v] = next_id
                next_id += 1
                continue
            ids = _entity_corner_ids(cor, mi, inter, s, d)
            if nint == 0:
                t[e, v] = ids[0]
                continue
            if nint == 1:
                p = mi[inter[0]]
                pos = p if ids[0] <= ids[1] else k - p
                key = (tuple(sorted(ids)), pos)
            elif nint == 2:
                pos = _tf_face_pos(ids, mi[inter[0]], mi[inter[1]], k)
                key = (tuple(sorted(ids)), pos)
            else:
                raise ValueError("tensor_dofmap: shared entities of dimension >= 3 unsupported")
            idv = reg.get(key)
            if idv is None:
                idv = next_id
                next_id += 1
                reg[key] = idv
            t[e, v] = idv
    return t


def _tf_build_geometry(d, k, x, t=None):
    """Isoparametric operator build (TensorFEM.jl:428-490). x: (N, n, e)."""
    ref = _TFRef(d, k)
    N, n, e = x.shape
    assert n == ref.n
    # grefs[b][N, i, dim] = sum_m Daxis[b][i, m] x[N, m, dim]
    J = np.stack([np.einsum("im,nmc->nic", ref.Daxis[b], x) for b in range(d)], axis=-1)  # (N,n,e,d)
    g = np.einsum("nicb,nicd->nibd", J, J)                 # first fundamental form (N,n,d,d)
    detg = np.linalg.det(g)
    P = np.linalg.solve(g, np.swapaxes(J, -1, -2))        # (N,n,d,e): P[b,dim]
    ops = {"id": np.broadcast_to(np.eye(n), (N, n, n)).copy()}
    Dax = np.stack(ref.Daxis, axis=0)                      # (d, n, n)
    names = ("dx", "dy", "dz")
    for dim in range(e):
        ops[names[dim]] = np.einsum("nib,bim->nim", P[:, :, :, dim], Dax)
    w = (ref.wref[None, :] * np.sqrt(np.maximum(detg, 0.0))).reshape(-1)
    if not np.all(w > 0):
        raise ValueError("fem%dd: non-positive quadrature weight (degenerate element map)" % d)
    if t is None:
        t = _dedupe(x.reshape(N * n, e)).reshape(N, n)
    return Geometry(kind="tensor", x=x, w=w, t=np.asarray(t, dtype=np.int64), operators=ops,
                    dim=d, k=k)


def _tf_resolve_mesh(K, k, d):
    ref = _TFRef(d, k)
    if K.shape[1] == ref.n:
        return K
    if K.shape[1] == (1 << d):
        return np.einsum("ic,ncd->nid", ref.q1_lift(), K)
    raise ValueError("mesh needs 2^d corners or (k+1)^d nodes per element")


def fem1d(nodes=(-1.0, 1.0), k=1, K=None, t=None):
    """1-D Q_k geometry (TensorFEM.jl:555-562)."""
    if K is None:
        nodes = np.asarray(nodes, dtype=float)
        K = np.stack([nodes[:-1], nodes[1:]], axis=1)[:, :, None]
    return _tf_build_geometry(1, k, _tf_resolve_mesh(np.asarray(K, float), k, 1), t)


def fem2d(k=1, K=None, t=None):
    """2-D Q_k quadrilateral geometry (TensorFEM.jl:589-595)."""
    if K is None:
        K = np.array([[[-1, -1], [1, -1], [-1, 1], [1, 1]]], dtype=float)
    return _tf_build_geometry(2, k, _tf_resolve_mesh(np.asarray(K, float), k, 2), t)


def fem3d(k=3, K=None, t=None):
    """3-D Q_k hexahedral geometry (TensorFEM.jl:624-630)."""
    if K is None:
        K = np.array([[[-1, -1, -1], [1, -1, -1], [-1, 1, -1], [1, 1, -1],
                       [-1, -1, 1], [1, -1, 1], [-1, 1, 1], [1, 1, 1]]], dtype=float)
    return _tf_build_geometry(3, k, _tf_resolve_mesh(np.asarray(K, float), k, 3), t)


def structured_box(d, cells, lo=-1.0, hi=1.0, k=1):
    """Vectorised regular grid of `cells`^d Q_k elements on [lo,hi]^d with lattice connectivity.
    Used for the large synthetic meshes of BASELINE.json (e.g. fem3d k=1, 100^3 hexes)."""
    m = cells
    h = (hi - lo) / m
    idx = np.indices((m,) * d).reshape(d, -1)[::-1]       # axis 1 fastest element order
    nc = 1 << d
    K = np.empty((m ** d, nc, d))
    tcorner = np.empty((m ** d, nc), dtype=np.int64)
    for c in range(nc):
        gid = np.zeros(m ** d, dtype=np.int64)
        stride = 1
        for a in range(d):
            bit = (c >> a) & 1
            K[:, c, a] = lo + h * (idx[a] + bit)
            gid += (idx[a] + bit) * stride
            stride *= m + 1
        tcorner[:, c] = gid
    x = _tf_resolve_mesh(K, k, d)
    t = tcorner if k == 1 else tensor_dofmap(tcorner, k, d)
    return _tf_build_geometry(d, k, x, t)


def _tf_refine_connectivity(t, k, d):
    """TensorFEM.jl:821-860: child corner ids from parent entities, then tensor_dofmap."""
    ref = _TFRef(d, k)
    nc = 1 << d
    cl = ref.corner_local()
    N = t.shape[0]
    child = np.empty((N * nc, nc), dtype=np.int64)
    ids = {}
    for e in range(N):
        pc = [int(t[e, cl[c]]) for c in range(nc)]
        for ch in range(nc):
            for c in range(nc):
                mi = tuple(((ch >> a) & 1) + ((c >> a) & 1) for a in range(d))   # 0,1,2 grid
                inter = [a for a in range(d) if mi[a] == 1]
                ent = _entity_corner_ids(pc, mi, inter, 3, d)
                if not inter:
                    key = (0, ent[0])
                elif len(inter) == d:
                    key = (-1, e)
                else:
                    key = (len(inter),) + tuple(sorted(ent))
                v = ids.get(key)
                if v is None:
                    v = len(ids)
                    ids[key] = v
                child[e * nc + ch, c] = v
    return tensor_dofmap(child, k, d)


def _tf_refine_local(k, d):
    ref = _TFRef(d, k)
    s, n = ref.s, ref.n
    nc = 1 << d
    P = np.zeros((nc * n, n))
    for ch in range(nc):
        lag = []
        for a in range(d):
            shift = -0.5 if ((ch >> a) & 1) == 0 else 0.5
            cn = ref.nodes1 * 0.5 + shift
            lag.append(np.array([_tf_lagrange(ref.nodes1, v) for v in cn]))   # (s child, s parent)
        for i in range(n):
            for j in range(n):
                wv = 1.0
                for a in range(d):
                    wv *= lag[a][ref.mi[i, a], ref.mi[j, a]]
                P[ch * n + i, j] = wv
    return P


# --------------------------------------------------------------------------
# P1 triangles
# --------------------------------------------------------------------------

_DEFAULT_TRI = np.array([[[-1, -1], [1, -1], [-1, 1]], [[1, -1], [1, 1], [-1, 1]]], dtype=float)


def _p1_operators(x):
    """fem2d_P1.jl:279-308. x: (N,3,2)."""
    N = x.shape[0]
    x1, y1 = x[:, 0, 0], x[:, 0, 1]
    x2, y2 = x[:, 1, 0], x[:, 1, 1]
    x3, y3 = x[:, 2, 0], x[:, 2, 1]
    det2 = (x2 - x1) * (y3 - y1) - (x3 - x1) * (y2 - y1)
    area = np.abs(det2) / 2
    b = np.stack([y2 - y3, y3 - y1, y1 - y2], axis=1) / det2[:, None]
    c = np.stack([x3 - x2, x1 - x3, x2 - x1], axis=1) / det2[:, None]
    dx = np.repeat(b[:, None, :], 3, axis=1)
    dy = np.repeat(c[:, None, :], 3, axis=1)
    w = np.repeat(area / 3, 3)
    ops = {"id": np.broadcast_to(np.eye(3), (N, 3, 3)).copy(), "dx": dx, "dy": dy}
    return ops, w


def _refine_p1_connectivity(t):
    """fem2d_P1.jl:251-277 (vectorised; edge ids by first occurrence in (ab, bc, ca) order)."""
    N = t.shape[0]
    a, b, c = t[:, 0], t[:, 1], t[:, 2]
    ends = np.stack([np.stack([a, b], 1), np.stack([b, c], 1), np.stack([c, a], 1)], axis=1)  # (N,3,2)
    keys = np.sort(ends.reshape(-1, 2), axis=1)
    nv = int(t.max()) + 1
    mids, _ = _first_occurrence_ids(keys[:, 0] * (nv + 1) + keys[:, 1], nv)
    mids = mids.reshape(N, 3)
    ab, bc, ca = mids[:, 0], mids[:, 1], mids[:, 2]
    out = np.empty((N, 4, 3), dtype=np.int64)
    out[:, 0] = np.stack([a, ab, ca], 1)
    out[:, 1] = np.stack([ab, b, bc], 1)
    out[:, 2] = np.stack([ca, bc, c], 1)
    out[:, 3] = np.stack([ab, bc, ca], 1)
    return out.reshape(4 * N, 3)


_P1_REFINE = np.array([
    [1, 0, 0], [.5, .5, 0], [.5, 0, .5],
    [.5, .5, 0], [0, 1, 0], [0, .5, .5],
    [.5, 0, .5], [0, .5, .5], [0, 0, 1],
    [.5, .5, 0], [0, .5, .5], [.5, 0, .5]])


def fem2d_P1(K=None, t=None):
    """P1 triangles (fem2d_P1.jl:39-45). K: (N,3,2)."""
    x = _DEFAULT_TRI.copy() if K is None else np.asarray(K, float)
    if t is None:
        t = _dedupe(x.reshape(-1, 2)).reshape(-1, 3)
    ops, w = _p1_operators(x)
    return Geometry(kind="p1", x=x, w=w, t=np.asarray(t, np.int64), operators=ops, dim=2, k=1)


def structured_triangles(m, lo=-1.0, hi=1.0):
    """m x m squares on [lo,hi]^2, each split into 2 triangles (synthetic ~1M-DOF P1 meshes)."""
    h = (hi - lo) / m
    j, i = np.indices((m, m)).reshape(2, -1)
    gid = lambda ii, jj: ii + (m + 1) * jj
    P = lambda ii, jj: np.stack([lo + h * ii, lo + h * jj], axis=1)
    x = np.empty((2 * m * m, 3, 2))
    t = np.empty((2 * m * m, 3), dtype=np.int64)
    x[0::2, 0], x[0::2, 1], x[0::2, 2] = P(i, j), P(i + 1, j), P(i, j + 1)
    x[1::2, 0], x[1::2, 1], x[1::2, 2] = P(i + 1, j), P(i + 1, j + 1), P(i, j + 1)
    t[0::2] = np.stack([gid(i, j), gid(i + 1, j), gid(i, j + 1)], 1)
    t[1::2] = np.stack([gid(i + 1, j), gid(i + 1, j + 1), gid(i, j + 1)], 1)
    return fem2d_P1(K=x, t=t)


# --------------------------------------------------------------------------
# P2 (+bubble) triangles
# --------------------------------------------------------------------------

def _p2_basis(lam, bubble):
    """Nodal basis values and d/dlambda gradients at barycentric points lam (m,3).
    Node order: c1, e12, c2, e23, c3, e31 [, centroid] (fem2d_P2.jl:19-20)."""
    l1, l2, l3 = lam[:, 0], lam[:, 1], lam[:, 2]
    m = lam.shape[0]
    V = 7 if bubble else 6
    phi = np.zeros((m, V))
    g = np.zeros((m, V, 3))
    corners = [(0, 0), (2, 1), (4, 2)]
    for slot, a in corners:
        la = lam[:, a]
        phi[:, slot] = la * (2 * la - 1)
        g[:, slot, a] = 4 * la - 1
    edges = [(1, 0, 1), (3, 1, 2), (5, 2, 0)]
    for slot, a, b in edges:
        phi[:, slot] = 4 * lam[:, a] * lam[:, b]
        g[:, slot, a] = 4 * lam[:, b]
        g[:, slot, b] = 4 * lam[:, a]
    if bubble:
        bub = 27 * l1 * l2 * l3
        gb = 27 * np.stack([l2 * l3, l1 * l3, l1 * l2], axis=1)
        cvals = [-1.0 / 9, 4.0 / 9, -1.0 / 9, 4.0 / 9, -1.0 / 9, 4.0 / 9]   # P2 basis at centroid
        for slot in range(6):
            phi[:, slot] -= cvals[slot] * bub
            g[:, slot, :] -= cvals[slot] * gb
        phi[:, 6] = bub
        g[:, 6, :] = gb
    return phi, g


def reference_triangle(bubble=True):
    """Reference tables K (node barycentrics), w, dx, dy with xi = lambda1, eta = lambda2
    (fem2d_P2.jl:74-154; the numeric tables there are reproduced by this construction and
    checked entry-by-entry in tests/test_geometry.py)."""
    lam = np.array([[1, 0, 0], [.5, .5, 0], [0, 1, 0], [0, .5, .5], [0, 0, 1], [.5, 0, .5]], float)
    if bubble:
        lam = np.vstack([lam, [1 / 3, 1 / 3, 1 / 3]])
        w = np.array([3, 8, 3, 8, 3, 8, 27], float) / 60
    else:
        w = np.array([0, 1, 0, 1, 0, 1], float) / 3
    phi, g = _p2_basis(lam, bubble)
    dx = g[:, :, 0] - g[:, :, 2]       # d/dxi  (lambda3 = 1 - xi - eta)
    dy = g[:, :, 1] - g[:, :, 2]       # d/deta
    return dict(K=lam, w=w, dx=dx, dy=dy)


def _p2_child_lams():
    """Barycentric (parent) coordinates of the child corners, in the reference's four-child
    order (fem2d_P2.jl:183-184): (ca,a,ab), (ab,b,bc), (bc,c,ca), (ab,bc,ca)."""
    a, b, c = np.eye(3)
    ab, bc, ca = (a + b) / 2, (b + c) / 2, (c + a) / 2
    return [(ca, a, ab), (ab, b, bc), (bc, c, ca), (ab, bc, ca)]


def _p2_refine_matrix(bubble):
    """(4V x V) interpolation of the parent element at the child nodes."""
    V = 7 if bubble else 6
    rows = []
    for (c1, c2, c3) in _p2_child_lams():
        pts = [c1, (c1 + c2) / 2, c2, (c2 + c3) / 2, c3, (c3 + c1) / 2]
        if bubble:
            pts.append((c1 + c2 + c3) / 3)
        phi, _ = _p2_basis(np.array(pts), bubble)
        rows.append(phi)
    return np.vstack(rows)


def _refine_p2_connectivity(t):
    """fem2d_P2.jl:169-207."""
    N, V = t.shape
    node_ids = {}
    for e in range(N):
        for v in range(6):
            node_ids.setdefault(int(t[e, v]), len(node_ids))
    edge_nodes = {}
    next_id = len(node_ids)
    out = np.empty((4 * N, V), dtype=np.int64)
    for e in range(N):
        a, ab, b, bc, c, ca = (node_ids[int(t[e, v])] for v in range(6))
        for s_, corners in enumerate(((ca, a, ab), (ab, b, bc), (bc, c, ca), (ab, bc, ca))):
            j = 4 * e + s_
            out[j, 0], out[j, 2], out[j, 4] = corners
            for slot, u, v in ((1, corners[0], corners[1]), (3, corners[1], corners[2]),
                               (5, corners[2], corners[0])):
                key = (u, v) if u < v else (v, u)
                eid = edge_nodes.get(key)
                if eid is None:
                    eid = next_id
                    next_id += 1
                    edge_nodes[key] = eid
                out[j, slot] = eid
            if V == 7:
                out[j, 6] = next_id
                next_id += 1
    return out


def _p2_operators(x, bubble):
    """fem2d_P2.jl:539-555 (isoparametric)."""
    R = reference_triangle(bubble)
    N, V, _ = x.shape
    X, Y = x[:, :, 0], x[:, :, 1]
    x_xi, x_eta = X @ R["dx"].T, X @ R["dy"].T
    y_xi, y_eta = Y @ R["dx"].T, Y @ R["dy"].T
    detJ = x_xi * y_eta - x_eta * y_xi
    if not np.all(detJ > 0):
        raise ValueError("fem2d_P2: non-positive Jacobian")
    inv = 1.0 / detJ
    dx = (y_eta[:, :, None] * R["dx"][None] - y_xi[:, :, None] * R["dy"][None]) * inv[:, :, None]
    dy = (-x_eta[:, :, None] * R["dx"][None] + x_xi[:, :, None] * R["dy"][None]) * inv[:, :, None]
    w = (detJ * R["w"][None, :]).reshape(-1)
    ops = {"id": np.broadcast_to(np.eye(V), (N, V, V)).copy(), "dx": dx, "dy": dy}
    return ops, w


def fem2d_P2(bubble=None, K=None, t=None):
    """P2(+bubble) triangles (fem2d_P2.jl:262-277). K: (N, V, 2) full node mesh."""
    b = (K is None or K.shape[1] == 7) if bubble is None else bubble
    if K is None:
        R = reference_triangle(b)
        x = np.einsum("vc,ncd->nvd", R["K"], _DEFAULT_TRI)
    else:
        x = np.asarray(K, float)
    if t is None:
        t = _dedupe(x.reshape(-1, 2)).reshape(x.shape[0], x.shape[1])
    ops, w = _p2_operators(x, b)
    return Geometry(kind="p2", x=x, w=w, t=np.asarray(t, np.int64), operators=ops, dim=2, k=2,
                    bubble=b)


# --------------------------------------------------------------------------
# Spectral
# --------------------------------------------------------------------------

def _cheb_eval(xs, n):
    xs = np.asarray(xs, float).reshape(-1)
    M = np.empty((xs.size, n))
    M[:, 0] = 1.0
    if n >= 2:
        M[:, 1] = xs
    for j in range(2, n):
        M[:, j] = 2 * xs * M[:, j - 1] - M[:, j - 2]
    return M


def _cheb_derivative(n):
    D = np.zeros((n, n))
    for j in range(n - 1):
        for k in range(j + 1, n, 2):
            D[j, k] = 2 * k
    D[0, :] /= 2
    return D


def _clenshaw_curtis(n):
    """QuadratureRules.ClenshawCurtisQuadrature (un-vendored; spectral1d.jl:75-78): standard
    Clenshaw-Curtis on the n Chebyshev-Lobatto points, ascending, weights summing to 2 on [-1,1]
    (pinned by the goldens of test/runtests.jl:25-32)."""
    if n == 1:
        return np.array([0.0]), np.array([2.0])
    return _tf_nodes(n - 1), _tf_weights(n - 1)


def _spectral1d_levels(n):
    """spectral1d.jl:63-109. Returns dict with geometry pieces and per-level subspaces/refine."""
    L = int(np.ceil(np.log2(n)))
    ls = [min(n, 2 ** k) for k in range(1, L + 1)]
    xs, sub_d, sub_f, sub_u, refine = [], [], [], [], [None] * L
    for l in range(L):
        nodes, w = _clenshaw_curtis(ls[l])
        xs.append(nodes)
        M = _cheb_eval(nodes, ls[l])
        CI = M[:, 2:].copy()
        CI[:, 0::2] -= M[:, 0:1]
        CI[:, 1::2] -= M[:, 1:2]
        sub_d.append(CI)
        sub_f.append(M)
        sub_u.append(np.ones((ls[l], 1)))
    dx = M @ _cheb_derivative(ls[-1]) @ np.linalg.inv(M)
    refine[L - 1] = np.eye(ls[-1])
    for l in range(L - 1):
        refine[l] = _cheb_eval(xs[l + 1], ls[l]) @ np.linalg.inv(sub_f[l])
    return dict(x=xs[-1], w=w, dx=dx, subspaces={"dirichlet": sub_d, "full": sub_f, "uniform": sub_u},
                refine=refine, L=L)


def spectral1d(n=16):
    lv = _spectral1d_levels(n)
    ops = {"id": np.eye(n)[None], "dx": lv["dx"][None]}
    return Geometry(kind="spectral1d", x=lv["x"].reshape(1, n, 1), w=lv["w"], t=np.arange(n)[None],
                    operators=ops, dim=1, k=n, dense=True)


def spectral2d(n=4):
    """spectral2d.jl:15-42 (note :dx = kron(DX, I) acts on the slow index = coordinate column 2)."""
    lv = _spectral1d_levels(n)
    w = np.outer(lv["w"], lv["w"]).reshape(-1)
    xl = lv["x"]
    y = np.tile(xl, n)               # fast index
    z = np.repeat(xl, n)             # slow index
    x = np.stack([y, z], axis=1)
    ID, DX = np.eye(n), lv["dx"]
    ops = {"id": np.kron(ID, ID)[None], "dx": np.kron(DX, ID)[None], "dy": np.kron(ID, DX)[None]}
    return Geometry(kind="spectral2d", x=x.reshape(1, n * n, 2), w=w, t=np.arange(n * n)[None],
                    operators=ops, dim=2, k=n, dense=True)


# --------------------------------------------------------------------------
# subdivide / find_boundary
# --------------------------------------------------------------------------

def subdivide(geom: Geometry, L: int) -> Geometry:
    """Refine by L-1 levels of geometric subdivision (multigrid.jl:472)."""
    if geom.dense or L <= 1:
        return geom
    if geom.kind == "p1":
        x, t = geom.x, geom.t
        for _ in range(L - 1):
            x = np.einsum("rc,ncd->nrd", _P1_REFINE, x).reshape(-1, 3, 2)
            t = _refine_p1_connectivity(t)
        return fem2d_P1(K=x, t=t)
    if geom.kind == "p2":
        V = geom.V
        Rm = _p2_refine_matrix(geom.bubble)
        x, t = geom.x, geom.t
        for _ in range(L - 1):
            x = np.einsum("rc,ncd->nrd", Rm, x).reshape(-1, V, 2)
            t = _refine_p2_connectivity(t)
        return fem2d_P2(bubble=geom.bubble, K=x, t=t)
    if geom.kind == "tensor":
        d, k = geom.dim, geom.k
        P = _tf_refine_local(k, d)
        n = geom.V
        x, t = geom.x, geom.t
        for _ in range(L - 1):
            x = np.einsum("rc,ncd->nrd", P, x).reshape(-1, n, x.shape[2])
            t = _tf_refine_connectivity(t, k, d)
        return _tf_build_geometry(d, k, x, t)
    raise ValueError(geom.kind)


def _boundary_ids_faces(t, faces_local):
    """Ids on (d-1)-faces used by exactly one element."""
    N = t.shape[0]
    sigs = np.concatenate([np.sort(t[:, fl], axis=1) for fl in faces_local], axis=0)
    _, inv, cnt = np.unique(sigs, axis=0, return_inverse=True, return_counts=True)
    once = cnt[inv.reshape(-1)] == 1
    return np.unique(sigs[once].reshape(-1))


def boundary_node_ids(geom: Geometry) -> np.ndarray:
    """Global node ids on the boundary (the id set behind find_boundary)."""
    t = geom.t
    if geom.kind == "tensor":
        ref = _TFRef(geom.dim, geom.k)
        faces = []
        for a in range(geom.dim):
            for layer in (0, ref.s - 1):
                faces.append(np.nonzero(ref.mi[:, a] == layer)[0])
        return _boundary_ids_faces(t, faces)
    if geom.kind == "p1":
        return _boundary_ids_faces(t, [np.array([0, 1]), np.array([1, 2]), np.array([2, 0])])
    if geom.kind == "p2":
        half = [np.array([a, (a + 1) % 6]) for a in range(6)]
        return _boundary_ids_faces(t, half)
    if geom.kind == "spectral1d":
        return np.array([0, geom.k - 1])
    if geom.kind == "spectral2d":
        n = geom.k
        j, i = np.indices((n, n)).reshape(2, -1)
        return np.nonzero((i == 0) | (i == n - 1) | (j == 0) | (j == n - 1))[0]
    raise ValueError(geom.kind)


def find_boundary(geom: Geometry):
    """(v, e) pairs (0-based) of nodes on the boundary (multigrid.jl:434-461)."""
    ids = boundary_node_ids(geom)
    mask = np.isin(geom.t, ids)
    e, v = np.nonzero(mask)
    return list(zip(v.tolist(), e.tolist()))
