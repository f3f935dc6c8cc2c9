"""Host-side logic of the multi-GPU element partition (multigridbarrier.jl_b200/partition.py), on the CPU.

(1) structural identities of the sliced prolongations / transfers;
(2) world_size-2 `gloo` run: every rank evaluates f0 / f1 / f2 of ITS shard with the CPU oracle, the partial sums are
    all-reduced exactly where libmgbx all-reduces them (objective scalars, shared part of R'g, shared block of R'HR),
    and the result must equal the single-rank evaluation.  This is the N > 1 data path with the GPU kernels replaced
    by the oracle."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp

import mgb_oracle as O
from mgbx import geometry as G, hierarchy as H, partition as PT, problem as P


def _problem():
    return P.assemble(H.amg(G.subdivide(G.fem2d_P1(), 3)), p=1.5)


def _local_vector(M, Ml, x, i0, i1):
    """global level-L vector -> this rank's layout (shared segments whole, node-local segments sliced)."""
    offs = M.var_offsets[-1]
    parts = []
    for v in range(M.nu):
        seg = x[offs[v]:offs[v + 1]]
        parts.append(seg[i0:i1] if Ml.var_local[v] else seg)
    return np.concatenate(parts)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_operators_are_consistent(world):
    prob = _problem()
    for which in (0, 1):
        M = prob.M[which]
        n = M.geometry.n
        R = sp.csr_matrix(M.R_fine[-1])
        rng = np.random.default_rng(which)
        x = rng.normal(size=R.shape[1])
        y = rng.normal(size=R.shape[0])
        Rx = R @ x
        Rty = R.T @ y
        acc = np.zeros_like(Rty)
        covered = np.zeros(M.geometry.N, bool)
        for r in range(world):
            e0, e1 = PT.element_range(M.geometry.N, r, world)
            covered[e0:e1] = True
            Ml = PT.shard_amg(M, e0, e1)
            i0, i1 = Ml.node_range
            nl = i1 - i0
            assert Ml.n_global == n and Ml.geometry.n == nl and len(Ml.w) == nl
            Rl = sp.csr_matrix(Ml.R_fine[-1])
            xl = _local_vector(M, Ml, x, i0, i1)
            rows = np.concatenate([np.arange(v * n + i0, v * n + i1) for v in range(M.nu)])
            assert np.allclose(Rl @ xl, Rx[rows], rtol=0, atol=1e-14)
            # R' y: shared columns are partial sums over ranks, node-local columns are complete on their owner
            gl = Rl.T @ y[rows]
            offs, lo = M.var_offsets[-1], Ml.var_offsets[-1]
            for v in range(M.nu):
                if Ml.var_local[v]:
                    assert np.allclose(gl[lo[v]:lo[v + 1]], Rty[offs[v] + i0:offs[v] + i1], atol=1e-14)
                else:
                    acc[offs[v]:offs[v + 1]] += gl[lo[v]:lo[v + 1]]
            # level transfer: R_fine[L-2] restricted to the local rows == R_loc * T_loc
            if M.L > 1:
                R2 = sp.csr_matrix(M.R_fine[-2])[rows]
                assert abs(R2 - Rl @ sp.csr_matrix(Ml.T[-1])).max() < 1e-14
        assert covered.all()
        offs = M.var_offsets[-1]
        Ml0 = PT.shard_amg(M, 0, 1)
        for v in range(M.nu):
            if not Ml0.var_local[v]:
                assert np.allclose(acc[offs[v]:offs[v + 1]], Rty[offs[v]:offs[v + 1]], atol=1e-13)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        prob = _problem()
        M = prob.M[0]
        n = M.geometry.n
        t = 0.7
        rng = np.random.default_rng(5)
        x = 1e-3 * rng.normal(size=M.R_fine[-1].shape[1])
        bw = np.full(n, 1.0 / n)                       # the 1/n average as explicit (global) barrier weights
        # single-rank reference
        B = O.Barrier(prob.Q, bw)
        z0 = prob.g.T.reshape(-1).copy()
        f0 = B.f0(x, M.w, t * prob.f, M.R_fine[-1], O.operators(M), z0)
        f1 = B.f1(x, M.w, t * prob.f, M.R_fine[-1], O.operators(M), z0)
        f2 = sp.csr_matrix(B.f2(x, M.w, t * prob.f, M.R_fine[-1], O.operators(M), z0))
        # this rank's shard
        lp = PT.shard_problem(prob, rank, world)
        Ml = lp.M[0]
        i0, i1 = lp.node_range
        Bl = O.Barrier(lp.Q, PT.shard_barrier_weights(bw, i0, i1))
        xl = _local_vector(M, Ml, x, i0, i1)
        z0l = lp.g.T.reshape(-1).copy()
        opsl = O.operators(Ml)
        f0l = Bl.f0(xl, Ml.w, t * lp.f, Ml.R_fine[-1], opsl, z0l)
        f1l = Bl.f1(xl, Ml.w, t * lp.f, Ml.R_fine[-1], opsl, z0l)
        f2l = sp.csr_matrix(Bl.f2(xl, Ml.w, t * lp.f, Ml.R_fine[-1], opsl, z0l)).toarray()
        offs, lo = M.var_offsets[-1], Ml.var_offsets[-1]
        sh = [v for v in range(M.nu) if not Ml.var_local[v]]
        # all-reduce exactly what the library all-reduces
        s0 = torch.tensor([f0l], dtype=torch.float64)
        dist.all_reduce(s0)
        ok = abs(s0.item() - f0) <= 1e-12 * max(1.0, abs(f0))
        for v in range(M.nu):
            if Ml.var_local[v]:
                ok = ok and np.allclose(f1l[lo[v]:lo[v + 1]], f1[offs[v] + i0:offs[v] + i1], rtol=1e-12, atol=1e-14)
            else:
                g = torch.from_numpy(f1l[lo[v]:lo[v + 1]].copy())
                dist.all_reduce(g)
                ok = ok and np.allclose(g.numpy(), f1[offs[v]:offs[v + 1]], rtol=1e-11, atol=1e-13)
        for va in sh:
            for vb in sh:
                blk = torch.from_numpy(np.ascontiguousarray(f2l[lo[va]:lo[va + 1], lo[vb]:lo[vb + 1]]))
                dist.all_reduce(blk)
                ref = f2[offs[va]:offs[va + 1], offs[vb]:offs[vb + 1]].toarray()
                ok = ok and np.allclose(blk.numpy(), ref, rtol=1e-11, atol=1e-12 * abs(ref).max())
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_gloo_world2_partial_sums_match_single_rank():
    import torch.multiprocessing as mp
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_rank_main, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


# ---- row-sharded solve (cfg.shard_solve): row ownership and the push-to-peers exchange, on the CPU -----------------------

@pytest.mark.parametrize("rows,lpr,grid,nranks", [(261121, 1, 148, 2), (261121, 4, 148, 8), (130561, 4, 148, 4), (970299, 4, 148, 8),
                                                  (5, 1, 148, 2), (0, 1, 148, 2), (4737, 8, 3, 3), (100000, 32, 148, 8)])
def test_shard_row_ranges_partition_the_level(rows, lpr, grid, nranks):
    """mgbx_shard_row_range (host-only, csrc/pcg2.hpp pcg2_rank_rows): the ranks' row ranges are contiguous, ordered, disjoint and
    cover the level; every range starts on a slice boundary (32 / lpr rows), as the kernel's slice ownership requires."""
    from mgbx import native
    rps = 32 // lpr
    prev = 0
    for r in range(nranks):
        r0, r1 = native.shard_row_range(rows, lpr, grid, nranks, r)
        assert r0 == prev and r0 <= r1 <= rows
        assert r0 % rps == 0 or r0 == rows
        prev = r1
    assert prev == rows
    with pytest.raises(native.MgbxError):
        native.shard_row_range(rows, 3, grid, nranks, 0)


def _rank_rowshard(rank, world, port, out):
    """One rank of the CPU emulation of a row-sharded smoothing phase + fixed-order reduction (csrc/pcg2.cu, DIST = true)."""
    import torch
    import torch.distributed as dist
    from mgbx import native
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)          # identical (replicated) matrix and vectors on every rank
        m, grid, lpr = 4737, 6, 4
        A = sp.random(m, m, density=0.004, random_state=7, format="csr") + sp.identity(m, format="csr") * 4.0
        x, b = rng.normal(size=m), rng.normal(size=m)
        idg = 1.0 / A.diagonal()
        r0, r1 = native.shard_row_range(m, lpr, grid, world, rank)
        # the phase: this rank computes ITS rows of xnew = x + idg (b - A x) and "stores them into every peer's copy"
        mine = x[r0:r1] + idg[r0:r1] * (b[r0:r1] - A[r0:r1] @ x)
        parts = [None] * world
        dist.all_gather_object(parts, (r0, r1, mine))
        xnew = np.empty(m)
        for (a0, a1, v) in parts:
            xnew[a0:a1] = v
        ref = x + idg * (b - A @ x)
        ok = np.array_equal(xnew, ref) or np.allclose(xnew, ref, rtol=0, atol=1e-15)
        # the reduction: one partial per CTA of the joint grid, deposited in every rank's slot array, summed in CTA order
        rps = 32 // lpr
        nsl = (m + rps - 1) // rps
        spc = -(-nsl // (grid * world))
        slots = np.zeros(grid * world)
        for c in range(grid):
            g = rank * grid + c
            c0, c1 = min(m, g * spc * rps), min(m, (g + 1) * spc * rps)
            slots[g] = float(np.dot(b[c0:c1], xnew[c0:c1]))
        t = torch.from_numpy(slots)
        dist.all_reduce(t)                       # every slot is written by exactly one rank: the sum IS the exchange
        total = 0.0
        for v in t.numpy():
            total += float(v)                    # same order on every rank -> bitwise identical control flow
        allt = [None] * world
        dist.all_gather_object(allt, total)
        ok = ok and all(v == allt[0] for v in allt) and abs(total - float(b @ ref)) <= 1e-10 * abs(float(b @ ref))
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_gloo_world2_row_sharded_phase_matches_single_rank():
    import torch.multiprocessing as mp
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_rank_rowshard, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
