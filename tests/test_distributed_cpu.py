"""The N > 1 host logic on CPU: world_size-2 (and 3) gloo runs of the sharded CG loop and of the
row sharding of the product, with an oracle-backed stand-in for the CUDA kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kernel_matrix_benchmarks_b200.solver import LocalComm, TorchDistComm, cg_solve, shard_bounds
from oracle import bruteforce_oracle as orc


class OracleShardOps:
    """Same interface as CudaShardOps, arithmetic by the float64 oracle (CPU tensors)."""

    def __init__(self, points, kernel, lo, hi):
        self.pts, self.kernel, self.lo, self.hi = points, kernel, lo, hi

    def init(self, a):
        x = torch.zeros_like(a)
        return x, a.clone(), a.clone(), (a * a).sum(0)

    def matvec(self, p_full):
        rows = np.arange(self.lo, self.hi)
        if len(rows) == 0:
            return torch.zeros((0, p_full.shape[1]), dtype=torch.float64)
        return torch.from_numpy(orc.kernel_product(self.kernel, self.pts, None, p_full.numpy(), rows=rows))

    def shift_dot(self, Ap, p, lam, out):
        Ap += lam * p
        out.copy_((p * Ap).sum(0))
        return out

    def update(self, x, r, p, Ap, rs, pAp, rs_new):
        alpha = torch.where(pAp != 0, rs / pAp, torch.zeros_like(rs))
        x += alpha * p
        r -= alpha * Ap
        rs_new.copy_((r * r).sum(0))
        return rs_new

    def direction(self, p, r, rs_new, rs):
        beta = torch.where(rs != 0, rs_new / rs, torch.zeros_like(rs))
        p.mul_(beta).add_(r)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, lam, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.RandomState(11)
        pts, b = rng.rand(n, 3), rng.randn(n, 2)
        rhs = orc.regularised_matvec("gaussian", pts, b, lam)
        lo, hi, _ = shard_bounds(n, rank, world)
        comm = TorchDistComm()
        ops = OracleShardOps(pts, "gaussian", lo, hi)
        res = cg_solve(ops, comm, torch.from_numpy(rhs[lo:hi]).clone(), n, lam=lam, rtol=1e-10, max_iter=200)
        # the sharded product itself: every rank owns a row block, no collective on the data path
        prod = orc.kernel_product("gaussian", pts, None, b, rows=np.arange(lo, hi)) if hi > lo else np.zeros((0, 2))
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), x=res.x.numpy(), it=res.iterations, conv=res.converged,
                 lo=lo, hi=hi, prod=prod)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 301), (3, 100)])
def test_sharded_cg_gloo(tmp_path, world, n):
    lam = 1.0
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, lam, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.RandomState(11)
    pts, b = rng.rand(n, 3), rng.randn(n, 2)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    assert [int(p["lo"]) for p in parts] == [shard_bounds(n, r, world)[0] for r in range(world)]
    x = np.concatenate([p["x"] for p in parts])
    assert x.shape == (n, 2) and all(bool(p["conv"]) for p in parts)
    assert len({int(p["it"]) for p in parts}) == 1, "every rank must run the same number of iterations"
    assert orc.rel_l2(x, b) <= 1e-8  # rhs = (K + lam I) b
    prod = np.concatenate([p["prod"] for p in parts])
    assert orc.rel_l2(prod, orc.kernel_product("gaussian", pts, None, b)) <= 1e-13


def test_single_rank_matches_sharded_semantics():
    rng = np.random.RandomState(2)
    pts, b = rng.rand(150, 3), rng.randn(150, 1)
    lam = 0.5
    rhs = orc.regularised_matvec("gaussian", pts, b, lam)
    res = cg_solve(OracleShardOps(pts, "gaussian", 0, 150), LocalComm(), torch.from_numpy(rhs).clone(), 150, lam=lam,
                   rtol=1e-10, max_iter=300)
    assert res.converged and orc.rel_l2(res.x.numpy(), b) <= 1e-7
    dense = orc.kernel_solve_spd("gaussian", pts, rhs, lam)
    assert orc.rel_l2(res.x.numpy(), dense) <= 1e-7


def test_shard_bounds_cover_rows_exactly():
    for n in (0, 1, 7, 1000, 1_000_000):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, r, world)[:2] for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            per = shard_bounds(n, 0, world)[2]
            assert all(hi - lo <= per for lo, hi in spans)


# ---- symmetric mode: the triangular unit list split across ranks + one all-reduce per matvec ----------
# Host-level restatement of the partition of csrc/kprod_sym.cuh (sym_prefix / sym_tile_of / unit ranges)
# with float64 oracle arithmetic, so the N > 1 logic of CudaSymmetricOps is exercised on CPU.

def sym_prefix(I, nsb, TB):
    """units of tiles 0 .. I-1: sum_{t<I} (nsb - TB t).  (A tile-major stand-in for the unit list: the CUDA kernel orders
    the same units strip by strip -- kprod_sym.cuh, checked in test_abi_cpu.py -- which the host logic never sees.)"""
    return I * nsb - (TB * I * (I - 1)) // 2


def sym_tile_of(u, nsb, n_tiles, TB):
    I = 0
    while I + 1 < n_tiles and sym_prefix(I + 1, nsb, TB) <= u:
        I += 1
    return I


def sym_part(pts, b, part, n_parts, tile_rows=8, sb=4):
    """This part's share of K b for targets == sources: units on / above the block diagonal, each kernel
    value used for its row sum and (off the diagonal blocks) its column sum."""
    n = len(pts)
    TB = tile_rows // sb
    nsb, n_tiles = -(-n // sb), -(-n // tile_rows)
    units = sym_prefix(n_tiles, nsb, TB)
    ub, ue = units * part // n_parts, units * (part + 1) // n_parts
    out = np.zeros((n, 1))
    for u in range(ub, ue):
        I = sym_tile_of(u, nsb, n_tiles, TB)
        jb = TB * I + (u - sym_prefix(I, nsb, TB))
        rows = np.arange(I * tile_rows, min(n, (I + 1) * tile_rows))
        cols = np.arange(jb * sb, min(n, (jb + 1) * sb))
        if len(rows) == 0 or len(cols) == 0:
            continue
        K = np.exp(-((pts[rows][:, None, :] - pts[cols][None, :, :]) ** 2).sum(-1))
        out[rows] += K @ b[cols]
        if jb >= TB * (I + 1):          # off the block diagonal: the same values feed the column sums
            out[cols] += K.T @ b[rows]
    return out, units


@pytest.mark.parametrize("n,n_parts", [(50, 1), (50, 3), (37, 2), (8, 4), (5, 8)])
def test_symmetric_partition_adds_up(n, n_parts):
    rng = np.random.RandomState(n)
    pts, b = rng.rand(n, 3), rng.randn(n, 1)
    total = sum(sym_part(pts, b, p, n_parts)[0] for p in range(n_parts))
    assert orc.rel_l2(total, orc.kernel_product("gaussian", pts, None, b)) <= 1e-13
    # every unit is owned by exactly one part
    units = sym_part(pts, b, 0, n_parts)[1]
    owned = [units * (p + 1) // n_parts - units * p // n_parts for p in range(n_parts)]
    assert sum(owned) == units and max(owned) - min(owned) <= 1


class OracleSymmetricOps(OracleShardOps):
    """Same role as solver.CudaSymmetricOps: replicated CG vectors, the matvec is this rank's range of the
    symmetric unit list followed by one all-reduce."""

    def __init__(self, points, kernel, dist_comm):
        super().__init__(points, kernel, 0, len(points))
        self.dist_comm = dist_comm

    def matvec(self, p_full):
        part, _ = sym_part(self.pts, p_full.numpy(), self.dist_comm.rank, self.dist_comm.world)
        out = torch.from_numpy(part)
        self.dist_comm.all_reduce(out)
        return out


def _sym_worker(rank, world, port, n, lam, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.RandomState(5)
        pts, b = rng.rand(n, 3), rng.randn(n, 1)
        rhs = orc.regularised_matvec("gaussian", pts, b, lam)
        ops = OracleSymmetricOps(pts, "gaussian", TorchDistComm())
        res = cg_solve(ops, LocalComm(), torch.from_numpy(rhs).clone(), n, lam=lam, rtol=1e-10, max_iter=200)
        np.savez(os.path.join(out_dir, f"sym{rank}.npz"), x=res.x.numpy(), it=res.iterations, conv=res.converged)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 61), (3, 40)])
def test_symmetric_cg_gloo(tmp_path, world, n):
    lam = 1.0
    mp.spawn(_sym_worker, args=(world, _free_port(), n, lam, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.RandomState(5)
    pts, b = rng.rand(n, 3), rng.randn(n, 1)
    parts = [np.load(tmp_path / f"sym{r}.npz") for r in range(world)]
    assert all(bool(p["conv"]) for p in parts) and len({int(p["it"]) for p in parts}) == 1
    for p in parts:  # every rank holds the whole solution
        assert p["x"].shape == (n, 1) and orc.rel_l2(p["x"], b) <= 1e-8
    assert all(np.array_equal(parts[0]["x"], p["x"]) for p in parts[1:]), "replicated vectors must stay bit-identical"


# ---- Nystrom-preconditioned CG: same host logic on CPU, the kernel block from the oracle ----------------

def _oracle_block(kernel):
    def block(p, q):
        d2 = ((p[:, None, :] - q[None, :, :]) ** 2).sum(-1)
        return torch.exp(-d2) if kernel == "gaussian" else torch.exp(-torch.sqrt(d2))
    return block


def test_pcg_single_rank_converges_in_a_few_iterations():
    from kernel_matrix_benchmarks_b200.solver import NystromPreconditioner, landmark_indices, pcg_solve

    n, lam = 1500, 1.0
    rng = np.random.RandomState(3)
    pts, b = rng.rand(n, 3), rng.randn(n, 2)
    rhs = orc.regularised_matvec("gaussian", pts, b, lam)
    ops = OracleShardOps(pts, "gaussian", 0, n)
    plain = cg_solve(ops, LocalComm(), torch.from_numpy(rhs).clone(), n, lam=lam, rtol=1e-9, max_iter=300)
    tp = torch.from_numpy(pts)
    idx = landmark_indices(n, 300)
    assert len(set(idx.tolist())) == 300 and torch.equal(idx, landmark_indices(n, 300))
    pc = NystromPreconditioner(tp, tp[idx], "gaussian", lam, block_fn=_oracle_block("gaussian"), dtype=torch.float64)
    res = pcg_solve(ops, LocalComm(), torch.from_numpy(rhs).clone(), n, pc, lam=lam, rtol=1e-9, max_iter=50)
    assert res.converged and res.iterations <= 4 < plain.iterations
    assert orc.rel_l2(res.x.numpy(), b) <= 1e-6
    # the preconditioner is the exact inverse of K_hat + mu I on the span it was built from
    U, S = pc.U, pc.eigenvalues
    assert torch.allclose(U.T @ U, torch.eye(pc.rank, dtype=torch.float64), atol=1e-4)   # eigenvalues down to 1e-10 of the largest
    v = torch.from_numpy(rng.randn(n, 1))
    Mv = U @ (S.unsqueeze(1) * (U.T @ v)) + pc.mu * v
    assert orc.rel_l2(pc.apply(Mv).numpy(), v.numpy()) <= 1e-9
    # lam = 0 (the reference's system): mu falls back to the smallest kept eigenvalue, still SPD
    pc.set_shift(0.0)
    assert pc.mu > 0


def _pcg_worker(rank, world, port, n, lam, out_dir):
    from kernel_matrix_benchmarks_b200.solver import NystromPreconditioner, landmark_indices, pcg_solve

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.RandomState(12)
        pts, b = rng.rand(n, 3), rng.randn(n, 1)
        rhs = orc.regularised_matvec("absolute-exponential", pts, b, lam)
        lo, hi, _ = shard_bounds(n, rank, world)
        comm = TorchDistComm()
        tp = torch.from_numpy(pts)
        pc = NystromPreconditioner(tp[lo:hi], tp[landmark_indices(n, 200)], "absolute-exponential", lam, comm,
                                   block_fn=_oracle_block("absolute-exponential"), dtype=torch.float64)
        ops = OracleShardOps(pts, "absolute-exponential", lo, hi)
        res = pcg_solve(ops, comm, torch.from_numpy(rhs[lo:hi]).clone(), n, pc, lam=lam, rtol=1e-9, max_iter=100)
        np.savez(os.path.join(out_dir, f"pcg{rank}.npz"), x=res.x.numpy(), it=res.iterations, conv=res.converged, lo=lo, hi=hi,
                 rank_kept=pc.rank)
    finally:
        dist.destroy_process_group()


def test_pcg_sharded_gloo(tmp_path):
    world, n, lam = 2, 901, 1.0
    mp.spawn(_pcg_worker, args=(world, _free_port(), n, lam, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.RandomState(12)
    pts, b = rng.rand(n, 3), rng.randn(n, 1)
    parts = [np.load(tmp_path / f"pcg{r}.npz") for r in range(world)]
    assert all(bool(p["conv"]) for p in parts) and len({int(p["it"]) for p in parts}) == 1
    x = np.concatenate([p["x"] for p in parts], axis=0)
    assert orc.rel_l2(x, b) <= 1e-6
    # plain CG on the same system needs several times as many iterations
    plain = cg_solve(OracleShardOps(pts, "absolute-exponential", 0, n), LocalComm(),
                     torch.from_numpy(orc.regularised_matvec("absolute-exponential", pts, b, lam)).clone(), n, lam=lam,
                     rtol=1e-9, max_iter=500)
    assert int(parts[0]["it"]) * 2 <= plain.iterations


# ---- symmetric mode with the preconditioner built from row shards and replicated, collective stop decision ----

def _sym_pcg_worker(rank, world, port, n, lam, out_dir):
    from kernel_matrix_benchmarks_b200.solver import NystromPreconditioner, ReplicatedComm, landmark_indices, pcg_solve

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.RandomState(21)
        pts, b = rng.rand(n, 3), rng.randn(n, 1)
        rhs = orc.regularised_matvec("gaussian", pts, b, lam)
        lo, hi, _ = shard_bounds(n, rank, world)
        comm = TorchDistComm()
        tp = torch.from_numpy(pts)
        pc = NystromPreconditioner(tp[lo:hi], tp[landmark_indices(n, 150)], "gaussian", lam, comm,
                                   block_fn=_oracle_block("gaussian"), dtype=torch.float64).replicate(n)
        assert pc.U.shape[0] == n and pc.comm.world == 1
        ops = OracleSymmetricOps(pts, "gaussian", comm)
        res = pcg_solve(ops, ReplicatedComm(comm), torch.from_numpy(rhs).clone(), n, pc, lam=lam, rtol=1e-9, max_iter=100)
        np.savez(os.path.join(out_dir, f"spcg{rank}.npz"), x=res.x.numpy(), it=res.iterations, conv=res.converged, U=pc.U.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 333), (3, 200)])
def test_symmetric_pcg_with_sharded_build_gloo(tmp_path, world, n):
    """B200Solver(distributed=True), symmetric matvec: every rank builds 1 / world of the Nystrom factors, U is all-gathered,
    the CG vectors are replicated and the ranks agree on when to stop (solver.ReplicatedComm.agree)."""
    from kernel_matrix_benchmarks_b200.solver import NystromPreconditioner, landmark_indices

    lam = 1.0
    mp.spawn(_sym_pcg_worker, args=(world, _free_port(), n, lam, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.RandomState(21)
    pts, b = rng.rand(n, 3), rng.randn(n, 1)
    parts = [np.load(tmp_path / f"spcg{r}.npz") for r in range(world)]
    assert all(bool(p["conv"]) for p in parts) and len({int(p["it"]) for p in parts}) == 1
    for p in parts:
        assert p["x"].shape == (n, 1) and orc.rel_l2(p["x"], b) <= 1e-6
        assert np.array_equal(p["U"], parts[0]["U"]), "every rank must hold the same U"
    # the replicated U spans the same space as a single-process build (columns may differ by sign / rotation within clusters)
    tp = torch.from_numpy(pts)
    one = NystromPreconditioner(tp, tp[landmark_indices(n, 150)], "gaussian", lam, block_fn=_oracle_block("gaussian"), dtype=torch.float64)
    v = torch.from_numpy(rng.randn(n, 1))
    many = NystromPreconditioner.__new__(NystromPreconditioner)
    many.U, many.comm, many.eigenvalues = torch.from_numpy(parts[0]["U"]), LocalComm(), one.eigenvalues
    many.set_shift(lam)
    assert orc.rel_l2(many.apply(v).numpy(), one.apply(v).numpy()) <= 1e-6
