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
