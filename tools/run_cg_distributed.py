#!/usr/bin/env python
"""Sharded CG solve under torchrun (one process per GPU, NCCL): config C5 semantics at a chosen N.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P \
        tools/run_cg_distributed.py [N] [lam] [rows|symmetric] [nystrom|none]

rows:      every rank owns a contiguous block of rows of x, r, p, Ap; each iteration all-gathers p and
           all-reduces two scalars (solver.CudaShardOps + TorchDistComm).
symmetric: (default) the matvec is the symmetric product -- the ranks split its triangular unit list and
           all-reduce the N-float result; the CG vectors are replicated (solver.CudaSymmetricOps).
nystrom:   (default) Nystrom-preconditioned CG (solver.NystromPreconditioner, 1024 landmarks; in rows mode every rank
           holds its rows of U and the small products are all-reduced); none: plain CG.
Rank 0 prints one JSON line with the iteration count, the device time per iteration (max over ranks)
and the error against the generating signal.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from kernel_matrix_benchmarks_b200 import datasets  # noqa: E402
from kernel_matrix_benchmarks_b200.product import kernel_product  # noqa: E402
from kernel_matrix_benchmarks_b200.solver import (CudaShardOps, CudaSymmetricOps, LocalComm, NystromPreconditioner,  # noqa: E402
                                                  TorchDistComm, cg_solve, landmark_indices, pcg_solve, shard_bounds)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    lam = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    mode = sys.argv[3] if len(sys.argv) > 3 else "symmetric"
    precond = sys.argv[4] if len(sys.argv) > 4 else "nystrom"
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    comm = TorchDistComm() if world > 1 else LocalComm()

    ds = datasets.config_c5(n, lam)
    y = torch.tensor(ds.source_points, dtype=torch.float32, device=dev)
    b = torch.tensor(ds.source_signal, dtype=torch.float32, device=dev)
    lo, hi, _ = shard_bounds(n, rank, world)
    # this rank's rows of the right-hand side a = K b + lam b
    rhs = kernel_product(y[lo:hi], y, b, row_offset=lo) + lam * b[lo:hi]
    if mode == "symmetric":
        rhs = comm.all_gather(rhs, n).clone()   # replicated right-hand side
        ops, loop_comm = CudaSymmetricOps(y, "gaussian", comm), LocalComm()
        lo, hi = 0, n
    else:
        ops, loop_comm = CudaShardOps(y, "gaussian", lo, hi, path="direct"), comm
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e2.record()
    pc = None
    if precond == "nystrom":   # what B200Solver.fit() does
        pc = NystromPreconditioner(y[lo:hi], y[landmark_indices(n, 1024).to(dev)], "gaussian", lam, loop_comm)
    e0.record()
    if pc is not None:
        res = pcg_solve(ops, loop_comm, rhs, n, pc, lam=lam, rtol=1e-6, max_iter=500)
    else:
        res = cg_solve(ops, loop_comm, rhs, n, lam=lam, rtol=1e-6, max_iter=500)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    err2 = ((res.x - b[lo:hi]) ** 2).sum().double().reshape(1)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if mode != "symmetric":
            dist.all_reduce(err2)
    if rank == 0:
        it = max(1, res.iterations)
        print(json.dumps({"config": "C5-sharded", "matvec": mode, "preconditioner": precond if pc is None else f"nystrom(rank={pc.rank})",
                          "fit_ms": float(e2.elapsed_time(e0)), "N": n, "n_gpus": world, "lam": lam, "cg_iterations": res.iterations,
                          "converged": res.converged, "rel_residual": res.rel_residual, "total_ms": float(ms),
                          "ms_per_iteration": float(ms) / it,
                          "matvec_gpairs_per_s": float(n) * n * it / (float(ms) * 1e-3) / 1e9,
                          "rel_l2_vs_generating_b": float(err2.sqrt() / b.double().norm())}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
