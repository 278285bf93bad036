#!/usr/bin/env python
"""Time every BASELINE.json config on one GPU through the plugin API (device time of query()).

    python tools/bench_configs.py [c1 c3 c4 c5 ...] > profiles/configs_rNN.jsonl

Not the headline bench (that is bench.py, config C2); this records where the other configs stand.
Under torchrun (WORLD_SIZE > 1; products only) the plugin runs with distributed=True: target rows sharded over the
ranks, sources replicated, no collective on the data path (the row blocks are gathered in get_result); the time
reported is the maximum over ranks.
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from kernel_matrix_benchmarks_b200 import datasets  # noqa: E402
from kernel_matrix_benchmarks_b200.algorithms.b200 import B200Product, B200Solver  # noqa: E402


WORLD = int(os.environ.get("WORLD_SIZE", "1"))
RANK = int(os.environ.get("RANK", "0"))


def run_product(name, ds, runs=3, **kw):
    if WORLD > 1:
        kw = dict(kw, distributed=True, device=int(os.environ.get("LOCAL_RANK", "0")))
    algo = B200Product(kernel=ds.kernel, dimension=ds.D, normalize_rows=ds.normalize_rows, precision="float32", **kw)
    algo.prepare_data(source_points=ds.source_points, target_points=ds.target_points, same_points=ds.same_points)
    algo.fit()
    algo.prepare_query(source_signal=ds.source_signal)
    from kernel_matrix_benchmarks_b200 import product as _product

    best = None
    _product.set_profiling(True)
    for _ in range(runs):
        t0 = time.perf_counter()
        algo.query()
        wall = time.perf_counter() - t0
        extra = algo.get_additional()
        try:
            extra["main_kernel_ms"] = _product.last_main_kernel_ms()   # tensor paths: the main kernel without the prepass
        except Exception:
            pass
        if best is None or extra["gpu_query_ms"] < best["gpu_query_ms"]:
            best = dict(extra, wall_ms=1e3 * wall)
    _product.set_profiling(False)
    res = algo.get_result()
    algo.done()
    if WORLD > 1:
        import torch
        import torch.distributed as dist

        t = torch.tensor([best["gpu_query_ms"], best.get("main_kernel_ms", 0.0)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best["gpu_query_ms"], best["main_kernel_ms"] = float(t[0]), float(t[1])
        best["gpairs_per_s"] = float(ds.N) * ds.M / (best["gpu_query_ms"] * 1e-3) / 1e9
        best["n_gpus"] = WORLD
        if RANK != 0:
            return
    pairs = float(ds.N) * ds.M
    out = {"config": name, "kernel": ds.kernel, "N": ds.N, "M": ds.M, "D": ds.D, "E": ds.E, "normalize_rows": ds.normalize_rows,
           "pairs": pairs, "finite": bool(np.isfinite(res).all()), **best}
    if ds.D > 16:
        pv = ds.E > 4 and ds.D <= 128  # second contraction on the tensor cores too
        flops = 2.0 * pairs * (ds.D + (ds.E if pv else 0))
        out["algorithmic_tflops"] = flops / (best["gpu_query_ms"] * 1e-3) / 1e12
        passes = -(-ds.E // 64) if pv else -(-ds.E // 4)
        out["executed_tf32_tflops"] = 3 * 2.0 * pairs * (ds.D * passes + (ds.E if pv else 0)) / (best["gpu_query_ms"] * 1e-3) / 1e12
    print(json.dumps(out), flush=True)


def run_solver(name, n, lam=1.0, rtol=1e-6):
    import torch
    from kernel_matrix_benchmarks_b200.product import kernel_product

    ds = datasets.config_c5(n, lam)
    # right-hand side a = K b + lam b, built with the product under test (the oracle cannot reach 1M)
    y = torch.tensor(ds.source_points, dtype=torch.float32, device="cuda")
    b = torch.tensor(ds.source_signal, dtype=torch.float32, device="cuda")
    rhs = (kernel_product(y, y, b) + lam * b).cpu().numpy().astype(np.float64)
    algo = B200Solver(kernel="gaussian", dimension=3, precision="float32", lam=lam, rtol=rtol, max_iter=500)
    algo.prepare_data(source_points=ds.source_points)
    algo.fit()
    algo.prepare_query(target_signal=rhs)
    t0 = time.perf_counter()
    algo.query()
    wall = time.perf_counter() - t0
    x = algo.get_result()
    extra = algo.get_additional()
    algo.done()
    err = float(np.linalg.norm(x - ds.source_signal) / np.linalg.norm(ds.source_signal))
    it = max(1, extra["cg_iterations"])
    print(json.dumps({"config": name, "N": n, "lam": lam, "rtol": rtol, "wall_ms": 1e3 * wall, "rel_l2_vs_generating_b": err,
                      "ms_per_iteration": extra["gpu_query_ms"] / it,
                      "matvec_gpairs_per_s": float(n) * n * it / (extra["gpu_query_ms"] * 1e-3) / 1e9, **extra}), flush=True)


def main():
    which = sys.argv[1:] or ["c1", "c2", "c3", "c4", "c5"]
    if WORLD > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
        which = [w for w in which if not w.startswith("c5")]
    if "c1" in which:
        run_product("C1", datasets.config_c1())
    if "c2" in which:
        run_product("C2", datasets.config_c2())
        run_product("C2-difference-form", datasets.config_c2(), path="direct_diff")
    if "c3" in which:
        run_product("C3", datasets.config_c3())
        run_product("C3-tf32-operands", datasets.config_c3(), path="tensor_tf32")
    if "c4s" in which:
        run_product("C4-small(32k)", datasets.config_c4(n=32768), runs=2)
    if "c4" in which:
        run_product("C4", datasets.config_c4(), runs=2)
        run_product("C4-gaussian", datasets.config_c4(kernel="gaussian"), runs=2)
    if "c5s" in which:
        run_solver("C5-small(100k)", 100_000)
    if "c5" in which:
        run_solver("C5", 1_000_000)


if __name__ == "__main__":
    main()
    if WORLD > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()
