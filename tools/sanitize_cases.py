#!/usr/bin/env python
"""One small case per kernel family, for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck python tools/sanitize_cases.py

Every case goes through the plugin (the C ABI) and is checked against the float64 oracle, so a run that the sanitizer
slows down 100x still says whether the kernels computed the right thing."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    import torch

    from kernel_matrix_benchmarks_b200 import product
    from kernel_matrix_benchmarks_b200.algorithms.b200 import B200Product, B200Solver
    from oracle import bruteforce_oracle as orc

    rng = np.random.RandomState(0)
    failures = 0

    def run(tag, kernel, y, x, b, tol, same=False, norm=False, precision="float32", path="auto"):
        nonlocal failures
        algo = B200Product(kernel=kernel, dimension=y.shape[1], normalize_rows=norm, precision=precision, path=path)
        algo.prepare_data(source_points=y, target_points=y if x is None else x, same_points=same)
        algo.fit()
        algo.prepare_query(source_signal=b)
        algo.query()
        out = algo.get_result()
        extra = algo.get_additional()
        algo.done()
        err = orc.rel_l2(out, orc.kernel_product(kernel, y, x, b, normalize_rows=norm))
        ok = err <= tol
        failures += not ok
        print(f"{'ok  ' if ok else 'FAIL'} {tag:<44} rel-L2 {err:.2e} (tol {tol:g}) path_used={extra.get('path_used')} launches={extra.get('gpu_launches')}", flush=True)

    # direct FP32 path: three kernels, plain / row-normalised, E = 1 and a wide signal (16 columns per pass)
    y, x = rng.rand(700, 3), rng.rand(300, 3)
    for kernel in ("gaussian", "absolute-exponential", "inverse-distance"):
        for norm in (False, True):
            run(f"direct {kernel} norm={norm}", kernel, y, x, rng.randn(700, 1), 1e-5, norm=norm)
    run("direct gaussian E=20", "gaussian", y, x, rng.randn(700, 20), 1e-5)
    run("direct gaussian difference form", "gaussian", 5 * y, 5 * x, rng.randn(700, 2), 1e-5)
    run("direct D=16", "gaussian", rng.rand(600, 16) * 0.4, rng.rand(200, 16) * 0.4, rng.randn(600, 3), 1e-5)
    # symmetric path (strip order, column slabs): product form, difference form, inverse-distance, 2 tiles
    ys = rng.rand(4500, 3)
    bs = rng.randn(4500, 1)
    for kernel, scale in (("gaussian", 1.0), ("gaussian", 6.0), ("absolute-exponential", 1.0), ("inverse-distance", 1.0)):
        pts = (scale * ys).astype(np.float32).astype(np.float64)
        t = torch.tensor(pts, dtype=torch.float32, device="cuda")
        tb = torch.tensor(bs, dtype=torch.float32, device="cuda")
        got = product.kernel_product(t, t, tb, kernel=kernel, path="direct_sym").cpu().numpy().astype(np.float64)
        parts = sum(product.kernel_product_sym_part(t, tb, p, 3, kernel=kernel).double() for p in range(3)).cpu().numpy()
        want = orc.kernel_product(kernel, pts, None, bs)
        e1, e2 = orc.rel_l2(got, want), orc.rel_l2(parts, want)
        ok = max(e1, e2) <= 1e-5
        failures += not ok
        print(f"{'ok  ' if ok else 'FAIL'} symmetric {kernel} scale={scale:<22} rel-L2 {e1:.2e} / 3 parts {e2:.2e}", flush=True)
    # float64 kernels
    run("float64 D=3", "gaussian", y, x, rng.randn(700, 2), 1e-12, precision="float64")
    run("float64 D=40 (tiled)", "absolute-exponential", rng.rand(300, 40) * 0.3, rng.rand(100, 40) * 0.3, rng.randn(300, 2), 1e-12, precision="float64")
    # tensor-core paths: E <= 4 (single CTA / CTA pairs), E > 4 FP16 planes (pv16, flush every 128 blocks), TF32 planes (pv)
    r = (3.0 / 96) ** 0.5
    yt, xt = r * rng.rand(700, 96), r * rng.rand(130, 96)
    run("tensor E=1 single CTA (1 row tile)", "gaussian", yt, xt[:100], rng.randn(700, 1), 1e-4)
    run("tensor E=2 CTA pairs", "gaussian", yt, xt, rng.randn(700, 2), 1e-4)
    run("tensor TF32 planes E=1", "gaussian", yt, xt, rng.randn(700, 1), 1e-4, path="tensor_tf32")
    r = (3.0 / 64) ** 0.5
    ya, xa = r * rng.rand(600, 64), r * rng.rand(300, 64)
    for kernel in ("gaussian", "absolute-exponential"):
        run(f"pv16 attention {kernel} E=64 (pairs)", kernel, ya, xa, rng.randn(600, 64), 1e-4, norm=True)
    run("pv16 product E=8 single CTA", "gaussian", ya, xa[:100], rng.randn(600, 8), 1e-4)
    run("pv (TF32 planes) inverse-distance E=8", "inverse-distance", ya, xa, rng.randn(600, 8), 1e-4)
    # CG vector kernels + the solver loop (plain and Nystrom-preconditioned)
    pts = rng.rand(1500, 3)
    bb = rng.randn(1500, 1)
    rhs = orc.regularised_matvec("gaussian", pts, bb, 1.0)
    for pc in ("none", "nystrom"):
        sol = B200Solver(kernel="gaussian", dimension=3, precision="float32", lam=1.0, rtol=1e-6, preconditioner=pc)
        sol.prepare_data(source_points=pts)
        sol.fit()
        sol.prepare_query(target_signal=rhs)
        sol.query()
        xs = sol.get_result()
        extra = sol.get_additional()
        sol.done()
        err = orc.rel_l2(xs, bb)
        ok = err <= 1e-4
        failures += not ok
        print(f"{'ok  ' if ok else 'FAIL'} solver preconditioner={pc:<22} rel-L2 {err:.2e} iterations={extra['cg_iterations']}", flush=True)
    torch.cuda.synchronize()
    print("FAILURES:", failures)
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
