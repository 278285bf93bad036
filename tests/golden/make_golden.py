#!/usr/bin/env python
"""Generate the golden vectors that pin the oracle to the *reference*.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports the unmodified reference classes
(/root/reference/kernel_matrix_benchmarks/algorithms/bruteforce.py:
``BruteForceProductBLAS`` :61-153, ``BruteForceSolverLAPACK`` :156-207), drives
them through the exact call sequence of ``runner.run`` (runner.py:77-143) and of
the ground-truth writer (datasets.py:180-195), and stores inputs + outputs as
small ``.npz`` fixtures next to this script.  /root/reference does not travel
to the GPU box, the fixtures do.

The reference has no golden vectors of its own, so these
reference-generated outputs are what pins parity.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, "/root/reference")

from kernel_matrix_benchmarks.algorithms.bruteforce import (  # noqa: E402  (the reference)
    BruteForceProductBLAS,
    BruteForceSolverLAPACK,
)

from kernel_matrix_benchmarks_b200 import datasets as gen  # noqa: E402


def reference_product(ds, precision="float64", fast_sqdists=False, pass_flags=True):
    """runner.py:73-143 for a product task (pass_flags=False: datasets.py:187-189)."""
    algo = BruteForceProductBLAS(
        kernel=ds.kernel,
        dimension=ds.D,
        normalize_rows=ds.normalize_rows,
        precision=precision,
        fast_sqdists=fast_sqdists,
    )
    kw = dict(same_points=ds.same_points, density_estimation=ds.density_estimation) if pass_flags else {}
    algo.prepare_data(source_points=ds.source_points, target_points=ds.target_points, **kw)
    algo.fit()
    algo.prepare_query(source_signal=ds.source_signal)
    algo.query()
    return algo.get_result()


def reference_solver(kernel, points, rhs, precision="float64"):
    algo = BruteForceSolverLAPACK(kernel=kernel, dimension=points.shape[1], precision=precision)
    algo.prepare_data(source_points=points)
    algo.fit()
    algo.prepare_query(target_signal=rhs)
    algo.query()
    return algo.get_result()


FORCE = "--force" in sys.argv   # default: only write fixtures that do not exist yet (committed ones stay byte-identical)


def exists(name):
    return os.path.exists(os.path.join(HERE, name + ".npz")) and not FORCE


def save(name, ds, **outputs):
    path = os.path.join(HERE, name + ".npz")
    if exists(name):
        print(f"{name:45s} kept")
        return
    np.savez_compressed(
        path,
        kernel=ds.kernel,
        task=ds.task,
        same_points=ds.same_points,
        normalize_rows=ds.normalize_rows,
        density_estimation=ds.density_estimation,
        source_points=ds.source_points,
        target_points=ds.target_points,
        source_signal=ds.source_signal,
        **outputs,
    )
    print(f"{name:45s} {os.path.getsize(path)/1024:8.1f} KiB  " + " ".join(f"{k}{v.shape}" for k, v in outputs.items()))


def check_generator_matches_reference():
    """uniform_cube: same legacy numpy calls in the same order as datasets.py:256-266."""
    n, D, radius = 257, 3, 1.0
    np.random.seed(n + D)
    pts = radius * np.random.rand(n, D)
    sig = np.random.randn(n, 1)
    ds = gen.uniform_cube(n, D, radius)
    assert np.array_equal(ds.source_points, pts) and np.array_equal(ds.source_signal, sig)
    assert ds.target_points is ds.source_points and ds.same_points


def main():
    warnings.simplefilter("ignore")  # the reference divides by zero on the inverse-distance diagonal
    check_generator_matches_reference()

    # 1. Gaussian product on the cube, x == y, E = 1 -- the shape of configs C1/C2
    ds = gen.uniform_cube(600, 3, 1.0, "gaussian")
    save(
        "product_gaussian_cube_d3",
        ds,
        truth=reference_product(ds),
        truth_groundtruth_call=reference_product(ds, pass_flags=False),
        ref_f32_fast=reference_product(ds, "float32", True),
        ref_f32_slow=reference_product(ds, "float32", False),
        ref_f64_fast=reference_product(ds, "float64", True),
    )

    # 2. inverse-distance on the sphere (the dataset the reference's CI runs,
    #    .github/workflows/benchmarks.yml:33)
    ds = gen.uniform_sphere(500, 1.0, "inverse-distance", seed=7)
    save("product_invdist_sphere_d3", ds, truth=reference_product(ds), ref_f32_slow=reference_product(ds, "float32", False))

    # 3. absolute-exponential, x != y, E = 3
    ds = gen.uniform_cube(700, 3, 1.0, "absolute-exponential", n_targets=333, signal_dim=3)
    save("product_absexp_cube_d3_e3_xy", ds, truth=reference_product(ds), ref_f32_fast=reference_product(ds, "float32", True))

    # 4. inverse-distance with N > M + 1: pins the flat-index zeroing rule (bruteforce.py:12-14)
    ds = gen.uniform_cube(37, 3, 1.0, "inverse-distance", n_targets=150)
    save("product_invdist_tall_d3", ds, truth=reference_product(ds))

    # 5. density estimation (b == 1, bruteforce.py:150) and its attention closed form (:134-138)
    ds = gen.uniform_cube(400, 3, 1.0, "gaussian", density_estimation=True)
    save("density_gaussian_cube_d3", ds, truth=reference_product(ds))
    ds = gen.uniform_cube(400, 3, 1.0, "gaussian", density_estimation=True, normalize_rows=True, task="attention")
    save("density_attention_gaussian_cube_d3", ds, truth=reference_product(ds))

    # 6. attention (row-normalised), D = 64, E = 8 -- the shape of config C4 in small
    for kernel in ("gaussian", "absolute-exponential"):
        ds = gen.uniform_cube(384, 64, gen.scaled_radius(64), kernel, "attention", normalize_rows=True, n_targets=256, signal_dim=8)
        save(
            f"attention_{kernel.replace('-', '')}_d64_e8",
            ds,
            truth=reference_product(ds),
            ref_f32_fast=reference_product(ds, "float32", True),
        )

    # 7. D = 784, scaled radius -- the shape of config C3 in small
    ds = gen.uniform_cube(192, 784, gen.scaled_radius(784), "gaussian", n_targets=96)
    save("product_gaussian_d784", ds, truth=reference_product(ds), ref_f32_fast=reference_product(ds, "float32", True))

    # 8. mid-size D (the direct FP32 path covers D <= 16)
    ds = gen.uniform_cube(300, 16, gen.scaled_radius(16), "gaussian", n_targets=200, signal_dim=2)
    save("product_gaussian_d16_e2", ds, truth=reference_product(ds))

    # 9. solver: the reference's lstsq (bruteforce.py:207) on a = K b, and the
    #    regularised SPD system (K + lam I) b = a that CG is scored on
    from oracle import bruteforce_oracle as orc

    for kernel, lam in (("gaussian", 1.0), ("gaussian", 1e-2)):
        ds = gen.uniform_cube(400, 3, 1.0, kernel, "solver")
        Kb = reference_product(ds)  # a = K b in float64 through the reference
        rhs = Kb + lam * ds.source_signal
        save(
            f"solver_{kernel}_cube_d3_lam{lam:g}",
            ds,
            rhs_unregularised=Kb,
            ref_lstsq_unregularised=reference_solver(kernel, ds.source_points, Kb),
            rhs=rhs,
            lam=np.float64(lam),
            spd_solution=orc.kernel_solve_spd(kernel, ds.source_points, rhs, lam),
        )

    # 10. attention with E = 64 -- the signal width of config C4 (the kernel with both contractions on the tensor
    #     cores, kprod_tensor_pv16), from the reference class
    for kernel in ("gaussian", "absolute-exponential"):
        name = f"attention_{kernel.replace('-', '')}_d64_e64"
        if not exists(name):
            ds = gen.uniform_cube(320, 64, gen.scaled_radius(64), kernel, "attention", normalize_rows=True, n_targets=160, signal_dim=64)
            save(name, ds, truth=reference_product(ds))

    # 11. the reference's OWN solve: K b = a with no regularisation (bruteforce.py:205-207, lstsq), on the point sets
    #     its solver datasets use (datasets.py:391-413: Fibonacci-sphere points for "solver-sphere-*-inverse-distance"
    #     AND for "solver-cube-*-gaussian"), right-hand side a = K b as write_output stores it (datasets.py:180-195)
    for label, kernel in (("sphere_invdist", "inverse-distance"), ("cube_gaussian", "gaussian")):
        name = f"refsolve_{label}_d3"
        if not exists(name):
            ds = gen.uniform_sphere(600, 1.0, kernel, "solver", seed=11)
            Kb = reference_product(ds)
            save(name, ds, rhs=Kb, lam=np.float64(0.0), ref_lstsq=reference_solver(kernel, ds.source_points, Kb),
                 ref_lstsq_f32=reference_solver(kernel, ds.source_points, Kb, "float32"))


if __name__ == "__main__":
    main()
