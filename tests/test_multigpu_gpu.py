"""Several GPUs behind one plugin object in one process (multigpu.DeviceGroup, kmb_product_*_multi_f32): what lets the
reference's single-threaded harness (runner.py:118-148) sweep ``n_gpus``.  The 2-GPU cases skip on a 1-GPU box; the
C-ABI reduction, the clamping rule and the n_gpus = 1 route run everywhere.  Oracle = checker only."""
import ctypes
import warnings

import numpy as np
import pytest

from oracle import bruteforce_oracle as orc
from oracle import c_oracle

pytestmark = pytest.mark.gpu


def _n_devices():
    import torch

    return torch.cuda.device_count()


needs2 = pytest.mark.skipif("_n_devices() < 2", reason="needs two GPUs in one box (gpurun --gpus 2)")


def run_plugin(kernel, y, x, b, *, n_gpus, same_points=False, normalize_rows=False, density=False, path="auto", as_query_arg=False):
    from kernel_matrix_benchmarks_b200.algorithms.b200 import B200Product

    algo = B200Product(kernel=kernel, dimension=y.shape[1], normalize_rows=normalize_rows, precision="float32", path=path,
                       n_gpus=1 if as_query_arg else n_gpus)
    try:
        algo.prepare_data(source_points=y, target_points=y if x is None else x, same_points=same_points, density_estimation=density)
        algo.fit()
        if as_query_arg:
            algo.set_query_arguments(n_gpus=n_gpus)   # runner.py:123: after fit(), before prepare_query()
        algo.prepare_query(source_signal=b)
        algo.query()
        out, extra, name = algo.get_result(), algo.get_additional(), str(algo)
    finally:
        algo.done()
    return out, extra, name


def test_reduce_parts_c_abi():
    """kmb_reduce_parts_f32: fixed-order sum of up to 16 part vectors (here all on one device), ragged length."""
    import torch
    from kernel_matrix_benchmarks_b200 import _lib

    lib = _lib.load()
    n = 100_003
    parts = [torch.randn(n, device="cuda") for _ in range(5)]
    out = torch.empty(n, device="cuda")
    arr = (ctypes.c_void_p * 5)(*[p.data_ptr() for p in parts])
    _lib.check(lib.kmb_reduce_parts_f32(ctypes.c_void_p(out.data_ptr()), arr, 5, n, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    want = parts[0].clone()
    for p in parts[1:]:
        want += p   # the same order, the same roundings
    assert torch.equal(out, want)
    assert lib.kmb_reduce_parts_f32(ctypes.c_void_p(out.data_ptr()), arr, 17, n, None) == _lib.KMB_ERR_INVALID


def test_n_gpus_is_clamped_to_the_box():
    """Asking for more GPUs than the box has must not abort the harness's sweep over query-argument groups."""
    from kernel_matrix_benchmarks_b200.algorithms.b200 import B200Product, B200Solver

    have = _n_devices()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        algo = B200Product(kernel="gaussian", dimension=3, n_gpus=64)
        sol = B200Solver(kernel="gaussian", dimension=3, lam=1.0)
        sol.set_query_arguments(n_gpus=64, rtol=1e-5)
    assert algo.n_gpus == have and sol.n_gpus == have and len(w) >= 2
    assert (f"gpus={have}" in str(algo)) == (have > 1)
    with pytest.raises(ValueError):
        B200Product(kernel="gaussian", dimension=3, n_gpus=2, distributed=True)


def test_n_gpus_one_is_the_plain_path():
    from kernel_matrix_benchmarks_b200 import datasets

    ds = datasets.uniform_cube(5000, 3, 1.0, "gaussian")
    out, extra, name = run_plugin("gaussian", ds.source_points, None, ds.source_signal, n_gpus=1, same_points=True, as_query_arg=True)
    assert extra["n_gpus"] == 1 and "gpus=" not in name
    assert orc.rel_l2(out, c_oracle.kernel_product("gaussian", ds.source_points, None, ds.source_signal)) <= 1e-5


@needs2
@pytest.mark.parametrize("as_query_arg", [False, True])
def test_symmetric_product_on_two_gpus(as_query_arg):
    """same_points Gaussian product: the unit list split over two devices, partial vectors added through peer memory."""
    from kernel_matrix_benchmarks_b200 import datasets

    ds = datasets.uniform_cube(70_001, 3, 1.0, "gaussian")
    rows = np.arange(0, ds.N, 97)
    want = c_oracle.kernel_product("gaussian", ds.source_points, None, ds.source_signal, rows=rows)
    out2, extra2, name2 = run_plugin("gaussian", ds.source_points, None, ds.source_signal, n_gpus=2, same_points=True, as_query_arg=as_query_arg)
    out1, extra1, _ = run_plugin("gaussian", ds.source_points, None, ds.source_signal, n_gpus=1, same_points=True)
    assert extra2["n_gpus"] == 2 and extra2["path_used"] == "direct_sym" and "gpus=2" in name2
    assert orc.rel_l2(out2[rows], want) <= 1e-5
    assert orc.rel_l2(out2, out1) <= 2e-6
    dens, _, _ = run_plugin("gaussian", ds.source_points, None, None, n_gpus=2, same_points=True, density=True)
    assert orc.rel_l2(dens[rows], c_oracle.kernel_product("gaussian", ds.source_points, None, None, density_estimation=True, rows=rows)) <= 1e-5


@needs2
@pytest.mark.parametrize("kernel,N,M,D,E,norm", [
    ("absolute-exponential", 3001, 5000, 3, 3, False),
    ("inverse-distance", 777, 300, 3, 1, False),      # the flat-index zeroing rule needs the global row offset
    ("gaussian", 2, 400, 3, 1, True),                 # fewer rows than a tile: the second device gets one row
    ("gaussian", 1, 400, 3, 1, False),                # the second device's shard is empty
    ("gaussian", 900, 2000, 784, 1, False),           # tensor path, E <= 4 (the shape of C3)
    ("absolute-exponential", 1000, 3000, 64, 64, True),   # tensor path with both contractions (the shape of C4)
])
def test_row_sharded_product_on_two_gpus(kernel, N, M, D, E, norm):
    rng = np.random.RandomState(N + M + D)
    r = 1.0 if D <= 3 else (3.0 / D) ** 0.5
    y, x, b = r * rng.rand(M, D), r * rng.rand(N, D), rng.randn(M, E)
    want = c_oracle.kernel_product(kernel, y, x, b, normalize_rows=norm)
    for as_query_arg in (False, True):
        out, extra, _ = run_plugin(kernel, y, x, b, n_gpus=2, normalize_rows=norm, as_query_arg=as_query_arg)
        assert out.shape == want.shape and extra["n_gpus"] == 2
        assert orc.rel_l2(out, want) <= (1e-5 if D <= 16 else 1e-4)


@needs2
def test_query_argument_sweep_reuses_the_fitted_object():
    """algos.yaml `query-args: [{n_gpus: 1}, {n_gpus: 2}]`: one fit(), then set_query_arguments / prepare_query / query per group."""
    from kernel_matrix_benchmarks_b200 import datasets
    from kernel_matrix_benchmarks_b200.algorithms.b200 import B200Product

    ds = datasets.config_c4(n=4096)
    want = c_oracle.kernel_product(ds.kernel, ds.source_points, ds.target_points, ds.source_signal, normalize_rows=True)
    algo = B200Product(kernel=ds.kernel, dimension=ds.D, normalize_rows=True, precision="float32")
    algo.prepare_data(source_points=ds.source_points, target_points=ds.target_points)
    algo.fit()
    names = []
    for g in (1, 2, 1, 2):
        algo.set_query_arguments(n_gpus=g)
        for _ in range(2):
            algo.prepare_query(source_signal=ds.source_signal)
            algo.query()
            assert orc.rel_l2(algo.get_result(), want) <= 1e-4
        assert algo.get_additional()["n_gpus"] == g
        names.append(str(algo))
    assert names[0] == names[2] != names[1] == names[3]
    algo.done()


@needs2
@pytest.mark.parametrize("E,path,mode", [(1, "auto", "symmetric"), (2, "auto", "rows")])
def test_solver_on_two_gpus(E, path, mode):
    """CG with the matvec spread over two devices of this process; vectors and preconditioner on the first one."""
    import torch
    from kernel_matrix_benchmarks_b200 import datasets, product
    from kernel_matrix_benchmarks_b200.algorithms.b200 import B200Solver

    n, lam = 40_000, 1.0
    ds = datasets.uniform_cube(n, 3, 1.0, "gaussian", "solver", signal_dim=E)
    y = torch.tensor(ds.source_points, dtype=torch.float32, device="cuda:0")
    b = torch.tensor(ds.source_signal, dtype=torch.float32, device="cuda:0")
    rhs = (product.kernel_product(y, y, b, path="direct") + lam * b).cpu().numpy().astype(np.float64)
    results = {}
    for g in (1, 2):
        algo = B200Solver(kernel="gaussian", dimension=3, precision="float32", lam=lam, rtol=1e-6, path=path, preconditioner="nystrom")
        algo.prepare_data(source_points=ds.source_points)
        algo.fit()
        algo.set_query_arguments(n_gpus=g)
        algo.prepare_query(target_signal=rhs)
        algo.query()
        results[g] = (algo.get_result(), algo.get_additional())
        algo.done()
    (x1, e1), (x2, e2) = results[1], results[2]
    assert e2["n_gpus"] == 2 and e2["matvec"] == mode and e1["cg_converged"] and e2["cg_converged"]
    assert abs(e1["cg_iterations"] - e2["cg_iterations"]) <= 2
    assert orc.rel_l2(x2, x1) <= 1e-4 and orc.rel_l2(x2, ds.source_signal) <= 5e-4
    rows = np.arange(0, n, 400)
    want = orc.kernel_product("gaussian", ds.source_points, None, x2, rows=rows) + lam * x2[rows]
    assert orc.rel_l2(want, rhs[rows]) <= 2e-5
