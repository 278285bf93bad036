"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the
header declares, the plugin classes mirror the reference interface, and nothing falls back to CPU."""
import inspect
import os
import re

import numpy as np
import pytest

from conftest import REPO

HEADER = os.path.join(REPO, "include", "kmb_b200.h")


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kmb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from kernel_matrix_benchmarks_b200 import _lib

    lib = _lib.load()  # raises if the .so has not been built: there is no fallback
    declared = _declared_symbols()
    assert len(declared) >= 11
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/kmb_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes signatures and header disagree"
    assert lib.kmb_abi_version() == 1


def test_error_codes_without_touching_the_gpu():
    from kernel_matrix_benchmarks_b200 import _lib
    import ctypes

    lib = _lib.load()
    need = ctypes.c_size_t(0)
    # argument validation happens before any CUDA call
    assert lib.kmb_product_workspace_bytes(10, 0, 3, 1, 0, 0, 0, ctypes.byref(need)) == _lib.KMB_ERR_INVALID
    assert b"bad sizes" in lib.kmb_last_error()
    assert lib.kmb_product_workspace_bytes(10, 10, 3, 1, 7, 0, 0, ctypes.byref(need)) == _lib.KMB_ERR_UNSUPPORTED
    assert lib.kmb_product_workspace_bytes(10, 10, 3, 2, 0, _lib.FLAG_DENSITY, 0, ctypes.byref(need)) == _lib.KMB_ERR_INVALID
    with pytest.raises(NotImplementedError):
        _lib.check(_lib.KMB_ERR_UNSUPPORTED)
    with pytest.raises(ValueError):
        _lib.check(_lib.KMB_ERR_INVALID)


def test_plugin_interface_mirrors_reference_base_classes():
    """Method names and keyword arguments of base.py:7-167 (what runner.py:73-176 calls)."""
    from kernel_matrix_benchmarks_b200.algorithms.b200 import B200Product, B200Solver

    assert B200Product.task == "product" and B200Solver.task == "solver"
    sig = lambda f: list(inspect.signature(f).parameters)
    assert sig(B200Product.prepare_data) == ["self", "source_points", "target_points", "same_points", "density_estimation"]
    assert sig(B200Product.prepare_query) == ["self", "source_signal"]
    assert sig(B200Solver.prepare_data) == ["self", "source_points"]
    assert sig(B200Solver.prepare_query) == ["self", "target_signal"]
    for cls in (B200Product, B200Solver):
        ctor = inspect.signature(cls.__init__).parameters
        assert all(ctor[k].kind is inspect.Parameter.KEYWORD_ONLY for k in ("kernel", "dimension", "normalize_rows", "precision"))
        for m in ("fit", "query", "get_result", "set_query_arguments", "get_additional", "get_memory_usage", "done", "__str__"):
            assert callable(getattr(cls, m))


def test_plugin_rejects_what_the_reference_rejects():
    from kernel_matrix_benchmarks_b200.algorithms.b200 import B200Product, B200Solver

    for cls in (B200Product, B200Solver):
        with pytest.raises(NotImplementedError):  # bruteforce.py:82-85
            cls(kernel="laplace", dimension=3)
        with pytest.raises(NotImplementedError):
            cls(kernel="gaussian", dimension=3, precision=np.int32)
    for cls in (B200Product, B200Solver):      # the reference's precision sweep (algos.yaml:156-181): all three accepted
        for precision in ("float16", "float32", "float64", np.float64):
            try:
                cls(kernel="gaussian", dimension=784, precision=precision)
            except RuntimeError as e:          # no GPU here: the only acceptable refusal
                assert "no CPU fallback" in str(e)


def test_no_cpu_fallback():
    """Without a GPU the plugin must refuse to run rather than compute on the host."""
    import torch
    from kernel_matrix_benchmarks_b200.algorithms.b200 import B200Product

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        B200Product(kernel="gaussian", dimension=3, precision="float32")


def test_product_code_does_not_import_the_oracle():
    pkg = os.path.join(REPO, "kernel_matrix_benchmarks_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+\.*oracle|libkmb_oracle|oracle/", src, flags=re.M), \
                    f"{f} reaches into oracle/"


def test_datasets_follow_reference_generator():
    from kernel_matrix_benchmarks_b200 import datasets

    ds = datasets.uniform_cube(100, 3)
    np.random.seed(103)
    assert np.array_equal(ds.source_points, np.random.rand(100, 3))
    assert ds.name == "product-cube-D3-E1-M100-N100-gaussian" and ds.same_points
    c4 = datasets.config_c4(n=64)
    assert c4.normalize_rows and c4.E == 64 and c4.D == 64 and not c4.same_points
    assert abs(datasets.scaled_radius(784) - (3 / 784) ** 0.5) < 1e-15


def _wave_work(plan, w, unit):
    """What the kernels' wave_work() gives unit (CTA or cluster) `unit` in wave w: (row tile, source blocks) or None."""
    R, C, W, R_last, C_last = plan[:5]
    Rw, Cw = (R_last, C_last) if w == W - 1 else (R, C)
    if unit >= Rw * Cw:
        return None
    t, c = divmod(unit, Cw)
    return w * R + t, range(plan_nsb * c // Cw, plan_nsb * (c + 1) // Cw), Cw


@pytest.mark.parametrize("n_tiles, nsb, grid, tile_bytes", [
    (79, 235, 148, 128 * 784 * 4), (40, 235, 74, 2 * 128 * 784 * 4),      # C3: single CTAs, CTA pairs
    (2048, 2048, 148, 128 * 64 * 4), (1024, 2048, 74, 128 * 64 * 8),      # C4
    (1, 1, 148, 1 << 20), (1, 500, 148, 1 << 20), (500, 1, 148, 1 << 12), (149, 3, 148, 1 << 16), (7813, 4, 148, 128 * 800 * 8),
    (3, 7, 2, 1 << 10), (150, 2, 74, 1 << 30),
])
def test_wave_schedule_covers_every_unit_once(n_tiles, nsb, grid, tile_bytes):
    """kmb_debug_plan_waves (the planner of the tensor kernels) + the kernels' wave_work rule: every (row tile, source
    block) is evaluated exactly once, a wave never uses more than `grid` units, row tiles split over C > 1 units have
    a partial record each, and the number of steps stays within a few percent of the ideal."""
    import ctypes

    from kernel_matrix_benchmarks_b200 import _lib

    global plan_nsb
    plan_nsb = nsb
    out = (ctypes.c_int64 * 7)()
    assert _lib.load().kmb_debug_plan_waves(n_tiles, nsb, grid, tile_bytes, out) == 0
    plan = list(out)
    R, C, W, R_last, C_last, slots_per_wave, partial_slots = plan
    assert 1 <= R and R * C <= grid and R_last * C_last <= grid and (W - 1) * R + R_last == n_tiles
    # the wave's row tiles stay L2-resident, unless there are too few source blocks to keep every unit busy otherwise
    assert R * tile_bytes <= max(48 << 20, tile_bytes) or R <= -(-grid // nsb)
    seen = np.zeros((n_tiles, nsb), dtype=np.int32)
    steps = 0
    for w in range(W):
        longest = 0
        for unit in range(grid):
            ww = _wave_work(plan, w, unit)
            if ww is None:
                continue
            tile, blocks, Cw = ww
            assert 0 <= tile < n_tiles
            seen[tile, blocks.start:blocks.stop] += 1
            longest = max(longest, len(blocks))
            if Cw > 1:   # this unit leaves a partial record
                assert w * slots_per_wave + unit < partial_slots
        steps += longest
    assert (seen == 1).all()
    ideal = n_tiles * nsb / grid
    assert steps <= max(1.12 * ideal + 1, ideal + 2), (steps, ideal, plan)


@pytest.mark.parametrize("n, ctas", [(1, 148), (4097, 1), (40000, 148), (100000, 148), (100000, 8 * 148), (300000, 2)])
def test_symmetric_unit_list_covers_the_upper_block_triangle_once(n, ctas):
    """kprod_sym.cuh: the strip-ordered unit list (kmb_debug_sym_unit, host logic only) enumerates every
    (row tile, source block) pair on or above the block diagonal exactly once, strip by strip, tile by tile, block by
    block, and the segment a unit reports contains it."""
    import ctypes

    from kernel_matrix_benchmarks_b200 import _lib

    lib = _lib.load()

    def unit(u):
        o = (ctypes.c_int64 * 8)()
        _lib.check(lib.kmb_debug_sym_unit(n, ctas, u, o))
        return list(o)

    total, n_strips, wb = unit(-1)[:3]
    nsb, n_tiles = (n + 511) // 512, (n + 4095) // 4096
    assert total == sum(nsb - 8 * t for t in range(n_tiles)) and wb % 8 == 0 and n_strips == -(-nsb // wb) <= 96
    seen, prev = set(), None
    for u in range(total):
        _, _, _, strip, tile, block, begin, length = unit(u)
        assert 8 * tile <= block < nsb and tile < n_tiles and strip == block // wb and begin <= u < begin + length
        key = (strip, tile, block)
        assert prev is None or key > prev
        prev = key
        seen.add((tile, block))
    assert len(seen) == total


def test_symmetric_strip_width_follows_the_cta_count():
    """Wb ~ nsb / sqrt(2 G): a CTA's contiguous range is a roughly square patch of the pair matrix."""
    import ctypes

    from kernel_matrix_benchmarks_b200 import _lib

    lib = _lib.load()
    o = (ctypes.c_int64 * 8)()
    _lib.check(lib.kmb_debug_sym_unit(1_000_000, 148, -1, o))
    assert (o[1], o[2]) == (17, 120)
    _lib.check(lib.kmb_debug_sym_unit(10_000_000, 8 * 148, -1, o))
    assert o[1] <= 96 and 380 <= o[2] <= 420


def test_auto_path_resolution():
    """KMB_PATH_AUTO: tensor path for D > 16 and for wide signals (E >= 32) of the kernels the FP16-plane P.B kernel covers;
    the direct FP32 kernel otherwise; explicit paths are returned unchanged (kmb_resolved_path, host logic only)."""
    from kernel_matrix_benchmarks_b200 import _lib
    from kernel_matrix_benchmarks_b200.product import resolved_path

    assert resolved_path(3, 1) == "direct" and resolved_path(16, 31) == "direct"
    assert resolved_path(17, 1) == "tensor_f16" and resolved_path(784, 1) == "tensor_f16"
    assert resolved_path(3, 64) == "tensor_f16" and resolved_path(3, 64, "absolute-exponential") == "tensor_f16"
    assert resolved_path(3, 64, "inverse-distance") == "direct"   # the P.B kernel with FP16 planes has no inverse-distance form
    assert resolved_path(3, 64, path="direct") == "direct" and resolved_path(64, 64, path="tensor_tf32") == "tensor_tf32"
    assert _lib.load().kmb_resolved_path(0, 1, 0, 0) == -1
