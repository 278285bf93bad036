"""The C oracle (used for row samples at N = M = 10^6) against the golden vectors and the NumPy oracle."""
import numpy as np
import pytest

from oracle import bruteforce_oracle as orc
from oracle import c_oracle
from conftest import golden_names, load_golden, product_golden_names

CASES = product_golden_names()


@pytest.mark.parametrize("name", CASES)
def test_c_oracle_matches_reference(name):
    g = load_golden(name)
    out = c_oracle.kernel_product(
        g["kernel"], g["source_points"], None if g["same_points"] else g["target_points"], g["source_signal"],
        normalize_rows=g["normalize_rows"], density_estimation=g["density_estimation"])
    assert orc.rel_l2(out, g["truth"]) <= 1e-12


def test_c_oracle_row_sample_keeps_global_indices():
    g = load_golden("product_invdist_tall_d3")
    rows = np.array([149, 38, 0, 77])
    out = c_oracle.kernel_product("inverse-distance", g["source_points"], g["target_points"], g["source_signal"], rows=rows)
    assert orc.rel_l2(out, g["truth"][rows]) <= 1e-12


def test_c_oracle_speed_sample():
    """Keeps the CPU suite honest about what a 10^6-row sample costs."""
    rng = np.random.RandomState(1)
    y, b = rng.rand(200_000, 3), rng.randn(200_000, 1)
    rows = rng.choice(200_000, 64, replace=False)
    a = c_oracle.kernel_product("gaussian", y, None, b, rows=rows)
    assert orc.rel_l2(a, orc.kernel_product("gaussian", y, None, b, rows=rows)) <= 1e-12
