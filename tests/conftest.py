"""Shared fixtures.  ``-m "not gpu"`` runs here without a GPU; ``-m gpu`` runs on a B200."""
import glob
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN_DIR = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def product_golden_names():
    """Fixtures of product / attention / density tasks (the others pin the solver: solver_*, refsolve_*)."""
    return [n for n in golden_names() if not n.startswith(("solver_", "refsolve_"))]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    for k in ("kernel", "task"):
        g[k] = str(g[k])
    for k in ("same_points", "normalize_rows", "density_estimation"):
        g[k] = bool(g[k])
    return g


@pytest.fixture(params=golden_names())
def golden(request):
    g = load_golden(request.param)
    g["name"] = request.param
    return g
