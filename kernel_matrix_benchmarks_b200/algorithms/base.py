"""The plugin interface the B200 classes implement.

When the reference package is importable (the harness is driving us) its own
``BaseProduct`` / ``BaseSolver`` are re-exported, so the plugin *is* a subclass
of the reference's base classes.  On a machine without the reference (the GPU
box) an interface-compatible stand-in is defined: same method names, keyword
arguments, ``task`` attributes and defaults as
/root/reference/kernel_matrix_benchmarks/algorithms/base.py:7-167, which is all
``runner.run`` (runner.py:73-176) relies on.
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - exercised when the harness is on sys.path
    from kernel_matrix_benchmarks.algorithms.base import BaseAlgorithm, BaseProduct, BaseSolver  # type: ignore

    USING_REFERENCE_BASE = True
except Exception:  # ImportError, or the reference's own missing dependencies
    USING_REFERENCE_BASE = False

    class BaseAlgorithm:
        """Constructor contract of base.py:7-29; hooks of base.py:31-48."""

        def __init__(self, *, kernel, dimension, normalize_rows=False, precision=np.float64):
            self.kernel = kernel
            self.dimension = dimension
            self.precision = precision
            self.normalize_rows = normalize_rows
            self.name = "BaseAlgorithm()"

        def done(self):
            pass

        def get_memory_usage(self):
            """Resident set size in kB, sampled by the runner around fit() (runner.py:96-100)."""
            import psutil

            return psutil.Process().memory_info().rss / 1024

        def set_query_arguments(self, **kwargs):
            pass

        def get_additional(self):
            return {}

        def __str__(self):
            return self.name

    class BaseProduct(BaseAlgorithm):
        task = "product"  # base.py:54; runner.py:77 dispatches on it

        def prepare_data(self, *, source_points, target_points, same_points=False, density_estimation=False):
            pass

        def fit(self):
            pass

        def prepare_query(self, *, source_signal):
            pass

        def query(self):
            self.res = None

        def get_result(self):
            return np.ascontiguousarray(self.res, dtype=np.float64)  # base.py:116

    class BaseSolver(BaseAlgorithm):
        task = "solver"  # base.py:122; runner.py:87

        def prepare_data(self, *, source_points):
            pass

        def fit(self):
            pass

        def prepare_query(self, *, target_signal):
            pass

        def query(self):
            raise NotImplementedError()

        def get_result(self):
            return np.ascontiguousarray(self.res, dtype=np.float64)  # base.py:167
