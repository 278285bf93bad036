"""Synthetic inputs with the reference generators' semantics.

The reference builds its point clouds in
/root/reference/kernel_matrix_benchmarks/datasets.py:

* ``uniform_cube``   (datasets.py:248-282): ``numpy.random.seed(n + D)``,
  ``radius * rand(n, D)`` sources, ``randn(n, 1)`` signal, targets == sources.
* ``uniform_sphere`` (datasets.py:200-244): Fibonacci lattice on the 2-sphere,
  ``randn(n, 1)`` signal drawn from the *unseeded* global generator.

MNIST/GloVe downloads are unavailable offline and disabled in the reference
(datasets.py:421-426), so every BASELINE.json config is driven by these
generators.  The same legacy ``numpy.random`` calls are issued in the same
order, so for the cases the reference can express (x == y, E == 1) the arrays
are bit-identical to what ``uniform_cube(...)`` hands to ``write_output``
(checked in tests/golden/make_golden.py).  Two extensions the reference
generators cannot express are needed by the configs:
independent targets (x != y) and E > 1; they draw *after* the reference's own
calls so the shared prefix of the stream is unchanged.

Dataset names follow the reference contract
``{task}-{label}-D{D}-E{E}-M{M}-N{N}-{kernel}`` (algos.yaml:38).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np


@dataclass
class KernelDataset:
    """The four arrays + attributes of the reference HDF5 schema (datasets.py:1-70)."""

    name: str
    task: str  # "product" | "attention" | "solver"
    kernel: str
    source_points: np.ndarray  # (M, D) float64
    target_points: np.ndarray  # (N, D) float64 (is source_points when same_points)
    source_signal: np.ndarray  # (M, E) float64
    same_points: bool
    normalize_rows: bool = False
    density_estimation: bool = False
    lam: float = 0.0  # solver regularisation the right-hand side was built with

    @property
    def M(self):
        return self.source_points.shape[0]

    @property
    def N(self):
        return self.target_points.shape[0]

    @property
    def D(self):
        return self.source_points.shape[1]

    @property
    def E(self):
        return self.source_signal.shape[1]


def dataset_name(task, label, D, E, M, N, kernel):
    return f"{task}-{label}-D{D}-E{E}-M{M}-N{N}-{kernel}"


def scaled_radius(dimension):
    """radius = sqrt(3 / D) keeps E|x - y|^2 = 0.5 as in the D = 3 unit cube, so
    high-D Gaussian kernels do not underflow (the reference assumes pre-scaled
    data, datasets.py:37-38; SURVEY.md section 7 'hard parts')."""
    return 1.0 if dimension <= 3 else math.sqrt(3.0 / dimension)


def uniform_cube(
    n_points=1000,
    dimension=3,
    radius=1.0,
    kernel="gaussian",
    task="product",
    normalize_rows=False,
    n_targets=None,
    signal_dim=1,
    density_estimation=False,
):
    """Uniform sample of the cube [0, radius)^D (datasets.py:248-282).

    ``n_targets=None`` reproduces the reference (targets are the sources);
    an integer draws that many independent targets from the same stream.
    """
    np.random.seed(n_points + dimension)  # datasets.py:258
    source_points = radius * np.random.rand(n_points, dimension)  # :261-263
    source_signal = np.random.randn(n_points, 1)  # :266
    if signal_dim > 1:
        extra = np.random.randn(n_points, signal_dim - 1)
        source_signal = np.concatenate((source_signal, extra), axis=1)
    if n_targets is None:
        target_points, same = source_points, True
    else:
        target_points, same = radius * np.random.rand(n_targets, dimension), False
    if density_estimation:
        source_signal = np.ones((n_points, 1))  # datasets.py:172-174
    N = n_points if n_targets is None else n_targets
    E = source_signal.shape[1]
    return KernelDataset(
        name=dataset_name(task, "cube", dimension, E, n_points, N, kernel),
        task=task,
        kernel=kernel,
        source_points=source_points,
        target_points=target_points,
        source_signal=source_signal,
        same_points=same,
        normalize_rows=normalize_rows,
        density_estimation=density_estimation,
    )


def uniform_sphere(n_points=1000, radius=1.0, kernel="inverse-distance", task="product", normalize_rows=False, seed=None):
    """Fibonacci lattice on the sphere of the given radius (datasets.py:200-244).

    The reference draws the signal from the unseeded global generator
    (datasets.py:228); pass ``seed`` to make it reproducible.
    """
    i = np.arange(n_points, dtype=np.float64)
    golden = math.pi * (3.0 - math.sqrt(5.0))  # datasets.py:212
    yy = 1 - (i / float(n_points - 1)) * 2  # :216
    ry = np.sqrt(1 - yy * yy)  # :217
    theta = golden * i  # :219
    pts = np.stack((radius * np.cos(theta) * ry, radius * yy, radius * np.sin(theta) * ry), axis=1)
    if seed is not None:
        np.random.seed(seed)
    signal = np.random.randn(n_points, 1)
    return KernelDataset(
        name=dataset_name(task, "sphere", 3, 1, n_points, n_points, kernel),
        task=task,
        kernel=kernel,
        source_points=pts,
        target_points=pts,
        source_signal=signal,
        same_points=True,
        normalize_rows=normalize_rows,
    )


# --- the five BASELINE.json configs ------------------------------------------------


def config_c1():
    """Gaussian product N=M=10k, D=3, E=1 (the reference's own CPU-runnable case)."""
    return uniform_cube(10_000, 3, 1.0, "gaussian", "product")


def config_c2(n=1_000_000):
    """Gaussian product N=M=1M, D=3, E=1 (headline metric)."""
    return uniform_cube(n, 3, 1.0, "gaussian", "product")


def config_c3(m=60_000, n=10_000, d=784):
    """MNIST-shaped synthetic: M=60k sources, N=10k targets, D=784, E=1."""
    return uniform_cube(m, d, scaled_radius(d), "gaussian", "product", n_targets=n)


def config_c4(n=262_144, d=64, e=64, kernel="absolute-exponential"):
    """Exponential-kernel attention N=M=262k, D=64, E=64 (row-normalised)."""
    return uniform_cube(n, d, scaled_radius(d), kernel, "attention", normalize_rows=True, n_targets=n, signal_dim=e)


def config_c5(n=1_000_000, lam=1.0):
    """Gaussian solve (K + lam I) b = a at N=M=1M, D=3.  The right-hand side is
    *not* materialised here (it needs one product): callers build
    a = K b + lam b with the product under test or the oracle."""
    ds = uniform_cube(n, 3, 1.0, "gaussian", "solver")
    ds.lam = lam
    return ds
