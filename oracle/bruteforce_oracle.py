"""CPU oracle for the kernel-product / attention / solve hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import it, and only as the checker (or as
the CPU arm being timed) -- never as a fallback for the CUDA path.

It restates, in blocked NumPy, the algorithm of the reference brute force
(all citations are into /root/reference/kernel_matrix_benchmarks/):

* kernel formulas              algorithms/bruteforce.py:18-22 and :8-15
* squared distances            algorithms/bruteforce.py:36-49 (BLAS form) and :53-54 (difference form)
* product / density / attention algorithms/bruteforce.py:130-153
* float64 contiguous output    algorithms/base.py:107-116
* the dense solve              algorithms/bruteforce.py:205-207

The reference materialises the dense (N, M) matrix (and an (N, M, D) temporary
on the difference path); this restatement walks the same matrix in
(row block x source block) tiles so that N = M = 10^6 row samples fit in RAM.
Summation order therefore differs from the reference's single GEMM, which
moves results by ~1e-15 relative in float64.

Parity pin: ``tests/golden/make_golden.py`` imports the *reference* classes
from /root/reference, runs them on seeded inputs and commits inputs+outputs
under ``tests/golden/``; ``tests/test_oracle.py`` checks this module against
every one of those vectors (<= 1e-12 relative in float64).
"""
from __future__ import annotations

import numpy as np

KERNELS = ("gaussian", "absolute-exponential", "inverse-distance")


def _as_dtype(precision):
    """``precision`` arrives as a numpy dtype or a YAML string (algos.yaml:156-162)."""
    return np.dtype(precision)


def apply_kernel(kernel, sqdists, row0, n_sources):
    """Pointwise kernel on a tile of squared distances (bruteforce.py:18-22).

    ``row0`` is the global index of the tile's first row and the tile spans
    *all* ``n_sources`` columns; both are needed for the reference's
    'inverse-distance' convention, which zeroes every flat index that is a
    multiple of ``M + 1`` in the (N, M) matrix (bruteforce.py:12-14) -- the
    diagonal when N <= M + 1.
    """
    if kernel == "gaussian":
        return np.exp(-sqdists)
    if kernel == "absolute-exponential":
        return np.exp(-np.sqrt(np.maximum(sqdists, 0)))
    if kernel == "inverse-distance":
        with np.errstate(divide="ignore"):
            k = 1 / np.sqrt(np.maximum(sqdists, 0))
        rows = row0 + np.arange(sqdists.shape[0])
        # flat = i*M + j is a multiple of M+1  <=>  j == i mod (M+1)
        jz = rows % (n_sources + 1)
        hit = jz < n_sources
        k[np.nonzero(hit)[0], jz[hit]] = 0
        return k
    raise NotImplementedError(f"oracle: unsupported kernel {kernel!r}")


def sqdists_tile(x_blk, y, fast_sqdists):
    """Squared distances of a row block against all sources.

    Difference form follows bruteforce.py:53-54 (non-negative by construction);
    BLAS form follows bruteforce.py:36-49 (may dip below zero).
    """
    if fast_sqdists:
        xn = (x_blk**2).sum(-1)
        yn = (y**2).sum(-1)
        return xn[:, None] + yn[None, :] - 2 * x_blk @ y.T
    out = np.zeros((x_blk.shape[0], y.shape[0]), dtype=x_blk.dtype)
    # accumulate one coordinate at a time: same terms as sum(diffs**2, -1)
    # without the (n, M, D) temporary
    for d in range(x_blk.shape[1]):
        diff = x_blk[:, d : d + 1] - y[None, :, d]
        out += diff * diff
    return out


def kernel_rows(kernel, x_blk, y, row0, precision=np.float64, fast_sqdists=False):
    """Rows ``row0 .. row0+len(x_blk)`` of the dense kernel matrix (bruteforce.py:25-58)."""
    dt = _as_dtype(precision)
    x_blk = np.ascontiguousarray(x_blk, dtype=dt)
    y = np.ascontiguousarray(y, dtype=dt)
    return apply_kernel(kernel, sqdists_tile(x_blk, y, fast_sqdists), row0, y.shape[0])


def kernel_product(
    kernel,
    source_points,
    target_points,
    source_signal,
    *,
    normalize_rows=False,
    density_estimation=False,
    precision=np.float64,
    fast_sqdists=False,
    rows=None,
    row_block=None,
):
    """a_i = sum_j k(x_i, y_j) b_j with the reference's four query modes
    (bruteforce.py:130-153), returned as float64 (base.py:116).

    ``rows``: optional index array -- evaluate only those target rows (used
    for sampled truth at N = 10^6); row indices stay global so the
    inverse-distance zeroing matches the full matrix.
    """
    dt = _as_dtype(precision)
    y = np.ascontiguousarray(source_points, dtype=dt)
    x = y if target_points is None else np.ascontiguousarray(target_points, dtype=dt)
    M, D = y.shape
    row_ids = np.arange(x.shape[0]) if rows is None else np.asarray(rows)
    n_out = len(row_ids)

    if normalize_rows and density_estimation:
        return np.ones((n_out, 1), dtype=np.float64)  # bruteforce.py:134-138

    if density_estimation:
        b = np.ones((M, 1), dtype=dt)  # bruteforce.py:150: K.sum(-1)
    else:
        b = np.ascontiguousarray(source_signal, dtype=dt)
    if normalize_rows:
        b = np.concatenate((b, np.ones_like(b[:, :1])), axis=1)  # bruteforce.py:142-143

    if row_block is None:
        # keep the (rows, M) tile around 64 MB
        row_block = max(1, min(4096, (8 << 20) // max(M, 1)))
    out = np.empty((n_out, b.shape[1]), dtype=dt)
    contiguous = rows is None
    for s in range(0, n_out, row_block):
        ids = row_ids[s : s + row_block]
        if contiguous:
            k = apply_kernel(kernel, sqdists_tile(x[ids], y, fast_sqdists), int(ids[0]), M)
        else:
            # arbitrary rows: apply the zeroing rule row by row through global ids
            sq = sqdists_tile(x[ids], y, fast_sqdists)
            if kernel == "inverse-distance":
                k = np.empty_like(sq)
                for t, i in enumerate(ids):
                    k[t : t + 1] = apply_kernel(kernel, sq[t : t + 1], int(i), M)
            else:
                k = apply_kernel(kernel, sq, 0, M)
        out[s : s + len(ids)] = k @ b
    if normalize_rows:
        out = out[:, :-1] / out[:, -1:]  # bruteforce.py:145
    return np.ascontiguousarray(out, dtype=np.float64)


def dense_kernel_matrix(kernel, source_points, target_points=None, precision=np.float64, fast_sqdists=False):
    """The full (N, M) matrix, for small cases only (bruteforce.py:25-58)."""
    y = np.ascontiguousarray(source_points, dtype=_as_dtype(precision))
    x = y if target_points is None else np.ascontiguousarray(target_points, dtype=y.dtype)
    return kernel_rows(kernel, x, y, 0, precision, fast_sqdists)


def kernel_solve_lstsq(kernel, source_points, target_signal, precision=np.float64, fast_sqdists=False):
    """The reference solver: minimum-norm least squares on the dense matrix
    (bruteforce.py:193-207, scipy.linalg.lstsq -> LAPACK gelsd)."""
    from scipy.linalg import lstsq

    K = dense_kernel_matrix(kernel, source_points, None, precision, fast_sqdists)
    a = np.ascontiguousarray(target_signal, dtype=K.dtype)
    return np.ascontiguousarray(lstsq(K, a)[0], dtype=np.float64)


def kernel_solve_spd(kernel, source_points, target_signal, lam=0.0):
    """Dense float64 solve of (K + lam I) b = a with the SPD LAPACK call the
    reference left commented out at bruteforce.py:206.  This is what the CG
    solver is compared with (the un-regularised Gaussian system is singular to
    working precision)."""
    from scipy.linalg import solve

    K = dense_kernel_matrix(kernel, source_points)
    K[np.diag_indices_from(K)] += lam
    a = np.ascontiguousarray(target_signal, dtype=np.float64)
    return np.ascontiguousarray(solve(K, a, assume_a="pos"), dtype=np.float64)


def regularised_matvec(kernel, points, v, lam):
    """(K + lam I) v in float64 through the blocked product (used to score CG residuals)."""
    return kernel_product(kernel, points, None, v) + lam * np.asarray(v, dtype=np.float64)


def rel_l2(result, truth):
    """Relative L2 error used by every parity test (BASELINE.json north_star)."""
    result = np.asarray(result, dtype=np.float64)
    truth = np.asarray(truth, dtype=np.float64)
    den = np.linalg.norm(truth)
    num = np.linalg.norm(result - truth)
    if den == 0:  # an all-zero truth (e.g. the 1 x 1 inverse-distance matrix): exact match or infinite error
        return 0.0 if num == 0 else float("inf")
    return float(num / den)
