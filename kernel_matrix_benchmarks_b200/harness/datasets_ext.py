"""Extra synthetic datasets for the reference harness.

The reference's generators hard-code x == y and E == 1, its registered "cube" names produce
Fibonacci-sphere points (datasets.py:400-413 call ``uniform_sphere``), and ``write_output``
computes the ground truth through one (N, M, D) float64 temporary (datasets.py:180-195), which
stops at N = M ~ 10^4.  The BASELINE configs need independent targets, E > 1, row-normalised
attention, radius scaling for high D and a lambda-consistent solver right-hand side.

The writers below emit the reference's dataset schema (datasets.py:1-70: four float64 arrays
+ the attributes ``write_output`` sets, datasets.py:150-177) for inputs drawn with the
reference's ``uniform_cube`` semantics (``..datasets.uniform_cube``).  The ground truth is
still produced by the reference's own ``GroundTruth`` class (= ``BruteForceProductBLAS`` in
float64 with the difference-form distances, datasets.py:81-83) -- it is merely fed blocks of
target rows so that its (rows, M, D) temporary stays below ``TEMP_BYTES``.  Rows of a kernel
product are independent, so blocking changes nothing for the Gaussian / exponential kernels;
the inverse-distance kernel zeroes entries by *flat* index (bruteforce.py:12-14) and is
therefore only written unblocked.

Beyond ``GPU_TRUTH_MIN_WORK`` (N M D; the BASELINE sizes C2-C5: 10^12 pairs would take the reference's float64
brute force hours) the ground truth is written by this library's float64 kernel (``kmb_product_f64``, csrc/kprod_f64.cu:
the same difference-form float64 arithmetic, reference outputs reproduced to 1e-12 on the golden vectors) and
``VERIFY_ROWS`` randomly sampled target rows are recomputed by the reference's ``GroundTruth`` and must agree to
``VERIFY_TOL``; what was done is recorded in the file's ``truth_source`` attribute.

Names follow the reference contract ``{task}-{label}-D{D}-E{E}-M{M}-N{N}-{kernel}``
(algos.yaml:38) with the label ``ucube`` ("uniform cube": the generator the reference defines
but never registers).
"""
from __future__ import annotations

import time

import numpy as np

from .. import datasets as gen

TEMP_BYTES = 1 << 30  # cap for the reference ground truth's (rows, M, D) float64 temporaries
UNBLOCKED_BYTES = 8 << 30  # what a kernel that cannot be blocked (inverse-distance) may use


def _ground_truth_blocked(*, kernel, source_points, target_points, source_signal, normalize_rows):
    """``target_signal`` exactly as datasets.py:180-195 computes it, in blocks of target rows."""
    from kernel_matrix_benchmarks.datasets import GroundTruth  # the reference's float64 brute force

    M, D = source_points.shape
    tp = source_points if target_points is None else target_points
    N = tp.shape[0]
    rows = int(max(1, TEMP_BYTES // (2 * 8 * M * D)))
    if kernel == "inverse-distance" and rows < N:
        if 2 * 8 * N * M * D > UNBLOCKED_BYTES:
            raise ValueError("inverse-distance zeroes entries by flat index (bruteforce.py:12-14): "
                             f"cannot block N={N}, M={M}, D={D} within {UNBLOCKED_BYTES} bytes")
        rows = N
    out = np.empty((N, source_signal.shape[1]), dtype=np.float64)
    for r0 in range(0, N, rows):
        gt = GroundTruth(kernel=kernel, dimension=D, normalize_rows=normalize_rows)
        gt.prepare_data(source_points=source_points, target_points=tp[r0:r0 + rows])
        gt.fit()
        gt.prepare_query(source_signal=source_signal)
        gt.query()
        out[r0:r0 + rows] = gt.get_result()
    return out


GPU_TRUTH_MIN_WORK = 2.0e10   # N * M * D from which the float64 GPU kernel writes the ground truth
VERIFY_ROWS = 512             # target rows of a GPU-written truth recomputed by the reference's GroundTruth
VERIFY_TOL = 1e-10            # relative L2 between the two on those rows


def _ground_truth_gpu(*, kernel, source_points, target_points, source_signal, normalize_rows, verify_rows=None, seed=11):
    """``target_signal`` by kmb_product_f64 (float64 on the GPU, K never materialised), spot-verified by the reference's
    ``GroundTruth`` (datasets.py:180-195) on sampled target rows.  Returns (truth, note)."""
    import torch

    from ..product import kernel_product_f64

    if kernel == "inverse-distance":
        raise ValueError("the sampled verification cannot reproduce the flat-index zeroing of inverse-distance (bruteforce.py:12-14)")
    tp = source_points if target_points is None else target_points
    N, M = tp.shape[0], source_points.shape[0]
    dev = torch.device("cuda", torch.cuda.current_device())
    y = torch.from_numpy(np.ascontiguousarray(source_points, dtype=np.float64)).to(dev)
    b = torch.from_numpy(np.ascontiguousarray(source_signal, dtype=np.float64)).to(dev)
    x = y if target_points is None else torch.from_numpy(np.ascontiguousarray(tp, dtype=np.float64)).to(dev)
    t0 = time.time()
    out = kernel_product_f64(x, y, b, kernel=kernel, normalize_rows=normalize_rows)
    torch.cuda.synchronize(dev)
    gpu_s = time.time() - t0
    truth = out.cpu().numpy()
    del out, x, y, b
    k = int(min(N, VERIFY_ROWS if verify_rows is None else verify_rows))
    rows = np.sort(np.random.RandomState(seed).choice(N, k, replace=False))
    t0 = time.time()
    ref = _ground_truth_blocked(kernel=kernel, source_points=source_points, target_points=np.ascontiguousarray(tp[rows]),
                                source_signal=source_signal, normalize_rows=normalize_rows)
    rel = float(np.linalg.norm(truth[rows] - ref) / max(np.linalg.norm(ref), 1e-300))
    if not rel <= VERIFY_TOL:
        raise RuntimeError(f"GPU float64 ground truth disagrees with the reference brute force on {k} sampled rows: rel-L2 {rel:.3e}")
    note = (f"kmb_product_f64 (float64 GPU kernel, {gpu_s:.1f} s); {k} sampled target rows recomputed by the reference GroundTruth "
            f"(datasets.py:180-195) in {time.time() - t0:.1f} s: rel-L2 {rel:.2e} <= {VERIFY_TOL:g}")
    return truth, note


def write_dataset(filename, ds, *, label, lam=0.0, verbose=True):
    """One dataset file in the reference schema.  For ``task == 'solver'`` the stored
    ``target_signal`` is ``K b + lam b`` so that ``true_answer = source_signal``
    (runner.py:87-90) stays the solution of the regularised system."""
    import h5py  # the real one, or the stand-in installed by bootstrap.install_import_shims

    t0 = time.time()
    args = dict(kernel=ds.kernel, source_points=ds.source_points, target_points=None if ds.same_points else ds.target_points,
                source_signal=ds.source_signal, normalize_rows=ds.normalize_rows)
    if float(ds.N) * ds.M * ds.D >= GPU_TRUTH_MIN_WORK:
        truth, truth_source = _ground_truth_gpu(**args)
    else:
        truth, truth_source = _ground_truth_blocked(**args), "the reference GroundTruth (datasets.py:180-195) on all rows"
    if ds.task == "solver" and lam:
        truth = truth + lam * ds.source_signal
    with h5py.File(filename, "w") as f:
        f.attrs["kernel"] = ds.kernel
        f.attrs["task"] = ds.task
        f.attrs["point_type"] = "float"
        f.attrs["normalize_rows"] = bool(ds.normalize_rows)
        f.attrs["short_description"] = f"{label} (N={ds.N}, D={ds.D})"
        f.attrs["description"] = f"{ds.task.capitalize()} on the cube, {ds.kernel} (N={ds.N}, M={ds.M}, D={ds.D}, E={ds.E})"
        f.attrs["same_points"] = bool(ds.same_points)
        f.attrs["density_estimation"] = bool(ds.density_estimation)
        if lam:
            f.attrs["lam"] = float(lam)
        f.attrs["truth_source"] = truth_source
        f["source_points"] = ds.source_points
        f["target_points"] = ds.target_points
        f["source_signal"] = ds.source_signal
        f["target_signal"] = truth
    if verbose:
        print(f"wrote {filename}: N={ds.N} M={ds.M} D={ds.D} E={ds.E} kernel={ds.kernel} "
              f"({time.time() - t0:.1f} s; ground truth: {truth_source})")


def _writer(make, label, lam=0.0):
    def write_to(filename):
        write_dataset(filename, make(), label=label, lam=lam)

    return write_to


def extra_datasets():
    """name -> writer(filename), the shape of the reference's ``DATASETS`` (datasets.py:416-427)."""
    out = {}
    # C1/C2 family: Gaussian product on the unit cube, x == y, E = 1
    for n in (1000, 10_000, 30_000, 100_000, 1_000_000):   # 10^4 = C1, 10^6 = C2
        out[gen.dataset_name("product", "ucube", 3, 1, n, n, "gaussian")] = _writer(
            lambda n=n: gen.uniform_cube(n, 3, 1.0, "gaussian", "product"), "ucube")
    for kernel in ("absolute-exponential", "inverse-distance"):
        n = 10_000
        out[gen.dataset_name("product", "ucube", 3, 1, n, n, kernel)] = _writer(
            lambda n=n, kernel=kernel: gen.uniform_cube(n, 3, 1.0, kernel, "product"), "ucube")
    # C3 family: MNIST-shaped, independent targets, radius sqrt(3/D)
    for m, n in ((4000, 1000), (60_000, 10_000)):   # the second is C3
        out[gen.dataset_name("product", "ucube", 784, 1, m, n, "gaussian")] = _writer(
            lambda m=m, n=n: gen.config_c3(m, n, 784), "ucube")
    # C4 family: row-normalised attention, E = 64
    for kernel in ("absolute-exponential", "gaussian"):
        for n in (4096, 262_144):   # the second is C4
            out[gen.dataset_name("attention", "ucube", 64, 64, n, n, kernel)] = _writer(
                lambda n=n, kernel=kernel: gen.config_c4(n, 64, 64, kernel), "ucube")
    # C5 family: (K + I) b = a
    for n in (2000, 10_000, 100_000, 1_000_000):   # 10^6 = C5
        out[gen.dataset_name("solver", "ucubelam1", 3, 1, n, n, "gaussian")] = _writer(
            lambda n=n: gen.uniform_cube(n, 3, 1.0, "gaussian", "solver"), "ucubelam1", lam=1.0)
    return out


def register():
    """Add the writers to the reference's ``DATASETS`` (idempotent)."""
    from kernel_matrix_benchmarks import datasets as ref_datasets  # the reference's

    for name, writer in extra_datasets().items():
        ref_datasets.DATASETS.setdefault(name, writer)
    return ref_datasets.DATASETS
