"""Run the *unmodified* reference harness offline, with the B200 plugin registered.

The reference is a benchmark driver (``run.py`` -> ``main.main`` -> ``runner.run`` ->
``results.store_result`` -> ``plotting.metrics``); this repository replaces only the
arithmetic plugin underneath it.  To show the plugin is a drop-in, the harness itself has
to run -- but in this image it cannot even be imported: ``h5py``, ``docker`` and ``colors``
are missing and there is no network.  ``activate()`` therefore

1. finds a reference tree (``$KMB_REFERENCE``, ``baseline/_ref`` inside this repo, or
   ``/root/reference``) and puts it on ``sys.path`` -- nothing from it is copied into git;
2. installs stand-ins for the three missing imports **only if the real module is absent**:
   ``h5py`` -> ``h5lite`` (this package), ``docker`` / ``colors`` -> empty modules (``--local``
   never touches docker, runner.py:242-316; ``colors.color`` is only used to tint docker logs,
   runner.py:319-338);
3. adds the dataset writers of ``datasets_ext`` to ``datasets.DATASETS`` *before*
   ``main()`` builds its argument parser (main.py:86 reads the dict's keys at call time).

``run_main(argv)`` then calls the reference's own ``main()`` from the reference tree's
directory (``logging.conf``, ``data/`` and ``results/`` are cwd-relative: main.py:165,
datasets.py:96-98, results.py:79).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
STAGED_REFERENCE = os.path.join(REPO, "baseline", "_ref")


def find_reference(explicit=None):
    """Directory that contains the reference's ``kernel_matrix_benchmarks`` package and ``run.py``."""
    candidates = [explicit, os.environ.get("KMB_REFERENCE"), STAGED_REFERENCE, "/root/reference"]
    for c in candidates:
        if c and os.path.isfile(os.path.join(c, "kernel_matrix_benchmarks", "runner.py")):
            return os.path.abspath(c)
    return None


def stage_reference(src="/root/reference", dst=STAGED_REFERENCE):
    """Copy the reference tree to the git-ignored ``baseline/_ref`` so that it travels to the GPU
    box with the repository snapshot (gpurun ships /root/repo only).  Returns the staged path, or
    None when there is no source tree."""
    import shutil

    if not os.path.isfile(os.path.join(src, "kernel_matrix_benchmarks", "runner.py")):
        return None
    keep = {}
    for sub in ("data", "results"):  # generated datasets / results survive a re-stage
        p = os.path.join(dst, sub)
        if os.path.isdir(p):
            keep[sub] = p + ".keep"
            os.replace(p, keep[sub])
    for name in os.listdir(src):
        if name in (".git", "data", "results"):
            continue
        s, d = os.path.join(src, name), os.path.join(dst, name)
        if os.path.isdir(s):
            shutil.copytree(s, d, dirs_exist_ok=True, ignore=shutil.ignore_patterns("__pycache__"))
        else:
            os.makedirs(dst, exist_ok=True)
            shutil.copy2(s, d)
    for sub, p in keep.items():
        os.replace(p, os.path.join(dst, sub))
    return dst


def _module_missing(name):
    if name in sys.modules:
        return False
    try:
        return importlib.util.find_spec(name) is None
    except (ImportError, ValueError):
        return True


def install_import_shims():
    """Stand-ins for imports the reference needs but this image lacks.  Returns what was shimmed."""
    shimmed = []
    if _module_missing("h5py"):
        from . import h5lite

        mod = types.ModuleType("h5py")
        mod.File, mod.Group, mod.Dataset = h5lite.File, h5lite.Group, h5lite.Dataset
        mod.__doc__ = "offline stand-in installed by kernel_matrix_benchmarks_b200.harness (see h5lite.py)"
        sys.modules["h5py"] = mod
        shimmed.append("h5py")
    if _module_missing("docker"):
        mod = types.ModuleType("docker")

        def from_env(*a, **k):
            raise RuntimeError("docker is not available offline: run the harness with --local")

        mod.from_env = from_env
        sys.modules["docker"] = mod
        shimmed.append("docker")
    if _module_missing("colors"):
        mod = types.ModuleType("colors")
        mod.color = lambda s, *a, **k: s
        sys.modules["colors"] = mod
        shimmed.append("colors")
    return shimmed


def activate(reference=None, extra_datasets=True):
    """Make ``import kernel_matrix_benchmarks`` resolve to the reference and register the extra
    datasets.  Returns ``(reference_root, shimmed_module_names)``."""
    root = find_reference(reference)
    if root is None:
        raise RuntimeError(
            "reference harness not found: set KMB_REFERENCE or stage it with "
            "`python tools/run_harness.py --stage` (copies /root/reference to baseline/_ref)")
    if root not in sys.path:
        sys.path.insert(0, root)
    if REPO not in sys.path:
        sys.path.insert(1, REPO)
    shimmed = install_import_shims()
    if "h5py" in shimmed or os.environ.get("KMB_OFFLINE", "1") != "0":
        # get_dataset (datasets.py:106-109) first fetches http://kernel-matrix-benchmarks.com/datasets/<name>.hdf5 over
        # plain HTTP and opens whatever arrived.  The stand-in cannot read real HDF5 files and this harness is the offline
        # one: refuse the download so that get_dataset falls through to creating the dataset locally (:113-117).
        import kernel_matrix_benchmarks.datasets as ref_datasets

        def _refuse_download(src, dst):
            if not os.path.exists(dst):
                raise OSError(f"offline harness: not downloading {src}")

        ref_datasets.download = _refuse_download
    if extra_datasets:
        from . import datasets_ext

        datasets_ext.register()
    return root, shimmed


def run_main(argv, reference=None):
    """``python run.py <argv>`` of the reference, in-process (main.py:74-308)."""
    root, _ = activate(reference)
    from kernel_matrix_benchmarks.main import main  # the reference's

    old_argv, old_cwd = sys.argv, os.getcwd()
    sys.argv = ["run.py"] + list(argv)
    os.chdir(root)
    try:
        main()
    finally:
        sys.argv = old_argv
        os.chdir(old_cwd)
    return root
