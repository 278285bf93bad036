"""ctypes front end of oracle/kprod_ref.c -- TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libkmb_oracle.so")
_KERNEL_IDS = {"gaussian": 0, "absolute-exponential": 1, "inverse-distance": 2}


def build(force=False):
    src = os.path.join(_HERE, "kprod_ref.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-s", "-C", _HERE, "-B"], check=True)
    return _LIB


_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.kmb_oracle_product_f64.restype = ctypes.c_int
        _lib.kmb_oracle_threads.restype = ctypes.c_int
    return _lib


def threads():
    return int(_load().kmb_oracle_threads())


def kernel_product(kernel, source_points, target_points, source_signal, *, normalize_rows=False,
                   density_estimation=False, rows=None):
    """Same contract as bruteforce_oracle.kernel_product (float64 only)."""
    lib = _load()
    if kernel not in _KERNEL_IDS:
        raise NotImplementedError(f"oracle: unsupported kernel {kernel!r}")
    y = np.ascontiguousarray(source_points, dtype=np.float64)
    x = y if target_points is None else np.ascontiguousarray(target_points, dtype=np.float64)
    M, D = y.shape
    ids = None
    if rows is not None:
        ids = np.ascontiguousarray(rows, dtype=np.int64)
        x = np.ascontiguousarray(x[ids])
    n = x.shape[0]
    if normalize_rows and density_estimation:
        return np.ones((n, 1))
    b = None if density_estimation else np.ascontiguousarray(source_signal, dtype=np.float64)
    E = 1 if b is None else b.shape[1]
    out = np.empty((n, E), dtype=np.float64)
    p = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
    rc = lib.kmb_oracle_product_f64(
        ctypes.c_int(_KERNEL_IDS[kernel]), p(x), p(ids), p(y), p(b), p(out),
        ctypes.c_int64(n), ctypes.c_int64(M), ctypes.c_int(D), ctypes.c_int(E), ctypes.c_int(int(normalize_rows)))
    if rc != 0:
        raise ValueError("kmb_oracle_product_f64: bad arguments")
    return out
