"""A small stand-in for the slice of ``h5py`` the reference harness uses.

``h5py`` is not installed in this image and cannot be fetched offline, but
the reference's ``datasets.py`` / ``results.py`` / ``plotting`` import it.  The
calls they make are few:

* ``h5py.File(fn, "r" | "w" | "r+")``, ``close()``, context manager
  (datasets.py:150,118; results.py:107,133)
* ``f[key] = array``; ``f[key][:]``; ``f[key].shape``; ``key in f``; ``del f[key]``
  (datasets.py:162-195; runner.py:31-36; plotting/utils.py:8-12,111-112)
* ``f.attrs[key]``, ``f.attrs.get``, ``dict(f.attrs)`` (main.py:172-174; results.py:116,134)
* ``f.create_group(name)`` with its own ``attrs`` (plotting/metrics.py:47-58)

``File`` below keeps one tree of ``{"attrs": {}, "items": {}}`` nodes in memory and
pickles it on ``close()`` when opened for writing.  The on-disk bytes are *not*
HDF5 -- this is a compatibility layer for running the unmodified harness
offline, installed only when the real ``h5py`` cannot be imported
(``bootstrap.install_import_shims``).
"""
from __future__ import annotations

import os
import pickle

import numpy as np


class _Attrs(dict):
    """``f.attrs``: a dict; numpy scalars are stored as Python scalars the way h5py returns them."""

    def __setitem__(self, key, value):
        if isinstance(value, np.generic):
            value = value.item()
        super().__setitem__(key, value)


class Dataset:
    """``f[key]``: supports ``[:]`` / any numpy index, ``.shape``, ``.dtype``, ``len``."""

    def __init__(self, array):
        self._a = np.asarray(array)

    def __getitem__(self, idx):
        out = self._a[idx]
        return out.copy() if isinstance(out, np.ndarray) else out

    def __len__(self):
        return len(self._a)

    def __array__(self, dtype=None, copy=None):
        return self._a if dtype is None else self._a.astype(dtype)

    @property
    def shape(self):
        return self._a.shape

    @property
    def dtype(self):
        return self._a.dtype


class Group:
    def __init__(self, node):
        self._node = node

    @property
    def attrs(self):
        return self._node["attrs"]

    def __contains__(self, key):
        return key in self._node["items"]

    def __getitem__(self, key):
        item = self._node["items"][key]
        return Group(item) if isinstance(item, dict) else Dataset(item)

    def __setitem__(self, key, value):
        self._require_writable()
        if isinstance(value, Dataset):
            value = value[:]
        self._node["items"][key] = np.array(value)

    def __delitem__(self, key):
        self._require_writable()
        del self._node["items"][key]

    def keys(self):
        return self._node["items"].keys()

    def create_group(self, name):
        self._require_writable()
        if name in self._node["items"]:
            raise ValueError(f"unable to create group (name already exists): {name}")
        self._node["items"][name] = _new_node()
        return Group(self._node["items"][name])

    def _require_writable(self):
        pass


def _new_node():
    return {"attrs": _Attrs(), "items": {}}


class File(Group):
    def __init__(self, name, mode="r"):
        if mode not in ("r", "r+", "w", "a"):
            raise ValueError(f"unsupported mode {mode!r}")
        self.filename, self.mode, self._open = str(name), mode, True
        if mode == "w" or (mode == "a" and not os.path.exists(name)):
            node = _new_node()
        else:
            with open(name, "rb") as fh:
                node = pickle.load(fh)
        super().__init__(node)

    def _require_writable(self):
        if self.mode == "r":
            raise OSError("file is open read-only")

    def flush(self):
        if self.mode != "r":
            tmp = self.filename + ".tmp"
            with open(tmp, "wb") as fh:
                pickle.dump(self._node, fh, protocol=pickle.HIGHEST_PROTOCOL)
            os.replace(tmp, self.filename)

    def close(self):
        if self._open:
            self.flush()
            self._open = False

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
