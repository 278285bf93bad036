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
writes it on ``close()`` when opened for writing.  The on-disk bytes are *not*
HDF5 -- this is a compatibility layer for running the unmodified harness
offline, installed only when the real ``h5py`` cannot be imported
(``bootstrap.install_import_shims``).

On-disk format: the 16-byte magic ``MAGIC`` followed by an ``.npz`` archive (``numpy.savez``) that holds
every array under ``a<k>`` and one member ``tree`` with the group structure and attributes as JSON.  It
is read back with ``numpy.load(allow_pickle=False)``: nothing in a file is ever unpickled or executed, a
file without the magic (in particular a real HDF5 file, or anything a download may have put there) is
rejected with a clear error.
"""
from __future__ import annotations

import io
import json
import os

import numpy as np

MAGIC = b"KMB-H5LITE-v2\n\0\0"
HDF5_SIGNATURE = b"\x89HDF\r\n\x1a\n"
assert len(MAGIC) == 16


def _encode_attr(value):
    if isinstance(value, np.generic):
        value = value.item()
    if isinstance(value, np.ndarray):
        return {"__ndarray__": value.tolist(), "dtype": value.dtype.str}
    if isinstance(value, bytes):
        return {"__bytes__": value.decode("latin-1")}
    if isinstance(value, (list, tuple)):
        return [_encode_attr(v) for v in value]
    if value is None or isinstance(value, (bool, int, float, str)):
        return value
    raise TypeError(f"h5lite: cannot store an attribute of type {type(value).__name__}")


def _decode_attr(value):
    if isinstance(value, dict):
        if "__ndarray__" in value:
            return np.array(value["__ndarray__"], dtype=np.dtype(value["dtype"]))
        if "__bytes__" in value:
            return value["__bytes__"].encode("latin-1")
        raise ValueError("h5lite: malformed attribute")
    if isinstance(value, list):
        return [_decode_attr(v) for v in value]
    return value


def _dump(node, arrays):
    """Tree -> JSON-able structure; arrays are moved to ``arrays`` and referenced by key."""
    items = {}
    for name, item in node["items"].items():
        if isinstance(item, dict):
            items[name] = {"group": _dump(item, arrays)}
        else:
            a = np.asarray(item)
            if a.dtype == object:
                raise TypeError(f"h5lite: cannot store an object array ({name})")
            key = f"a{len(arrays)}"
            arrays[key] = a
            items[name] = {"array": key}
    return {"attrs": {k: _encode_attr(v) for k, v in node["attrs"].items()}, "items": items}


def _restore(desc, arrays):
    node = _new_node()
    for k, v in desc["attrs"].items():
        node["attrs"][k] = _decode_attr(v)
    for name, item in desc["items"].items():
        node["items"][name] = _restore(item["group"], arrays) if "group" in item else arrays[item["array"]]
    return node


def _read_file(name):
    with open(name, "rb") as fh:
        head = fh.read(16)
        if head[:8] == HDF5_SIGNATURE:
            raise OSError(f"{name} is a real HDF5 file; the offline h5py stand-in cannot read it (install h5py)")
        if head != MAGIC:
            raise OSError(f"{name} is not an h5lite file (bad magic); refusing to parse it")
        payload = io.BytesIO(fh.read())
    with np.load(payload, allow_pickle=False) as z:
        arrays = {k: z[k] for k in z.files if k != "tree"}
        desc = json.loads(bytes(z["tree"]).decode("utf-8"))
    return _restore(desc, arrays)


def _write_file(name, node):
    arrays = {}
    desc = _dump(node, arrays)
    arrays["tree"] = np.frombuffer(json.dumps(desc).encode("utf-8"), dtype=np.uint8)
    tmp = name + ".tmp"
    with open(tmp, "wb") as fh:
        fh.write(MAGIC)
        np.savez(fh, **arrays)
    os.replace(tmp, name)


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
            node = _read_file(name)
        super().__init__(node)

    def _require_writable(self):
        if self.mode == "r":
            raise OSError("file is open read-only")

    def flush(self):
        if self.mode != "r":
            _write_file(self.filename, self._node)

    def close(self):
        if self._open:
            self.flush()
            self._open = False

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
