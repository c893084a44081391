"""Where Gram blocks are stored: ``open_store`` and the worker-file merge.

``*.h5`` / ``*.hdf5`` paths are real HDF5 files, the reference's format (``h5py.File`` at
exp_mnist_resnet/save_kernel.py:26,33; layout cnn_gp/kernel_save_tools.py:7-23): written and read
by this repository's native implementation of the file format (``cnn_gp.h5store`` ->
libcnngp_h5.so), because h5py and libhdf5 are absent from the B200 image; set ``CNNGP_USE_H5PY=1``
to go through an installed h5py instead -- the files are interchangeable.  Any other path is a
directory of ``.npy`` memmaps (``NpyStore``) that answers the same subset of the h5py API
(``create_dataset``, ``keys``, item access, slicing, ``read_direct``, context manager).
"""
import json
import os

import numpy as np


class NpyDataset:
    def __init__(self, path, meta, mode):
        self._arr = np.lib.format.open_memmap(path, mode=mode)
        self.chunks = tuple(meta["chunks"]) if meta.get("chunks") else None
        self.maxshape = tuple(meta["maxshape"]) if meta.get("maxshape") else None
        self.fillvalue = meta.get("fillvalue")

    shape = property(lambda self: self._arr.shape)
    dtype = property(lambda self: self._arr.dtype)

    def __len__(self):
        return self._arr.shape[0]

    def __getitem__(self, key):
        return np.asarray(self._arr[key])

    def __setitem__(self, key, value):
        self._arr[key] = value

    def read_direct(self, dest, source_sel=None, dest_sel=None):
        src = self._arr if source_sel is None else self._arr[source_sel]
        if dest_sel is None:
            dest[...] = src
        else:
            dest[dest_sel] = src

    def flush(self):
        self._arr.flush()


class NpyStore:
    def __init__(self, path, mode="r"):
        if mode not in ("r", "r+", "a", "w", "w-", "x"):
            raise ValueError(mode)
        self.path, self.mode = path, mode
        exists = os.path.isdir(path)
        if mode == "r" or mode == "r+":
            if not exists:
                raise FileNotFoundError(path)
        elif mode in ("w-", "x") and exists:
            raise FileExistsError(path)
        os.makedirs(path, exist_ok=True)
        if mode == "w":  # truncate, like h5py.File(path, "w")
            for fn in os.listdir(path):
                if fn.endswith(".npy") or fn == "meta.json":
                    os.remove(os.path.join(path, fn))
        self._meta_path = os.path.join(path, "meta.json")
        self._meta = {}
        if os.path.exists(self._meta_path):
            with open(self._meta_path) as fh:
                self._meta = json.load(fh)
        self._open = {}

    # -- h5py-like API ------------------------------------------------------------------
    def keys(self):
        return list(self._meta.keys())

    def __contains__(self, name):
        return name in self._meta

    def __iter__(self):
        return iter(self.keys())

    def __getitem__(self, name):
        if name not in self._meta:
            raise KeyError(name)
        if name not in self._open:
            self._open[name] = NpyDataset(self._file(name), self._meta[name],
                                          "r" if self.mode == "r" else "r+")
        return self._open[name]

    def create_dataset(self, name, shape, dtype=np.float32, fillvalue=None, chunks=None, maxshape=None,
                       data=None):
        if self.mode == "r":
            raise OSError("store opened read-only")
        if name in self._meta:
            raise ValueError(f"dataset {name} exists")
        arr = np.lib.format.open_memmap(self._file(name), mode="w+", dtype=np.dtype(dtype), shape=tuple(shape))
        if data is not None:
            arr[...] = data
        elif fillvalue is not None:
            arr[...] = fillvalue
        arr.flush()
        del arr
        fv = None if fillvalue is None else (None if isinstance(fillvalue, float) and np.isnan(fillvalue)
                                             else fillvalue)
        self._meta[name] = dict(chunks=list(chunks) if chunks else None,
                                maxshape=[m for m in maxshape] if maxshape else None,
                                fillvalue="nan" if fillvalue is not None and fv is None else fv)
        self._write_meta()
        return self[name]

    def flush(self):
        for ds in self._open.values():
            ds.flush()

    def close(self):
        self.flush()
        self._open.clear()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    # -- internals ----------------------------------------------------------------------
    def _file(self, name):
        return os.path.join(self.path, name + ".npy")

    def _write_meta(self):
        tmp = self._meta_path + ".tmp"
        with open(tmp, "w") as fh:
            json.dump(self._meta, fh)
        os.replace(tmp, self._meta_path)


def open_store(path, mode="r"):
    """An HDF5 file for ``*.h5`` / ``*.hdf5`` paths (native store; h5py on request), else ``NpyStore``."""
    path = str(path)
    if path.endswith((".h5", ".hdf5")):
        if os.environ.get("CNNGP_USE_H5PY") == "1":
            import h5py
            return h5py.File(path, mode)
        from . import h5store
        return h5store.File(path, mode)
    return NpyStore(path, mode)


def merge_into(dest, src):
    """exp_mnist_resnet/merge_h5_files.py:15-30: for every dataset present in both stores copy
    ``src`` into ``dest`` wherever ``dest`` is NaN, one leading index at a time."""
    from . import h5store
    for k in [k for k in dest.keys() if k in src.keys()]:
        d, s = dest[k], src[k]
        if isinstance(dest, h5store.File) and isinstance(src, h5store.File) and d.chunks is not None \
                and (d.shape, d.chunks, d.dtype) == (s.shape, s.chunks, s.dtype):
            # chunk by chunk in native code: cost proportional to what the workers wrote
            h5store.merge_nan(dest, src, k)
            continue
        for i in range(len(d)):
            block = d[i, ...]
            other = s[i, ...]
            todo = np.isnan(block)
            block[todo] = other[todo]
            d[i, ...] = block
