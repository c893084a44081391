"""h5py-shaped front end of the native HDF5 block store (libcnngp_h5.so, include/cnngp_h5.h).

``File`` / ``Dataset`` answer the h5py calls the cnn-gp scripts make --
``h5py.File(path, mode)`` as a context manager, ``f.keys()``, ``name in f``, ``f[name]``,
``f.create_dataset(name, shape=, dtype=, fillvalue=, chunks=, maxshape=)``
(reference cnn_gp/kernel_save_tools.py:21-23), ``dset[0, i:i+n, j:j+m] = k`` (:55-58),
``dset[i, ...]`` (exp_mnist_resnet/merge_h5_files.py:24-30), ``dset.read_direct(A, source_sel=...)``
(exp_mnist_resnet/classify_gp.py:45-48), ``dset.resize`` -- on real HDF5 files, written and read
by this repository's own C++ implementation of the file format (h5py / libhdf5 are absent from
the B200 image).  Selections are integers, unit-step slices and ``...``.
"""
import ctypes
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_PKG), "libcnngp_h5.so")
MAX_RANK = 8
UNLIMITED = -1


class Info(ctypes.Structure):
    """struct cnngp_h5_info of include/cnngp_h5.h."""
    _fields_ = [("rank", ctypes.c_int32), ("dtype", ctypes.c_int32), ("chunked", ctypes.c_int32),
                ("has_fill", ctypes.c_int32), ("shape", ctypes.c_int64 * MAX_RANK),
                ("maxshape", ctypes.c_int64 * MAX_RANK), ("chunks", ctypes.c_int64 * MAX_RANK),
                ("fill", ctypes.c_double), ("n_chunks_stored", ctypes.c_int64)]


_I64P = ctypes.POINTER(ctypes.c_int64)
_SIGNATURES = {
    "cnngp_h5_last_error": (ctypes.c_char_p, []),
    "cnngp_h5_open": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]),
    "cnngp_h5_flush": (ctypes.c_int, [ctypes.c_void_p]),
    "cnngp_h5_close": (ctypes.c_int, [ctypes.c_void_p]),
    "cnngp_h5_count": (ctypes.c_int, [ctypes.c_void_p]),
    "cnngp_h5_name": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int]),
    "cnngp_h5_find": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p]),
    "cnngp_h5_create_dataset": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, _I64P, _I64P, _I64P,
                                               ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]),
    "cnngp_h5_dataset_info": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(Info)]),
    "cnngp_h5_write": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _I64P, _I64P, ctypes.c_void_p]),
    "cnngp_h5_read": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _I64P, _I64P, ctypes.c_void_p]),
    "cnngp_h5_write_strided": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _I64P, _I64P, ctypes.c_void_p, _I64P]),
    "cnngp_h5_read_strided": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _I64P, _I64P, ctypes.c_void_p, _I64P]),
    "cnngp_h5_resize": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _I64P]),
    "cnngp_h5_merge_nan": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]),
}
EXPORTS = tuple(_SIGNATURES)
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not found: build it with `python cnn-gp_b200/build.py`")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _check(rc, what):
    if rc != 0:
        raise OSError(f"{what}: {lib().cnngp_h5_last_error().decode('utf-8', 'replace')}")


def _i64(values):
    return (ctypes.c_int64 * max(1, len(values)))(*[int(v) for v in values])


_DTYPES = {0: np.dtype("<f4"), 1: np.dtype("<f8")}


class Dataset:
    def __init__(self, file, ds_id, name):
        self._f, self._id, self.name = file, ds_id, "/" + name

    def _info(self):
        info = Info()
        _check(lib().cnngp_h5_dataset_info(self._f._handle(), self._id, ctypes.byref(info)), "dataset info")
        return info

    @property
    def shape(self):
        i = self._info()
        return tuple(i.shape[:i.rank])

    @property
    def maxshape(self):
        i = self._info()
        return tuple(None if m == UNLIMITED else m for m in i.maxshape[:i.rank])

    @property
    def chunks(self):
        i = self._info()
        return tuple(i.chunks[:i.rank]) if i.chunked else None

    @property
    def dtype(self):
        i = self._info()
        if i.dtype not in _DTYPES:
            raise TypeError(f"{self.name}: element type not supported by the native HDF5 store")
        return _DTYPES[i.dtype]

    @property
    def fillvalue(self):
        i = self._info()
        return self.dtype.type(i.fill if i.has_fill else 0)

    @property
    def n_chunks_stored(self):
        return int(self._info().n_chunks_stored)

    ndim = property(lambda self: len(self.shape))
    size = property(lambda self: int(np.prod(self.shape)))

    def __len__(self):
        return self.shape[0]

    def _select(self, key):
        """key -> (start, count, result shape, which dimensions the result keeps)."""
        shape = self.shape
        if not isinstance(key, tuple):
            key = (key,)
        if sum(1 for k in key if k is Ellipsis) > 1:
            raise IndexError("more than one Ellipsis")
        if any(k is Ellipsis for k in key):
            e = next(i for i, k in enumerate(key) if k is Ellipsis)
            key = key[:e] + (slice(None),) * (len(shape) - len(key) + 1) + key[e + 1:]
        if len(key) > len(shape):
            raise IndexError("too many indices")
        key = key + (slice(None),) * (len(shape) - len(key))
        start, count, out = [], [], []
        kept = []  # which dataset dimensions survive in the result (slices do, integers do not)
        for k, n in zip(key, shape):
            if isinstance(k, (int, np.integer)):
                k = int(k)
                if k < 0:
                    k += n
                if not 0 <= k < n:
                    raise IndexError(f"index {k} out of range for extent {n}")
                start.append(k)
                count.append(1)
                kept.append(False)
            elif isinstance(k, slice):
                a, b, s = k.indices(n)
                if s != 1:
                    raise IndexError("only unit-step slices are supported")
                start.append(a)
                count.append(max(0, b - a))
                out.append(max(0, b - a))
                kept.append(True)
            else:
                raise IndexError(f"unsupported index {k!r}")
        return start, count, tuple(out), kept

    def _strides_of(self, arr, kept, out_shape):
        """Byte strides per dataset dimension if ``arr`` (shape ``out_shape``) can be handed to the
        library as it is -- right dtype, last dimension contiguous, no negative strides -- else None."""
        if not isinstance(arr, np.ndarray) or arr.dtype != self.dtype or arr.shape != out_shape or arr.ndim == 0:
            return None
        if arr.strides[-1] != arr.itemsize or any(st < 0 for st in arr.strides) or not kept[-1]:
            return None
        it = iter(arr.strides)  # dimensions indexed by an integer have no axis in arr: stride 0
        return [next(it) if k else 0 for k in kept]

    def __getitem__(self, key):
        start, count, out_shape, _ = self._select(key)
        arr = np.empty(count, dtype=self.dtype)
        if arr.size:
            _check(lib().cnngp_h5_read(self._f._handle(), self._id, _i64(start), _i64(count),
                                       arr.ctypes.data_as(ctypes.c_void_p)), f"read {self.name}")
        arr = arr.reshape(out_shape)
        return arr[()] if arr.ndim == 0 else arr

    def __setitem__(self, key, value):
        start, count, out_shape, kept = self._select(key)
        strides = self._strides_of(value, kept, out_shape)
        if strides is not None and value.size:  # a strided view (e.g. columns of a row buffer): no packing copy
            _check(lib().cnngp_h5_write_strided(self._f._handle(), self._id, _i64(start), _i64(count),
                                                value.ctypes.data_as(ctypes.c_void_p), _i64(strides)), f"write {self.name}")
            return
        arr = np.ascontiguousarray(np.broadcast_to(np.asarray(value, dtype=self.dtype), out_shape)).reshape(count)
        if arr.size:
            _check(lib().cnngp_h5_write(self._f._handle(), self._id, _i64(start), _i64(count),
                                        arr.ctypes.data_as(ctypes.c_void_p)), f"write {self.name}")

    def read_direct(self, dest, source_sel=None, dest_sel=None):
        """Read straight into ``dest`` (C-contiguous destination selections avoid the copy)."""
        start, count, out_shape, kept = self._select(source_sel if source_sel is not None else Ellipsis)
        target = dest if dest_sel is None else dest[dest_sel]
        if target.shape != out_shape:
            raise TypeError(f"cannot read {out_shape} into {target.shape}")
        strides = self._strides_of(target, kept, out_shape) if target.flags.writeable else None
        if strides is not None:
            if target.size:
                _check(lib().cnngp_h5_read_strided(self._f._handle(), self._id, _i64(start), _i64(count),
                                                   target.ctypes.data_as(ctypes.c_void_p), _i64(strides)),
                       f"read {self.name}")
        else:
            target[...] = self[source_sel if source_sel is not None else Ellipsis]

    def resize(self, size, axis=None):
        shape = list(self.shape)
        if axis is not None:
            shape[axis] = int(size)
        else:
            shape = [int(s) for s in size]
        _check(lib().cnngp_h5_resize(self._f._handle(), self._id, _i64(shape)), f"resize {self.name}")

    def flush(self):
        self._f.flush()


class File:
    """``h5py.File(path, mode)`` on the native store.  Modes: r, r+, w, w- / x, a."""

    def __init__(self, path, mode="r"):
        self.filename, self.mode = str(path), mode
        h = ctypes.c_void_p()
        rc = lib().cnngp_h5_open(os.fsencode(self.filename), mode.encode(), ctypes.byref(h))
        if rc != 0:
            msg = lib().cnngp_h5_last_error().decode("utf-8", "replace")
            if mode in ("r", "r+") and not os.path.exists(self.filename):
                raise FileNotFoundError(msg)
            if mode in ("w-", "x") and os.path.exists(self.filename):
                raise FileExistsError(msg)
            raise OSError(msg)
        self._h = h
        self._open = {}

    def _handle(self):
        if self._h is None:
            raise ValueError("file is closed")
        return self._h

    def keys(self):
        L, h = lib(), self._handle()
        names = []
        for i in range(L.cnngp_h5_count(h)):
            n = L.cnngp_h5_name(h, i, None, 0)
            buf = ctypes.create_string_buffer(n)
            L.cnngp_h5_name(h, i, buf, n)
            names.append(buf.value.decode())
        return names

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return lib().cnngp_h5_count(self._handle())

    def __contains__(self, name):
        return lib().cnngp_h5_find(self._handle(), name.lstrip("/").encode()) >= 0

    def __getitem__(self, name):
        name = name.lstrip("/")
        ds = lib().cnngp_h5_find(self._handle(), name.encode())
        if ds < 0:
            raise KeyError(f"unable to open object (object '{name}' doesn't exist)")
        if name not in self._open:
            self._open[name] = Dataset(self, ds, name)
        return self._open[name]

    def create_dataset(self, name, shape=None, dtype=None, data=None, fillvalue=None, chunks=None, maxshape=None):
        if data is not None:
            data = np.asarray(data)
            shape = data.shape if shape is None else shape
            dtype = data.dtype if dtype is None else dtype
        if shape is None:
            raise TypeError("one of data, shape is required")
        shape = (int(shape),) if isinstance(shape, (int, np.integer)) else tuple(int(s) for s in shape)
        dt = np.dtype(np.float32 if dtype is None else dtype)
        code = {np.dtype("float32"): 0, np.dtype("float64"): 1}.get(dt.newbyteorder("="))
        if code is None:
            raise TypeError(f"the native HDF5 store holds float32 / float64 datasets, not {dt}")
        if maxshape is not None:
            maxshape = tuple(UNLIMITED if m is None else int(m) for m in maxshape)
            if chunks is None or chunks is True:  # resizable datasets are chunked: pick ~1 MiB chunks
                chunks = self._guess_chunks(shape, maxshape, dt.itemsize)
        if chunks is True:
            chunks = self._guess_chunks(shape, shape, dt.itemsize)
        fill = None
        if fillvalue is not None:
            fill = np.array([fillvalue], dtype=_DTYPES[code])
        ds = ctypes.c_int(-1)
        _check(lib().cnngp_h5_create_dataset(
            self._handle(), name.lstrip("/").encode(), len(shape), _i64(shape),
            _i64(maxshape) if maxshape is not None else None,
            _i64(chunks) if chunks is not None else None, code,
            fill.ctypes.data_as(ctypes.c_void_p) if fill is not None else None, ctypes.byref(ds)),
            f"create_dataset {name}")
        d = self[name]
        if data is not None:
            d[...] = data
        return d

    @staticmethod
    def _guess_chunks(shape, maxshape, itemsize):
        dims = [max(1, s if m == UNLIMITED or m is None else m) for s, m in zip(shape, maxshape)]
        chunk = list(dims)
        i = 0
        while np.prod(chunk) * itemsize > (1 << 20) and any(c > 1 for c in chunk):
            if chunk[i % len(chunk)] > 1:
                chunk[i % len(chunk)] = (chunk[i % len(chunk)] + 1) // 2
            i += 1
        return tuple(int(c) for c in chunk)

    def flush(self):
        _check(lib().cnngp_h5_flush(self._handle()), "flush")

    def close(self):
        if self._h is not None:
            h, self._h = self._h, None
            self._open.clear()
            _check(lib().cnngp_h5_close(h), "close")

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def merge_nan(dest, src, name):
    """exp_mnist_resnet/merge_h5_files.py:24-30 for one dataset, natively and chunk by chunk."""
    _check(lib().cnngp_h5_merge_nan(dest._handle(), dest[name]._id, src._handle(), src[name]._id), f"merge {name}")


def _main(argv):
    """``python -m cnn_gp.h5store FILE...``: list the datasets of HDF5 files (an ``h5ls`` for this store)."""
    if not argv:
        print("usage: python -m cnn_gp.h5store FILE...")
        return 1
    for path in argv:
        with File(path, "r") as f:
            print(f"{path}: {len(f)} dataset(s)")
            for name in f.keys():
                d = f[name]
                info = d._info()
                if info.dtype not in _DTYPES:
                    print(f"  {name:12s} shape {d.shape}  (element type not read by this store)")
                    continue
                layout = f"chunks {d.chunks}, {d.n_chunks_stored} stored" if d.chunks else "contiguous"
                fill = f", fill {d.fillvalue}" if info.has_fill else ""
                print(f"  {name:12s} shape {d.shape}  maxshape {d.maxshape}  {d.dtype}  {layout}{fill}")
    return 0


if __name__ == "__main__":
    import sys
    sys.exit(_main(sys.argv[1:]))
