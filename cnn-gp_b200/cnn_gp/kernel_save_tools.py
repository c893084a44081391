"""Writing Gram matrices block by block, API-compatible with the reference's
cnn_gp/kernel_save_tools.py.

Layout contract (reference kernel_save_tools.py:7-23): dataset ``name`` has shape ``(1, N, N2)``
(``(1, N)`` for diagonals), float32, fill value NaN, chunks ``(1, min(bs, N), min(bs, N2))`` and
an unlimited leading dimension.  NaN therefore marks "block not computed by this worker", which
is what the merge step relies on (exp_mnist_resnet/merge_h5_files.py:27-28).

``f`` may be an ``h5py.File`` or a ``cnn_gp.block_store.NpyStore`` (same calls, ``.npy`` memmaps;
h5py / libhdf5 are not part of this image).
"""
import numpy as np

from .data import ProductIterator, DiagIterator, print_timings

__all__ = ('create_h5py_dataset', 'save_K')


def create_h5py_dataset(f, batch_size, name, diag, N, N2):
    """Create dataset ``name`` on ``f`` with ``batch_size`` chunks and a resizable leading
    dimension of length 1."""
    tail = (N,) if diag else (N, N2)
    chunk_tail = (min(batch_size, N),) if diag else (min(batch_size, N), min(batch_size, N2))
    return f.create_dataset(name, shape=(1,) + tail, dtype=np.float32, fillvalue=np.nan,
                            chunks=(1,) + chunk_tail, maxshape=(None,) + tail)


def save_K(f, kern, name, X, X2, diag, batch_size, worker_rank=0, n_workers=1,
           print_interval=2.):
    """Compute this worker's tiles of the kernel between ``X`` and ``X2`` (``X`` itself when
    ``X2 is None``) with ``kern(x, x2, same, diag) -> np.ndarray`` and store them in ``f[name]``.
    An existing dataset is left untouched."""
    if name in f.keys():
        print("Skipping {} (group exists)".format(name))
        return
    N = len(X)
    N2 = N if X2 is None else len(X2)
    out = create_h5py_dataset(f, batch_size, name, diag, N, N2)

    if diag:  # cheap: never split across workers
        tiles = DiagIterator(batch_size, X, X2)
    else:
        tiles = ProductIterator(batch_size, X, X2, worker_rank=worker_rank, n_workers=n_workers)
    tiles = print_timings(tiles, desc=f"{name} (worker {worker_rank}/{n_workers})",
                          print_interval=print_interval)

    for same, (i, (x, _y)), (j, (x2, _y2)) in tiles:
        k = kern(x, x2, same, diag)
        if not np.all(np.isfinite(k)):
            # the reference drops into ipdb here (kernel_save_tools.py:51-53)
            raise FloatingPointError(f"nan or inf in kernel block {name}[{i},{j}]")
        if diag:
            out[0, i:i + len(x)] = k
        else:
            out[0, i:i + len(x), j:j + len(x2)] = k
