"""Writing Gram matrices block by block, API-compatible with the reference's
cnn_gp/kernel_save_tools.py.

Layout contract (reference kernel_save_tools.py:7-23): dataset ``name`` has shape ``(1, N, N2)``
(``(1, N)`` for diagonals), float32, fill value NaN, chunks ``(1, min(bs, N), min(bs, N2))`` and
an unlimited leading dimension.  NaN therefore marks "block not computed by this worker", which
is what the merge step relies on (exp_mnist_resnet/merge_h5_files.py:27-28).

``f`` may be a ``cnn_gp.h5store.File`` (real HDF5 files written and read by this repository's own
implementation of the format, csrc/h5store.cpp: h5py / libhdf5 are not part of this image), an
``h5py.File`` where one is installed, or a ``cnn_gp.block_store.NpyStore`` (same calls, a directory
of ``.npy`` memmaps); ``cnn_gp.block_store.open_store`` picks by path.
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


def _images_of(dataset, device):
    import torch
    from .data import _batch
    x = dataset.images if hasattr(dataset, "images") and torch.is_tensor(dataset.images) else \
        _batch(dataset, 0, len(dataset))[0]
    return x.to(device, non_blocking=True).contiguous()


def save_K_resident(f, model, name, X, X2, diag, batch_size, worker_rank=0, n_workers=1,
                    print_interval=2., device=None):
    """``save_K`` with the datasets resident in HBM: same dataset layout, same tile ownership
    (the reference's contiguous slice of its tile list, cnn_gp/data.py:11-29, 54-60) and the same
    values, but the images are uploaded once, the per-image variance maps are computed once per
    dataset, every block row is at most two kernel launches, and finished rows stream to the
    store through double-buffered pinned memory while the next row is being computed.

    Replaces the per-tile H2D / D2H round trips of exp_mnist_resnet/save_kernel.py:21-24."""
    import torch
    from .tiles import GramJob, row_segments
    from .data import worker_tiles
    if name in f.keys():
        print("Skipping {} (group exists)".format(name))
        return
    if device is None:
        bufs = list(model.buffers())
        device = bufs[0].device if bufs else torch.device("cuda")
    N = len(X)
    N2 = N if X2 is None else len(X2)
    out = create_h5py_dataset(f, batch_size, name, diag, N, N2)
    with torch.cuda.device(device):
        x = _images_of(X, device)
        if diag:
            # DiagIterator semantics (data.py:99-126): index-aligned, truncated to the shorter set
            n = N if X2 is None else min(N, N2)
            x2 = x if X2 is None else _images_of(X2, device)
            k = model(x[:n], x2[:n], same=(X2 is None), diag=True)
            if not bool(torch.isfinite(k).all()):
                raise FloatingPointError(f"nan or inf in kernel diagonal {name}")
            out[0, 0:n] = k.cpu().numpy()
            return
        job = GramJob(model, x, None if X2 is None else _images_of(X2, device))
        tiles = worker_tiles(N, None if X2 is None else N2, batch_size, worker_rank, n_workers)
        segs = row_segments(tiles)
        copy_stream = torch.cuda.Stream()
        # two device row buffers and two pinned host buffers, allocated once and used alternately:
        # row k is copied out and written to the store while row k + 1 is being computed
        # (flat: every row segment is viewed as a CONTIGUOUS [rows, cols] matrix, so the device-to-host
        # copy is one plain async memcpy -- torch stages non-contiguous cross-device copies through
        # pageable memory and blocks)
        # device rows in the images' dtype (a .double() dataset runs the float64 kernel); the store is
        # float32 like the reference's (kernel_save_tools.py:21: the numpy tile is cast on assignment), so
        # a float64 row is narrowed on the device before it leaves
        dev_bufs = [torch.empty(batch_size * N2, dtype=x.dtype, device=device) for _ in range(2)]
        narrow = x.dtype != torch.float32
        dev32 = [torch.empty(batch_size * N2, dtype=torch.float32, device=device) for _ in range(2)] if narrow else None
        host_bufs = [torch.empty(batch_size * N2, dtype=torch.float32).pin_memory() for _ in range(2)]
        flags = torch.ones(2, dtype=torch.bool).pin_memory()  # per buffer: "every entry of the row is finite"
        pending = []  # (event, buffer index, host view, i0, i1, j0, j1)

        import os
        import time
        spent = {"wait_gpu": 0.0, "write": 0.0, "t0": time.perf_counter()}

        def drain(keep):
            while len(pending) > keep:
                ev, b, host, i0, i1, j0, j1 = pending.pop(0)
                t = time.perf_counter()
                ev.synchronize()
                spent["wait_gpu"] += time.perf_counter() - t
                if not bool(flags[b]):
                    raise FloatingPointError(f"nan or inf in kernel block row {name}[{i0}:{i1}]")
                t = time.perf_counter()
                out[0, i0:i1, j0:j1] = host.numpy()
                spent["write"] += time.perf_counter() - t

        segs = print_timings(segs, desc=f"{name} rows (worker {worker_rank}/{n_workers})",
                             print_interval=print_interval)
        for k, (r, has_diag, c0, c1) in enumerate(segs):
            i0, i1 = r * batch_size, min(N, (r + 1) * batch_size)
            j0 = i0 if has_diag else c0 * batch_size
            j1 = min(N2, c1 * batch_size) if c0 is not None else i1
            # buffer k % 2 was last used by row k - 2, which the drain of the previous iteration wrote
            rows, cols = i1 - i0, j1 - j0
            buf = dev_bufs[k % 2][:rows * cols].view(rows, cols)
            if has_diag:
                job.block_into(buf[:, :i1 - i0], i0, i1, i0, i1, symmetric=True)
            if c0 is not None:
                js = c0 * batch_size
                job.block_into(buf[:, js - j0:], i0, i1, js, j1, symmetric=False)
            finite = torch.isfinite(buf).all()  # stays on the device: no host sync in this loop
            if narrow:
                src = dev32[k % 2][:rows * cols].view(rows, cols)
                src.copy_(buf)
            else:
                src = buf
            host = host_bufs[k % 2][:rows * cols].view(rows, cols)
            copy_stream.wait_stream(torch.cuda.current_stream())
            finite.record_stream(copy_stream)  # read there: the allocator must not hand it out again before
            with torch.cuda.stream(copy_stream):
                host.copy_(src, non_blocking=True)
                flags[k % 2:k % 2 + 1].copy_(finite.reshape(1), non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            pending.append((ev, k % 2, host, i0, i1, j0, j1))
            # row k is queued on the GPU: write row k - 1 to the store underneath it
            drain(keep=1)
        drain(keep=0)
        if os.environ.get("CNNGP_SAVE_TIMING"):
            print(f"{name}: {time.perf_counter() - spent['t0']:.2f} s in the row loop, of which "
                  f"{spent['wait_gpu']:.2f} s waiting for the GPU and {spent['write']:.2f} s writing to the store")
