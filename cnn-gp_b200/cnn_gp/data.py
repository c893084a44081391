"""Tile iteration for Gram matrices, API-compatible with the reference's cnn_gp/data.py.

``ProductIterator`` / ``DiagIterator`` yield ``(same, (i, (x, labels)), (j, (x2, labels2)))``
with ``i``, ``j`` the first row / column of the tile, enumerate tiles in the reference's order
(upper triangle incl. diagonal when ``X2 is None``, data.py:22-29) and hand worker ``r`` the
same contiguous slice of that list (data.py:11-19, 54-60), so block ownership in the output
file is identical.

Datasets that expose their images as one tensor (``.images`` / ``.labels``, e.g.
``ResidentDataset`` below, typically already on the GPU) are sliced directly; any other
map-style dataset is batched item by item like the reference's ``DataLoader(Subset(...))``.
"""
import itertools
import os
import time

import torch
from torch.utils.data import ConcatDataset, Dataset, Subset
from torch.utils.data.dataloader import default_collate

__all__ = ('DatasetFromConfig', 'ProductIterator', 'DiagIterator',
           'print_timings')


def _round_up_div(a, b):
    return -(-a // b)


def _this_worker_batch(N_batches, worker_rank, n_workers):
    """(first tile, number of tiles) of ``worker_rank``: an even contiguous split where the
    first ``N_batches % n_workers`` workers take one extra tile."""
    base, extra = divmod(N_batches, n_workers)
    start = worker_rank * base + min(worker_rank, extra)
    return int(start), int(base + (1 if worker_rank < extra else 0))


def _product_generator(N_batches_X, N_batches_X2, same):
    for i in range(N_batches_X):
        first_col = 0
        if same:
            yield (True, i, i)  # diagonal tile, then only the columns right of it
            first_col = i + 1
        for j in range(first_col, N_batches_X2):
            yield (False, i, j)


def tile_count(N, N2, batch_size):
    nbx = _round_up_div(N, batch_size)
    if N2 is None:
        return max(1, nbx * (nbx + 1) // 2)
    return nbx * _round_up_div(N2, batch_size)


def worker_tiles(N, N2, batch_size, worker_rank=0, n_workers=1):
    """The (same, block_i, block_j) list ``ProductIterator`` serves to this worker."""
    nbx = _round_up_div(N, batch_size)
    same = N2 is None
    nb2 = nbx if same else _round_up_div(N2, batch_size)
    start, count = _this_worker_batch(tile_count(N, N2, batch_size), worker_rank, n_workers)
    return list(itertools.islice(_product_generator(nbx, nb2, same), start, start + count))


def tile_pairs(same, i, j, N, N2, batch_size):
    """Unique pair entries of tile (i, j): a diagonal tile of a symmetric Gram only computes j >= i."""
    n = min(batch_size, N - i * batch_size)
    if same:
        return n * (n + 1) // 2
    return n * min(batch_size, N2 - j * batch_size)


def worker_tiles_balanced(N, N2, batch_size, worker_rank=0, n_workers=1):
    """Like ``worker_tiles`` -- a contiguous slice of the reference's tile list -- but the cut
    points equalise the number of PAIR ENTRIES per worker instead of the number of tiles
    (the reference's count split, data.py:11-19, gives the ranks that own many diagonal and edge
    tiles up to ~8 % less work).  A pure function of (N, N2, batch_size, n_workers), so block
    ownership is reproducible and the NaN-fill / merge contract holds unchanged."""
    nbx = _round_up_div(N, batch_size)
    same = N2 is None
    nb2 = nbx if same else _round_up_div(N2, batch_size)
    tiles = list(_product_generator(nbx, nb2, same))
    cost, acc = [], 0
    for s, i, j in tiles:
        acc += tile_pairs(s, i, j, N, N if same else N2, batch_size)
        cost.append(acc)
    total = acc

    def cut(r):  # first tile of worker r: the first tile whose preceding cost reaches r/n of the total
        if r <= 0:
            return 0
        if r >= n_workers:
            return len(tiles)
        target = total * r / n_workers
        lo, hi = 0, len(tiles)
        while lo < hi:
            mid = (lo + hi) // 2
            if cost[mid] < target:
                lo = mid + 1
            else:
                hi = mid
        # tile `lo` straddles the target: give it to the side that leaves the smaller error
        before = cost[lo - 1] if lo > 0 else 0
        return lo + 1 if (cost[lo] - target) < (target - before) else lo
    return tiles[cut(worker_rank):cut(worker_rank + 1)]


class ResidentDataset(Dataset):
    """Images and labels held as two tensors (on any device).  Slicing a batch is a view, so a
    dataset kept in HBM feeds tiles with no host work at all."""

    def __init__(self, images, labels=None):
        self.images = images
        self.labels = labels if labels is not None else torch.zeros(len(images), dtype=torch.long)

    def __len__(self):
        return self.images.shape[0]

    def __getitem__(self, i):
        return self.images[i], self.labels[i]

    def to(self, device):
        return ResidentDataset(self.images.to(device), self.labels)


def _batch(dataset, lo, hi):
    hi = min(hi, len(dataset))
    if hasattr(dataset, "images") and hasattr(dataset, "labels"):
        return [dataset.images[lo:hi], dataset.labels[lo:hi]]
    return default_collate([dataset[k] for k in range(lo, hi)])


class ProductIterator(object):
    """Tiles of the product X x X2 (or the upper triangle of X x X) for one worker."""

    def __init__(self, batch_size, X, X2=None, worker_rank=0, n_workers=1):
        self.same = X2 is None
        self.X, self.X2 = X, (X if X2 is None else X2)
        self.batch_size = batch_size
        self.worker_rank = worker_rank
        self._tiles = worker_tiles(len(X), None if self.same else len(X2), batch_size, worker_rank, n_workers)
        self.batches_this_worker = len(self._tiles)
        self._pos = 0
        self._row, self.x_batch = None, None

    def __len__(self):
        return self.batches_this_worker

    def __iter__(self):
        return self

    def __next__(self):
        if self._pos >= len(self._tiles):
            raise StopIteration
        same, i, j = self._tiles[self._pos]
        self._pos += 1
        bs = self.batch_size
        if i != self._row:  # the X batch only changes with the row
            self._row, self.x_batch = i, _batch(self.X, i * bs, (i + 1) * bs)
        x2_batch = self.x_batch if (same and self.X2 is self.X) else _batch(self.X2, j * bs, (j + 1) * bs)
        return (same, (i * bs, self.x_batch), (j * bs, x2_batch))


class DiagIterator(object):
    """Index-aligned batches of X (and X2) for diagonal kernels; never split across workers."""

    def __init__(self, batch_size, X, X2=None):
        self.batch_size = batch_size
        self.same = X2 is None
        self.X, self.X2 = X, X2
        n = len(X) if self.same else min(len(X), len(X2))
        # zip(DataLoader(X), DataLoader(X2)) stops with the shorter loader (reference data.py:107-110)
        self.length = _round_up_div(len(X), batch_size) if self.same else min(
            _round_up_div(len(X), batch_size), _round_up_div(len(X2), batch_size))
        self._n, self._k = n, 0

    def __iter__(self):
        return self

    def __len__(self):
        return self.length

    def __next__(self):
        if self._k >= self.length:
            raise StopIteration
        ib = self._k * self.batch_size
        self._k += 1
        xy = _batch(self.X, ib, ib + self.batch_size)
        xy2 = xy if self.same else _batch(self.X2, ib, ib + self.batch_size)
        return (self.same, (ib, xy), (ib, xy2))


def _as_tensors(ds):
    """(images [N,C,H,W] float32 in [0,1], labels [N]) of a whole torchvision-style dataset without
    going through per-item PIL decoding, or None when the dataset offers no array view.  Equals
    what ``ToTensor`` yields item by item: uint8 HW / HWC -> float CHW / 255."""
    if hasattr(ds, "images") and hasattr(ds, "labels") and torch.is_tensor(ds.images):
        return ds.images, torch.as_tensor(ds.labels)
    data, targets = getattr(ds, "data", None), getattr(ds, "targets", None)
    if data is None or targets is None:
        return None
    data = torch.as_tensor(data)
    if data.dtype != torch.uint8 or data.dim() not in (3, 4):
        return None
    if data.dim() == 3:
        data = data[:, None]            # MNIST: [N, H, W]
    else:
        data = data.permute(0, 3, 1, 2)  # CIFAR: [N, H, W, C]
    return data.to(torch.float32).div(255), torch.as_tensor(targets, dtype=torch.long)


class DatasetFromConfig(object):
    """train / validation / test subsets described by a config module
    (``dataset``, ``dataset_name``, ``transforms``, ``train_range``, ``validation_range``,
    ``test_range``), as the reference's data.py:129-162.

    When the config adds no transforms and the dataset exposes its pixels as one array
    (torchvision MNIST / CIFAR ``.data``, ``SyntheticImages.images``), the three subsets are
    ``ResidentDataset`` tensors -- bit-identical to per-item ``ToTensor`` output -- that the tile
    iterators slice without any per-item work; ``resident()`` moves one to the GPU."""

    def __init__(self, datasets_path, config, download=True):
        self.config = config
        root = os.path.join(datasets_path, config.dataset_name)
        trans = None
        try:
            import torchvision
            trans = torchvision.transforms.ToTensor()
            if len(config.transforms) > 0:
                trans = torchvision.transforms.Compose([trans] + list(config.transforms))
        except ImportError:  # synthetic datasets do not need torchvision
            if len(config.transforms) > 0:
                raise
        train_full = config.dataset(root, train=True, download=download, transform=trans)
        test_full = config.dataset(root, train=False, transform=trans)
        self.data_full = ConcatDataset([train_full, test_full])
        arrays = None
        if len(config.transforms) == 0:
            a, b = _as_tensors(train_full), _as_tensors(test_full)
            if a is not None and b is not None:
                arrays = torch.cat([a[0], b[0]]), torch.cat([a[1], b[1]])
        if arrays is not None:
            def pick(rng):
                idx = torch.as_tensor(list(rng), dtype=torch.long)
                return ResidentDataset(arrays[0][idx].contiguous(), arrays[1][idx].contiguous())
        else:
            def pick(rng):
                return Subset(self.data_full, rng)
        self.train = pick(config.train_range)
        self.validation = pick(config.validation_range)
        self.test = pick(config.test_range)

    @staticmethod
    def load_full(dataset):
        return _batch(dataset, 0, len(dataset))

    @staticmethod
    def resident(dataset, device="cuda"):
        """The whole subset as a ``ResidentDataset`` on ``device`` (one upload; MNIST train is
        188 MB, CIFAR-10 614 MB -- trivial next to 180 GB of HBM)."""
        x, y = _batch(dataset, 0, len(dataset))
        return ResidentDataset(x.to(device), y)


def _hhmmss(s):
    m, s = divmod(int(s), 60)
    h, m = divmod(m, 60)
    return f"{m:02d}:{s:02d}" if h == 0 else f"{h:02d}:{m:02d}:{s:02d}"


def print_timings(iterator, desc="time", print_interval=2.):
    """Yield from ``iterator`` printing ``it/s`` and an ETA every ``print_interval`` seconds, one
    line per report so that several workers can share a terminal."""
    t0 = time.perf_counter()
    total = len(iterator)
    next_report = 0.0
    for done, value in enumerate(iterator, start=1):
        yield value
        elapsed = time.perf_counter() - t0
        if elapsed >= next_report:
            rate = done / max(elapsed, 1e-9)
            print(f"{desc}: {done}/{total} it, {rate:.02f} it/s,"
                  f"[{_hhmmss(elapsed)}<{_hhmmss(total / rate)}]")
            next_report = elapsed + print_interval
