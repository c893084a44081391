"""Gram matrices of GPU-resident datasets, tile list sharded over workers.

The reference computes one 200 x 200 tile per Python iteration, copying both image batches to
the device and the result back every time (exp_mnist_resnet/save_kernel.py:21-24,
cnn_gp/kernel_save_tools.py:49-58).  Here the images stay in HBM, the per-image variance maps
are computed once per dataset, and a worker's contiguous slice of the reference's tile list
(cnn_gp/data.py:11-29) is evaluated with ONE launch for every run of whole block rows (a band of the
symmetric Gram, `cnngp_gram_band`: the diagonal tiles and everything to their right) and at most two launches --
the diagonal tile and the rectangle to its right -- for the partial block rows at either end of the slice.

Tiles are independent, so workers never talk to each other while computing.  The single exchange
step (the reference does it through files, merge_h5_files.py) is `exchange_rows`: every worker
keeps only the block rows its tiles touch (`RowShard`), and finished block rows travel from their
owner(s) to whoever consumes them -- the block-cyclic rows of the distributed Cholesky, or one
rank that writes the store -- by NCCL broadcasts of row panels over NVLink.  `gather_blocks` is the
older whole-matrix form (NaN-marked N x N matrices on every rank, reduced), kept for small
matrices and for the CPU (gloo) tests.
"""
import torch

from . import engine
from .data import worker_tiles, worker_tiles_balanced


class GramJob:
    """Variance maps + plan for one (model, X[, X2]) pair, ready to evaluate any tile."""

    def __init__(self, model, X, X2=None):
        self.model = model
        self.X = X.contiguous()
        self.same = X2 is None
        self.X2 = self.X if self.same else X2.contiguous()
        assert self.X.is_cuda and self.X2.is_cuda
        with torch.cuda.device(self.X.device):
            self.plan = engine.plan_for(model, X.shape[2], X.shape[3], X.dtype)
            self.aux_x, _, self.kdiag = engine.variances(self.plan, self.X)
            self.aux_x2 = self.aux_x if self.same else engine.variances(self.plan, self.X2)[0]
        self.launches = 1 if self.same else 2

    def block(self, out, i0, i1, j0, j1, symmetric):
        """out[i0:i1, j0:j1] = K(X[i0:i1], X2[j0:j1]) (upper triangle mirrored when symmetric)."""
        self.block_into(out[i0:i1, j0:j1], i0, i1, j0, j1, symmetric)

    def block_into(self, view, i0, i1, j0, j1, symmetric):
        """view[...] = K(X[i0:i1], X2[j0:j1]); ``view`` is any [i1-i0, j1-j0] tensor with unit
        column stride (e.g. a slice of a row buffer) and the images' dtype.
        The fused kernels' variance maps interleave images 2k and 2k+1, so a block whose origin is
        odd (an odd ``batch_size``: the reference accepts any) goes to the generic kernel, which
        reads the plain per-image rows."""
        path = "generic" if (i0 % 2 or j0 % 2) else None
        with torch.cuda.device(self.X.device):
            engine.gram_with_aux(self.plan, self.X[i0:i1], self.X2[j0:j1], self.aux_x[i0:i1],
                                 self.aux_x2[j0:j1], same=symmetric, diag=False, symmetric=symmetric,
                                 out=view, kdiag=self.kdiag[i0:i1] if symmetric else None, path=path)
        self.launches += engine.last_launches()

    def can_band(self, i0):
        """Whole block rows starting at image i0 can go out as ONE launch (cnngp_gram_band): a symmetric job on a
        program a fused kernel covers, float32, even origin, kernel family not pinned to the generic one."""
        return (self.same and self.plan.fused_kind in (2, 3) and self.X.dtype == torch.float32 and i0 % 2 == 0
                and engine._force_path != "generic")

    def band_into(self, view, i0, i1, block):
        """view[...] = rows [i0, i1) x columns [i0, N) of K(X, X): entries j >= i, mirrored inside the diagonal
        blocks of ``block`` rows (the reference's same=True tiles); the block triangle below them is not touched."""
        with torch.cuda.device(self.X.device):
            engine.gram_band(self.plan, self.X[i0:], i1 - i0, self.aux_x[i0:], self.kdiag[i0:], block, view)
        self.launches += engine.last_launches()


def launch_groups(segments, n_block_rows, max_rows=None):
    """Group the block-row segments of a worker into launches: runs of WHOLE block rows of a symmetric Gram (the
    diagonal tile and everything to its right) become ("band", first_row, last_row) -- one launch each, at most
    ``max_rows`` block rows --, everything else stays ("row", segment) with its one or two launches."""
    groups, run = [], []

    def flush():
        if run:
            groups.append(("band", run[0], run[-1]))
            run.clear()
    for seg in segments:
        r, has_diag, c0, c1 = seg
        whole = has_diag and ((c0 == r + 1 and c1 == n_block_rows) or (c0 is None and r == n_block_rows - 1))
        if whole and (not run or (run[-1] == r - 1 and (max_rows is None or len(run) < max_rows))):
            run.append(r)
        elif whole:
            flush()
            run.append(r)
        else:
            flush()
            groups.append(("row", seg))
    flush()
    return groups


def row_segments(tiles):
    """Group a contiguous slice of the reference tile list by block row:
    -> [(block_row, has_diag_tile, first_col, last_col_exclusive)], columns excluding the
    diagonal tile (first_col is None when the row holds only its diagonal tile)."""
    rows, order = {}, []
    for same, i, j in tiles:
        if i not in rows:
            rows[i] = [False, None, None]
            order.append(i)
        seg = rows[i]
        if same:
            seg[0] = True
        elif seg[1] is None:
            seg[1], seg[2] = j, j + 1
        else:
            assert j == seg[2], "tile slice is not contiguous within a row"
            seg[2] = j + 1
    return [(i, rows[i][0], rows[i][1], rows[i][2]) for i in order]


def compute_worker_blocks(job, out, batch_size, worker_rank=0, n_workers=1, balanced=False, on_row=None,
                          rows_per_launch=None):
    """Fill ``out`` ([N, N2], any float dtype matching the job) with this worker's tiles; entries
    owned by other workers are left untouched.  Returns the number of unique pairs computed.
    ``balanced`` cuts the reference's tile list by pair count instead of tile count.
    ``on_row(i0, i1)`` is called after the launches of each block row (or band of block rows, see
    ``rows_per_launch`` / ``_compute_segments``) have been queued (e.g. to record an event and start copying
    the rows out on another stream)."""
    N, N2 = job.X.shape[0], job.X2.shape[0]
    split = worker_tiles_balanced if balanced else worker_tiles
    tiles = split(N, None if job.same else N2, batch_size, worker_rank, n_workers)
    return _compute_segments(job, row_segments(tiles), batch_size, lambda i0, i1: out[i0:i1], on_row, rows_per_launch)


def _band_pairs(N, bs, i0, i1):
    """unique pairs of the whole block rows [i0, i1): per block row its diagonal tile's triangle + the rectangle"""
    pairs = 0
    for a in range(i0, i1, bs):
        h = min(N, a + bs) - a
        pairs += h * (h + 1) // 2 + h * (N - a - h)
    return pairs


def _streaming_order(groups):
    """Launch order when somebody streams whole block rows out while later ones are computed.  A block row costs
    the same to copy wherever it lies, but the rows at the bottom of the triangle are short and cheap to compute:
    they go FIRST, so that every band's copy hides behind the longer rows that follow; the partial row at the top
    of the slice comes next, and the topmost whole block row -- the longest launch of the slice -- goes last and
    alone: its compute time hides every copy before it, and what is left to copy when it ends is one block row."""
    groups = groups[::-1]
    bands = [k for k, g in enumerate(groups) if g[0] == "band"]
    if bands:
        k = bands[-1]
        _, a, b = groups[k]
        groups = groups[:k] + ([("band", a + 1, b)] if b > a else []) + groups[k + 1:] + [("band", a, a)]
    return groups


def _compute_segments(job, segments, bs, rows_of, on_row, rows_per_launch):
    """Evaluate block-row segments into ``rows_of(i0, i1)`` ([i1 - i0, N2] views).  Runs of whole block rows go out
    as one band launch each (``rows_per_launch`` block rows at most; default: all of them when nobody waits for
    rows via ``on_row``, one otherwise), so a worker needs a handful of launches instead of two per block row."""
    N, N2 = job.X.shape[0], job.X2.shape[0]
    if rows_per_launch is None:
        rows_per_launch = 1 if on_row is not None else 1 << 30
    nbx = -(-N // bs)
    groups = launch_groups(segments, nbx, rows_per_launch) if job.same and rows_per_launch > 1 else [("row", s) for s in segments]
    if on_row is not None and rows_per_launch > 1 and job.same:
        groups = _streaming_order(groups)
    pairs = 0
    for g in groups:
        if g[0] == "band" and job.can_band(g[1] * bs) and bs % 2 == 0:
            i0, i1 = g[1] * bs, min(N, (g[2] + 1) * bs)
            job.band_into(rows_of(i0, i1)[:, i0:], i0, i1, bs)
            pairs += _band_pairs(N, bs, i0, i1)
            if on_row is not None:
                on_row(i0, i1)
            continue
        segs = [g[1]] if g[0] == "row" else [(r, True, r + 1 if r + 1 < nbx else None, nbx if r + 1 < nbx else None)
                                             for r in range(g[1], g[2] + 1)]
        for r, has_diag, c0, c1 in segs:
            i0, i1 = r * bs, min(N, (r + 1) * bs)
            view = rows_of(i0, i1)
            if has_diag:
                job.block_into(view[:, i0:i1], i0, i1, i0, i1, symmetric=True)
                pairs += (i1 - i0) * (i1 - i0 + 1) // 2
            if c0 is not None:
                j0, j1 = c0 * bs, min(N2, c1 * bs)
                job.block_into(view[:, j0:j1], i0, i1, j0, j1, symmetric=False)
                pairs += (i1 - i0) * (j1 - j0)
            if on_row is not None:
                on_row(i0, i1)
    return pairs


class RowShard:
    """The block rows of K(X, X2) that one worker's slice of the tile list touches: rows
    [row_lo, row_hi) x all N2 columns, NaN where another worker owns the entry (or, for a symmetric
    Gram, below the block diagonal).  A worker of an 8-way split of a 60 000^2 Gram holds 1.1 .. 3.6 GB
    instead of the 14.4 GB a full NaN-marked matrix costs."""

    def __init__(self, N, N2, batch_size, worker_rank, n_workers, same, device, dtype=torch.float32, balanced=True):
        self.N, self.N2, self.bs, self.rank, self.world, self.same = N, N2, batch_size, worker_rank, n_workers, same
        split = worker_tiles_balanced if balanced else worker_tiles
        self.balanced = balanced
        self.tiles = split(N, None if same else N2, batch_size, worker_rank, n_workers)
        self.segments = row_segments(self.tiles)
        rows = [r for r, _, _, _ in self.segments]
        self.row_lo = min(rows) * batch_size if rows else 0
        self.row_hi = min(N, (max(rows) + 1) * batch_size) if rows else 0
        self.data = torch.full((self.row_hi - self.row_lo, N2), float("nan"), dtype=dtype, device=device)

    def compute(self, job, on_row=None, rows_per_launch=None):
        """Evaluate this worker's tiles into the shard; returns the number of unique pairs."""
        return _compute_segments(job, self.segments, self.bs,
                                 lambda i0, i1: self.data[i0 - self.row_lo:i1 - self.row_lo], on_row, rows_per_launch)


def row_owners(N, N2, batch_size, n_workers, same, balanced=True):
    """For every block row of the tile grid: [(worker, col_lo, col_hi)] -- who computed which columns
    (a pure function of the split, identical on every rank; at most two workers share a row)."""
    split = worker_tiles_balanced if balanced else worker_tiles
    nbx = -(-N // batch_size)
    owners = [[] for _ in range(nbx)]
    N2 = N if (same or N2 is None) else N2
    for w in range(n_workers):
        for r, has_diag, c0, c1 in row_segments(split(N, None if same else N2, batch_size, w, n_workers)):
            lo = r * batch_size if has_diag else c0 * batch_size
            hi = min(N2, c1 * batch_size) if c0 is not None else min(N2, (r + 1) * batch_size)
            owners[r].append((w, lo, hi))
    return owners


def exchange_rows(shard, consume, group=None, wanted=None):
    """Send every finished block row from its owner(s) to all ranks, one row panel at a time, and call
    ``consume(i0, i1, panel)`` with the [i1 - i0, N2] float32 panel (NaN where nobody computed: the
    lower block triangle of a symmetric Gram) on every rank for which ``wanted(i0, i1)`` holds (default:
    all).  The panel buffer is reused: consumers copy what they keep.  Broadcasts are ring collectives on
    the communicator that already exists (no point-to-point channels to set up)."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    owners = row_owners(shard.N, shard.N2, shard.bs, world, shard.same, shard.balanced) if world > 1 else None
    dev, bs, N, N2 = shard.data.device, shard.bs, shard.N, shard.N2
    panel = torch.empty((min(bs, N), N2), dtype=shard.data.dtype, device=dev)
    for r in range(-(-N // bs)):
        i0, i1 = r * bs, min(N, (r + 1) * bs)
        buf = panel[:i1 - i0]
        if world == 1:
            buf.copy_(shard.data[i0 - shard.row_lo:i1 - shard.row_lo])
        else:
            buf.fill_(float("nan"))
            for w, lo, hi in owners[r]:
                piece = torch.empty((i1 - i0, hi - lo), dtype=buf.dtype, device=dev)
                if w == rank:
                    piece.copy_(shard.data[i0 - shard.row_lo:i1 - shard.row_lo, lo:hi])
                dist.broadcast(piece, src=dist.get_global_rank(group, w) if group is not None else w, group=group)
                buf[:, lo:hi].copy_(piece)
        if wanted is None or wanted(i0, i1):
            consume(i0, i1, buf)


def gather_blocks(out, dst=0, group=None):
    """Combine the workers' partial matrices on rank ``dst``: every entry is NaN on all workers
    but its owner (the NaN-fill / merge-where-NaN contract of kernel_save_tools.py:21-23 and
    merge_h5_files.py:27-28), so a NaN-ignoring reduction reproduces the merge.  Uses a SUM of
    NaN-zeroed matrices plus a SUM of ownership masks over NCCL (NVLink / NVSwitch)."""
    import torch.distributed as dist
    owned = ~torch.isnan(out)
    vals = torch.where(owned, out, torch.zeros_like(out))
    cnt = owned.to(out.dtype)
    dist.reduce(vals, dst=dst, op=dist.ReduceOp.SUM, group=group)
    dist.reduce(cnt, dst=dst, op=dist.ReduceOp.SUM, group=group)
    if dist.get_rank(group) == dst:
        vals[cnt == 0] = float("nan")
        return vals
    return None
