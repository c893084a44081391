"""Float64 Cholesky solve with the matrix spread over several GPUs (one process per GPU,
``torch.distributed`` over NCCL / NVLink) -- SURVEY.md 8(f) rank 1: after the Gram kernels the
single-GPU solve of exp_mnist_resnet/classify_gp.py:17-27 is the longest step of the pipeline.

Layout: block rows of 256 rows, dealt out cyclically (block b lives on rank b % world), each
stored full width, so a rank holds n / world rows (3.6 GB of the 28.8 GB at n = 60 000, world 8).
Right-looking factorisation, per block row b:

  owner      cnngp_potrf_panel_f64 on its rows      (the same panel code as the one-GPU potrf)
  everybody  receives the factored panel X = U[b, b+1:] by NCCL broadcast (<= 123 MB)
  everybody  cnngp_syrk_upper_f64: rank-256 update of the block rows it owns (DMMA)

with one block row of look-ahead: the owner of block b + 1 updates that block first, factorises
it and starts its broadcast on a side stream while all ranks are still busy with update b, so the
NVLink transfer and the latency-bound panel hide under the tensor-pipe work.

The two triangular sweeps stay distributed as well -- no rank ever holds all of U:

  forward  U^T y = b   every rank keeps an accumulator a_r [n, nrhs] (rank src starts from b, the others
                       from zero).  Block b: all-reduce of the ranks' a_r[b] (20 KB) gives b minus all
                       earlier updates; the owner solves with its diagonal block and applies
                       a_owner[later rows] -= U[b, later columns]^T y_b from its own block row.
  backward U x = y     block b (descending): the owner solves x_b with its diagonal block and
                       broadcasts it (20 KB); every rank applies y[own rows above] -= U[., b] x_b.

Both are chains of n / 256 steps of one small collective and two small launches: milliseconds, against
0.26 s for gathering a 60 000^2 factor to one GPU and sweeping there.

The compute calls go through a small backend object; the CUDA backend is the C ABI
(include/cnngp.h).  Tests drive the same orchestration over gloo with a numpy backend.
"""
import ctypes

import torch
import torch.distributed as dist

BLK = 256   # rows per block row (= the panel height of cnngp_potrf_panel_f64)
TILE = 128  # row-tile height of cnngp_syrk_upper_f64


def n_blocks(n):
    return -(-n // BLK)


def owned_blocks(rank, world, n):
    return list(range(rank, n_blocks(n), world))


def block_rows(b, n):
    return min(BLK, n - b * BLK)


class CudaBackend:
    """The sm_100a kernels behind include/cnngp.h."""

    def __init__(self):
        from . import _native as nat
        self.nat = nat
        self._work = {}  # device -> 128 x 128 doubles of panel scratch

    @staticmethod
    def _stream():
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def panel(self, rows, col0, n, info):
        """Factorise the block row held in ``rows`` [nb, n] whose diagonal starts at column col0."""
        work = self._work.get(rows.device)
        if work is None:
            work = self._work[rows.device] = torch.empty(128 * 128, dtype=torch.float64, device=rows.device)
        self.nat.check(self.nat.lib().cnngp_potrf_panel_f64(
            rows.data_ptr() + 8 * col0, rows.stride(0), n - col0, col0, info.data_ptr(), work.data_ptr(),
            self._stream()), "cnngp_potrf_panel_f64")

    def syrk(self, X, K, m, rows, col0, t_row0):
        """rows[r, col0 + j] -= sum_k X[k, t_row0 + r] X[k, j] for j >= t_row0 + r: the update of one
        owned block row that starts at trailing-relative row t_row0 (a multiple of 128)."""
        nr = rows.shape[0]
        base = rows.data_ptr() + 8 * col0 - 8 * t_row0 * rows.stride(0)
        self.nat.check(self.nat.lib().cnngp_syrk_upper_f64(
            X.data_ptr(), m, K, ctypes.c_void_p(base), rows.stride(0), m, t_row0 // TILE,
            (t_row0 + nr + TILE - 1) // TILE, self._stream()), "cnngp_syrk_upper_f64")

    def syrk_strided(self, X, K, m, local, col0, ti0, stride, q0, n_blocks):
        """The same update for the owned trailing blocks ti0, ti0 + stride, ... (stacked in ``local``
        from local block q0) in one launch."""
        self.nat.check(self.nat.lib().cnngp_syrk_upper_strided_f64(
            X.data_ptr(), m, K, ctypes.c_void_p(local.data_ptr() + 8 * col0), local.stride(0), m, ti0, stride, q0,
            n_blocks, self._stream()), "cnngp_syrk_upper_strided_f64")

    def potrs(self, U, B):
        from . import linalg
        return linalg.potrs_upper_(U, B)

    def fwd_panel(self, rows, col0, n, acc):
        """rows [nb, n]: a block row whose diagonal starts at column col0; acc [n - col0, nrhs]: its
        right-hand sides (solved in place) followed by this rank's accumulator for the later rows."""
        self.nat.check(self.nat.lib().cnngp_trsm_fwd_panel_f64(
            rows.data_ptr() + 8 * col0, rows.stride(0), rows.shape[0], n - col0, acc.data_ptr(), acc.shape[1],
            acc.stride(0), self._stream()), "cnngp_trsm_fwd_panel_f64")

    def bwd_diag(self, rows, col0, yb):
        self.nat.check(self.nat.lib().cnngp_trsm_bwd_diag_f64(
            rows.data_ptr() + 8 * col0, rows.stride(0), rows.shape[0], yb.data_ptr(), yb.shape[1], yb.stride(0),
            self._stream()), "cnngp_trsm_bwd_diag_f64")

    def rows_update(self, local, nrows, col0, nb, xb, y_local):
        """y_local[:nrows] -= local[:nrows, col0:col0 + nb] @ xb"""
        if nrows <= 0:
            return
        self.nat.check(self.nat.lib().cnngp_rows_update_f64(
            local.data_ptr() + 8 * col0, local.stride(0), nrows, nb, xb.data_ptr(), xb.stride(0), y_local.data_ptr(),
            y_local.stride(0), xb.shape[1], self._stream()), "cnngp_rows_update_f64")


class DistributedCholesky:
    """Block rows of one symmetric positive definite matrix on this rank + the factorisation."""

    def __init__(self, n, device, group=None, backend=None):
        self.n, self.device, self.group = n, device, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.backend = backend if backend is not None else CudaBackend()
        self.blocks = owned_blocks(self.rank, self.world, n)
        self.offsets, off = {}, 0
        for b in self.blocks:
            self.offsets[b] = off
            off += block_rows(b, n)
        self.local = torch.empty((max(off, 1), n), dtype=torch.float64, device=device)
        self.n_local = off

    def rows_of(self, b):
        o = self.offsets[b]
        return self.local[o:o + block_rows(b, self.n)]

    # -- data movement ---------------------------------------------------------------------------
    # Collectives only (broadcast / gather): they run on the communicator's established rings,
    # whereas first-use point-to-point channels between every pair of GPUs cost seconds to set up.
    def scatter_from(self, K, src=0, chunk_rows=16 * BLK):
        """Deal the block rows of ``K`` (on rank ``src``; float32 or float64, only j >= i is used)
        out to their owners, widened to float64 on arrival: ``K`` is broadcast in chunks of
        ``chunk_rows`` rows and every rank keeps the blocks it owns."""
        n = self.n
        meta = torch.zeros(1, dtype=torch.int64, device=self.device)
        if self.rank == src:
            meta[0] = 1 if K.dtype == torch.float64 else 0
        dist.broadcast(meta, src=src, group=self.group)
        dtype = torch.float64 if int(meta[0]) else torch.float32
        buf = None
        for r0 in range(0, n, chunk_rows):
            r1 = min(n, r0 + chunk_rows)
            if self.rank == src:
                chunk = K[r0:r1]
                if not chunk.is_contiguous():
                    chunk = chunk.contiguous()
            else:
                if buf is None:
                    buf = torch.empty((chunk_rows, n), dtype=dtype, device=self.device)
                chunk = buf[:r1 - r0]
            if self.world > 1:
                dist.broadcast(chunk, src=src, group=self.group)
            for b in range(r0 // BLK, -(-r1 // BLK)):
                if b in self.offsets:
                    lo = b * BLK - r0
                    self.rows_of(b).copy_(chunk[lo:lo + block_rows(b, n)])

    def fill_rows(self, i0, i1, panel):
        """Consumer for ``tiles.exchange_rows``: keep the 256-row blocks of panel rows [i0, i1) that
        this rank owns, widened to float64 (only j >= i is ever read)."""
        for b in range(i0 // BLK, -(-i1 // BLK)):
            if b in self.offsets:
                lo, hi = max(i0, b * BLK), min(i1, b * BLK + block_rows(b, self.n))
                o = self.offsets[b] + lo - b * BLK
                self.local[o:o + hi - lo].copy_(panel[lo - i0:hi - i0])

    def wants_rows(self, i0, i1):
        return any(b in self.offsets for b in range(i0 // BLK, -(-i1 // BLK)))

    def add_to_diagonal(self, jitter):
        """Kxx += jitter * I in float64, after the widening (classify_gp.py:30-36,64): a jitter added to
        the float32 matrix would be rounded away exactly when it is needed."""
        if not jitter:
            return
        for b in self.blocks:
            nb = block_rows(b, self.n)
            self.rows_of(b)[:, b * BLK:b * BLK + nb].diagonal().add_(jitter)

    def solve(self, Y, src=0):
        """X = (U^T U)^-1 Y with the factor left where it is.  Y [n, nrhs] float64 on rank ``src`` (ignored
        elsewhere).  Returns X [n, nrhs] on EVERY rank (it is small: the predictions need it everywhere)."""
        n, rank, world, be = self.n, self.rank, self.world, self.backend
        nblk = n_blocks(n)
        meta = torch.zeros(1, dtype=torch.int64, device=self.device)
        if rank == src:
            meta[0] = Y.shape[1]
        if world > 1:
            dist.broadcast(meta, src=src, group=self.group)
        nrhs = int(meta[0])
        acc = torch.zeros((n, nrhs), dtype=torch.float64, device=self.device)
        if rank == src:
            acc.copy_(Y)
        for b in range(nblk):  # U^T y = b
            kb, nb = b * BLK, block_rows(b, n)
            if world > 1:
                dist.all_reduce(acc[kb:kb + nb], group=self.group)
            if b in self.offsets:
                be.fwd_panel(self.rows_of(b), kb, n, acc[kb:])
        y_local = torch.zeros((max(self.n_local, 1), nrhs), dtype=torch.float64, device=self.device)
        for b in self.blocks:
            y_local[self.offsets[b]:self.offsets[b] + block_rows(b, n)].copy_(acc[b * BLK:b * BLK + block_rows(b, n)])
        X = acc  # reuse: from here on it holds the solution blocks as they are broadcast
        for b in range(nblk - 1, -1, -1):  # U x = y
            kb, nb = b * BLK, block_rows(b, n)
            xb = X[kb:kb + nb]
            if b in self.offsets:
                yb = y_local[self.offsets[b]:self.offsets[b] + nb]
                be.bwd_diag(self.rows_of(b), kb, yb)
                xb.copy_(yb)
            if world > 1:
                dist.broadcast(xb, src=b % world if self.group is None else dist.get_global_rank(self.group, b % world),
                               group=self.group)
            # this rank's rows above block b: its owned blocks < b are the first rows of the stack
            nrows = sum(block_rows(q, n) for q in self.blocks if q < b)
            be.rows_update(self.local, nrows, kb, nb, xb, y_local)
        return X

    def gather_to(self, dst=0):
        """The factor's block rows back on rank ``dst`` as one [n, n] matrix (None elsewhere)."""
        n, world = self.n, self.world
        rows_max = max(sum(block_rows(b, n) for b in owned_blocks(r, world, n)) for r in range(world))
        send = self.local if self.local.shape[0] == rows_max else torch.cat(
            [self.local[:self.n_local], self.local.new_zeros((rows_max - self.n_local, n))])
        # all_gather (a ring collective) rather than gather, which NCCL runs as point-to-point
        # transfers over channels that first have to be set up
        if world > 1:
            parts = torch.empty((world, rows_max, n), dtype=torch.float64, device=self.device)
            dist.all_gather_into_tensor(parts.view(world * rows_max, n), send[:rows_max].contiguous(), group=self.group)
        else:
            parts = send[:rows_max].unsqueeze(0)
        if self.rank != dst:
            return None
        full = torch.empty((n, n), dtype=torch.float64, device=self.device)
        for r in range(world):
            o = 0
            for b in owned_blocks(r, world, n):
                nb = block_rows(b, n)
                full[b * BLK:b * BLK + nb].copy_(parts[r][o:o + nb])
                o += nb
        return full

    # -- factorisation ---------------------------------------------------------------------------
    def factorize(self, lookahead=True):
        """In place: every block row becomes the matching rows of U (A = U^T U).  Returns LAPACK's
        info (0 = success) as a Python int, identical on every rank."""
        n, rank, world, be = self.n, self.rank, self.world, self.backend
        nblk = n_blocks(n)
        info = torch.zeros(1, dtype=torch.int32, device=self.device)
        use_cuda = self.device.type == "cuda"
        side = torch.cuda.Stream(device=self.device, priority=-1) if (use_cuda and lookahead) else None
        xbufs = [torch.empty(BLK * max(n - BLK, 1), dtype=torch.float64, device=self.device) for _ in range(2)]

        def factor_and_pack(b, xb):
            """Owner of block b: factorise it and pack U[b, b+1:] contiguously into xb."""
            kb, nb = b * BLK, block_rows(b, n)
            rows = self.rows_of(b)
            be.panel(rows, kb, n, info)
            m = n - kb - nb
            if m > 0:
                xb[:nb * m].view(nb, m).copy_(rows[:, kb + nb:])

        def update(b, xb, blocks):
            kb, nb = b * BLK, block_rows(b, n)
            m = n - kb - nb
            X = xb[:nb * m].view(nb, m)
            if not blocks:
                return
            if hasattr(be, "syrk_strided"):
                # owned blocks are an arithmetic progression (stride = world) stacked 256 rows apart
                i0 = blocks[0]
                be.syrk_strided(X, nb, m, self.local, kb + nb, i0 - b - 1, world, (i0 - rank) // world, len(blocks))
            else:
                for i in blocks:
                    be.syrk(X, nb, m, self.rows_of(i), kb + nb, (i - b - 1) * BLK)

        # block 0: factor + broadcast up front
        if rank == 0 % world:
            factor_and_pack(0, xbufs[0])
        pending = None
        for b in range(nblk):
            kb, nb = b * BLK, block_rows(b, n)
            m = n - kb - nb
            if m <= 0:
                break
            xb = xbufs[b % 2]
            if pending is not None:
                pending()  # the broadcast of panel b started during update b - 1
                pending = None
            else:
                dist.broadcast(xb[:nb * m], src=b % world, group=self.group)
            mine = [i for i in self.blocks if i > b]
            nxt = b + 1
            if side is not None and nxt < nblk:
                # look-ahead: block b + 1 first; its owner factorises it and everybody starts the
                # broadcast on the side stream while the remaining updates of step b run
                if nxt in self.offsets:
                    update(b, xb, [nxt])
                    mine = [i for i in mine if i != nxt]
                ev = torch.cuda.Event()
                ev.record()
                nb2 = block_rows(nxt, n)
                m2 = n - nxt * BLK - nb2
                xb2 = xbufs[nxt % 2]
                with torch.cuda.stream(side):
                    side.wait_event(ev)
                    if nxt in self.offsets:
                        factor_and_pack(nxt, xb2)
                    work = dist.broadcast(xb2[:nb2 * m2], src=nxt % world, group=self.group, async_op=True) if m2 > 0 else None

                def finish(work=work):
                    if work is not None:
                        work.wait()  # orders the current stream after the collective
                    torch.cuda.current_stream().wait_stream(side)
                pending = finish
                update(b, xb, mine)
            else:
                update(b, xb, mine)
                if nxt < nblk and nxt in self.offsets:
                    factor_and_pack(nxt, xbufs[nxt % 2])
        if pending is not None:
            pending()
        code = info.clone()
        dist.all_reduce(code, op=dist.ReduceOp.MAX, group=self.group)
        # the first failing pivot is the smallest non-zero info; MAX is enough to know about failure,
        # the exact index comes from the minimum over the ranks that saw one
        if int(code) != 0:
            big = torch.where(info > 0, info, torch.full_like(info, 2 ** 31 - 1))
            dist.all_reduce(big, op=dist.ReduceOp.MIN, group=self.group)
            return int(big)
        return 0


@torch.no_grad()
def solve_pos_upper_distributed(K, Y, n, device, group=None, src=0, backend=None, lookahead=True, jitter=0.0, fill=None):
    """A = (K + jitter I)^-1 Y on all ranks of ``group``: factorisation AND both triangular sweeps run on
    the block rows where they lie.  ``K`` (upper triangle, float32 or float64) and ``Y`` live on rank
    ``src``; alternatively ``fill(ch)`` places the block rows itself (``tiles.exchange_rows`` with
    ``ch.fill_rows``: the Gram matrix then never exists on one GPU).  Returns A [n, nrhs] on every rank."""
    import os
    import time
    from .linalg import NotPositiveDefiniteError
    verbose = bool(os.environ.get("CNNGP_DIST_VERBOSE")) and dist.get_rank(group) == src

    def lap(what, t0):
        if verbose:
            if device.type == "cuda":
                torch.cuda.synchronize()
            print(f"  dist solve: {what} {time.perf_counter() - t0:.3f} s")
        return time.perf_counter()

    t = time.perf_counter()
    ch = DistributedCholesky(n, device, group=group, backend=backend)
    if fill is not None:
        fill(ch)
    else:
        ch.scatter_from(K, src=src)
    ch.add_to_diagonal(jitter)
    t = lap("distribute", t)
    info = ch.factorize(lookahead=lookahead)
    t = lap("potrf", t)
    if info != 0:
        raise NotPositiveDefiniteError(info)
    A = ch.solve(Y.to(torch.float64) if ch.rank == src else None, src=src)
    lap("potrs (distributed sweeps)", t)
    return A
