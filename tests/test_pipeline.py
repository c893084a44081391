"""The experiment pipeline around the Gram kernel: datasets, the multi-worker exchange, and
save_kernel -> merge -> classify_gp (reference exp_mnist_resnet/*.py, cnn_gp/data.py:129-162)."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _synthetic_config(train=120, val=40, test=60):
    os.environ.update(CNNGP_SYNTH_TRAIN=str(train), CNNGP_SYNTH_VAL=str(val), CNNGP_SYNTH_TEST=str(test))
    import configs.synthetic as cfg
    return importlib.reload(cfg)


# ------------------------------------------------------------------------------- datasets (CPU)
def test_synthetic_dataset_splits_are_disjoint_and_deterministic():
    from cnn_gp import DatasetFromConfig
    cfg = _synthetic_config()
    a = DatasetFromConfig("/nonexistent", cfg)
    b = DatasetFromConfig("/nonexistent", cfg)
    assert (len(a.train), len(a.validation), len(a.test)) == (120, 40, 60)
    xa, ya = a.load_full(a.train)
    xb, yb = b.load_full(b.train)
    assert xa.shape == (120, 1, 28, 28) and xa.dtype == torch.float32
    assert torch.equal(xa, xb) and torch.equal(ya, yb)
    assert 0.0 <= float(xa.min()) and float(xa.max()) < 1.0
    xt, _ = a.load_full(a.test)
    xv, _ = a.load_full(a.validation)
    allx = torch.cat([xa, xv, xt]).flatten(1)
    assert len(torch.unique(allx, dim=0)) == 220, "no image may appear in two splits"


class _FakeMNIST(torch.utils.data.Dataset):
    """uint8 ``.data`` / ``.targets`` like torchvision.datasets.MNIST, items through ``transform``."""

    def __init__(self, root, train=True, download=False, transform=None):
        g = torch.Generator().manual_seed(7 if train else 8)
        n = 50 if train else 20
        self.data = torch.randint(0, 256, (n, 28, 28), generator=g, dtype=torch.uint8)
        self.targets = torch.randint(0, 10, (n,), generator=g)
        self.transform = transform

    def __len__(self):
        return len(self.data)

    def __getitem__(self, i):
        from PIL import Image
        img = Image.fromarray(self.data[i].numpy(), mode="L")
        return self.transform(img), int(self.targets[i])


class _FakeCIFAR(_FakeMNIST):
    def __init__(self, root, train=True, download=False, transform=None):
        g = np.random.default_rng(3 if train else 4)
        n = 30 if train else 10
        self.data = g.integers(0, 256, (n, 32, 32, 3), dtype=np.uint8)
        self.targets = list(g.integers(0, 10, n))
        self.transform = transform

    def __getitem__(self, i):
        from PIL import Image
        return self.transform(Image.fromarray(self.data[i])), int(self.targets[i])


@pytest.mark.parametrize("cls,shape", [(_FakeMNIST, (1, 28, 28)), (_FakeCIFAR, (3, 32, 32))])
def test_resident_fast_path_equals_per_item_totensor(cls, shape):
    """The array view of a torchvision dataset must give the bytes ToTensor gives item by item
    (reference data.py:143-151), in the config's index order."""
    pytest.importorskip("torchvision")
    pytest.importorskip("PIL")
    from types import SimpleNamespace
    from cnn_gp import DatasetFromConfig
    n_tr = 50 if cls is _FakeMNIST else 30
    cfg = SimpleNamespace(dataset=cls, dataset_name="FAKE", transforms=[], train_range=range(5, n_tr - 5),
                          validation_range=list(range(n_tr - 5, n_tr)) + list(range(0, 5)),
                          test_range=range(n_tr, n_tr + 10))
    ds = DatasetFromConfig("/nonexistent", cfg)
    from cnn_gp.data import ResidentDataset
    assert isinstance(ds.train, ResidentDataset)
    for name in ("train", "validation", "test"):
        sub = getattr(ds, name)
        rng = getattr(cfg, name + "_range")
        assert sub.images.shape == (len(rng),) + shape
        for k in (0, len(rng) // 2, len(rng) - 1):
            x, y = ds.data_full[list(rng)[k]]
            assert torch.equal(sub.images[k], x) and int(sub.labels[k]) == y


# --------------------------------------------------------------- multi-worker exchange (gloo, CPU)
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gather_worker(rank, world, port, N, N2, bs, same, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path[:0] = [os.path.join(ROOT, "cnn-gp_b200"), ROOT]
    import torch.distributed as dist
    from cnn_gp.data import worker_tiles
    from cnn_gp.tiles import gather_blocks, row_segments
    dist.init_process_group("gloo", rank=rank, world_size=world)
    truth = torch.arange(N * N2, dtype=torch.float32).reshape(N, N2) + 1.0
    K = torch.full((N, N2), float("nan"))
    pairs = 0
    for r, has_diag, c0, c1 in row_segments(worker_tiles(N, None if same else N2, bs, rank, world)):
        i0, i1 = r * bs, min(N, (r + 1) * bs)
        if has_diag:
            K[i0:i1, i0:i1] = truth[i0:i1, i0:i1]
            pairs += (i1 - i0) ** 2
        if c0 is not None:
            j0, j1 = c0 * bs, min(N2, c1 * bs)
            K[i0:i1, j0:j1] = truth[i0:i1, j0:j1]
            pairs += (i1 - i0) * (j1 - j0)
    counts = torch.tensor([pairs], dtype=torch.int64)
    dist.all_reduce(counts)
    merged = gather_blocks(K, dst=0)
    if rank == 0:
        ret["merged"] = merged.numpy()
        ret["pairs"] = int(counts[0])
    else:
        assert merged is None
    dist.destroy_process_group()


@pytest.mark.parametrize("N,N2,bs,same", [(50, 50, 8, True), (37, 53, 10, False)])
def test_two_workers_cover_the_matrix_and_gather_like_merge(N, N2, bs, same):
    """world_size 2 over gloo: the reference's contiguous tile split (data.py:11-19) gives every
    tile to exactly one worker, and gather_blocks reproduces merge_h5_files' NaN-fill semantics:
    for Kxx the strictly lower block triangle stays NaN."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, N, N2, bs, same, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    merged = ret["merged"]
    truth = np.arange(N * N2, dtype=np.float32).reshape(N, N2) + 1.0
    bi, bj = np.arange(N)[:, None] // bs, np.arange(N2)[None, :] // bs
    owned = (bj >= bi) if same else np.ones((N, N2), bool)
    assert ret["pairs"] == owned.sum(), "tiles must be disjoint and complete"
    np.testing.assert_array_equal(merged[owned], truth[owned])
    assert np.isnan(merged[~owned]).all()


@pytest.mark.parametrize("suffix", ["", ".h5"])  # .npy directory store / native HDF5 file
def test_merge_script_fills_only_nan(tmp_path, suffix):
    from cnn_gp.block_store import open_store
    from exp_mnist_resnet.merge_h5_files import merge
    a, b = str(tmp_path / ("a" + suffix)), str(tmp_path / ("b" + suffix))
    with open_store(a, "w") as f:
        d = f.create_dataset("Kxx", shape=(1, 4, 4), dtype=np.float32, fillvalue=np.nan)
        d[0, :2, :] = 1.0
        f.create_dataset("only_a", shape=(1, 2), dtype=np.float32, fillvalue=np.nan)
    with open_store(b, "w") as f:
        d = f.create_dataset("Kxx", shape=(1, 4, 4), dtype=np.float32, fillvalue=np.nan)
        d[0, 1:3, :] = 2.0
    merge(a, [b])
    with open_store(a, "r") as f:
        got = f["Kxx"][0]
        assert np.isnan(f["only_a"][0]).all()
    assert (got[:2] == 1.0).all() and (got[2] == 2.0).all() and np.isnan(got[3]).all()


# ------------------------------------------------------------------------ whole pipeline (GPU)
@pytest.mark.gpu
def test_save_kernel_resident_equals_reference_loop_and_classify_matches_scipy(tmp_path):
    """Two workers, both save paths, merge, classify.  The resident path must write the very
    bytes of the literal save_K loop; decisions must equal the oracle's scipy solve on the same
    Kxx bytes (reference classify_gp.py:17-42)."""
    from cnn_gp import DatasetFromConfig
    from cnn_gp.block_store import open_store
    from exp_mnist_resnet import classify_gp, merge_h5_files, save_kernel
    from oracle import oracle
    cfg = _synthetic_config(300, 80, 100)
    ds = DatasetFromConfig("/nonexistent", cfg)
    paths = {}
    for resident in (True, False):
        for rank in range(2):
            p = str(tmp_path / f"{'res' if resident else 'ref'}_{rank}.h5")  # real HDF5 files (native store)
            save_kernel.compute_all(cfg, ds, p, batch_size=64, n_workers=2, worker_rank=rank, resident=resident)
            paths[resident, rank] = p
    for rank in range(2):
        with open_store(paths[True, rank], "r") as a, open_store(paths[False, rank], "r") as b:
            assert sorted(a.keys()) == sorted(b.keys())
            assert ("Kv_diag" in a.keys()) == (rank == 0)
            for k in a.keys():
                x, y = a[k][...], b[k][...]
                assert x.shape == y.shape and x.dtype == np.float32
                np.testing.assert_array_equal(np.isnan(x), np.isnan(y))
                np.testing.assert_array_equal(x[~np.isnan(x)], y[~np.isnan(y)])
    merge_h5_files.merge(paths[True, 0], [paths[True, 1]])
    with open_store(paths[True, 0], "r") as f:
        Kxx, Kxtx, Kxvx = f["Kxx"][0], f["Kxtx"][0], f["Kxvx"][0]
        kt = f["Kt_diag"][0]
    bi = np.arange(300) // 64
    assert np.isnan(Kxx[bi[:, None] > bi[None, :]]).all() and np.isfinite(Kxx[bi[:, None] <= bi[None, :]]).all()
    assert np.isfinite(Kxtx).all() and np.isfinite(Kxvx).all() and np.isfinite(kt).all()
    # kernel values against the oracle on a corner
    xtr, ytr = ds.load_full(ds.train)
    want = oracle.gram(cfg.initial_model, xtr[:24].numpy(), xtr[:24].numpy(), same=True)
    np.testing.assert_allclose(Kxx[:24, :24], want, rtol=1e-5)
    res = classify_gp.classify(cfg, ds, paths[True, 0])
    Y = -np.ones((300, 10))
    Y[np.arange(300), ytr.numpy()] = 1.0
    A = oracle.solve_system(np.triu(Kxx.astype(np.float64)), Y)
    np.testing.assert_allclose(res["A"].cpu().numpy(), A, rtol=0, atol=1e-8 * np.abs(A).max())
    for key, K, sub in (("validation", Kxvx, ds.validation), ("test", Kxtx, ds.test)):
        pred = oracle.predict(K.astype(np.float64), A)
        _, ys = ds.load_full(sub)
        assert res[key] == pytest.approx(float((pred == ys.numpy()).mean()))
        got = classify_gp.predict(res["A"], torch.from_numpy(K))
        np.testing.assert_array_equal(got.numpy(), pred)
    assert res["test"] > 0.9, "class-template images are separable"


@pytest.mark.gpu
@pytest.mark.parametrize("batch_size,double", [(63, False), (64, True), (45, True)])
def test_resident_rows_with_odd_batch_size_and_float64_images(tmp_path, batch_size, double):
    """The reference's save_K accepts any batch size and any image dtype (the numpy tile is cast into
    the float32 dataset, kernel_save_tools.py:21,55-58).  The resident path must too: blocks with an
    odd origin go to the generic kernel (the fused kernels' variance maps interleave image pairs), a
    float64 dataset computes in float64 and is narrowed on the way out -- same bytes as the literal loop."""
    from cnn_gp import save_K
    from cnn_gp.block_store import open_store
    from cnn_gp.kernel_save_tools import save_K_resident
    cfg = _synthetic_config(150, 40, 50)
    model = cfg.initial_model.cuda()
    g = torch.Generator().manual_seed(5)
    X = torch.rand(150, 1, 28, 28, generator=g)
    X2 = torch.rand(70, 1, 28, 28, generator=g)
    if double:
        X, X2, model = X.double(), X2.double(), model.double()
    ds, ds2 = torch.utils.data.TensorDataset(X, torch.zeros(150)), torch.utils.data.TensorDataset(X2, torch.zeros(70))

    def kern(x, x2, same, diag):
        with torch.no_grad():
            return model(x.cuda(), x2.cuda(), same, diag).cpu().numpy()
    try:
        with open_store(str(tmp_path / "a.h5"), "w") as fa, open_store(str(tmp_path / "b.h5"), "w") as fb:
            for name, A, B in (("Kxx", ds, None), ("Kx2x", ds2, ds)):
                save_K_resident(fa, model, name, A, B, False, batch_size, 1, 2)
                save_K(fb, kern, name, A, B, False, batch_size, 1, 2)
                a, b = fa[name][...], fb[name][...]
                assert a.dtype == np.float32 and a.shape == b.shape
                np.testing.assert_array_equal(np.isnan(a), np.isnan(b))
                assert np.isfinite(a).any()
                np.testing.assert_allclose(a[~np.isnan(a)], b[~np.isnan(b)], rtol=1e-5 if not double else 0, atol=0)
    finally:
        model.float().cpu()


@pytest.mark.gpu
def test_classify_api_matches_reference_functions():
    """solve_system / diag_add / print_accuracy keep the reference's signatures and semantics."""
    from exp_mnist_resnet import classify_gp
    g = np.load(os.path.join(ROOT, "tests", "golden", "solve.npz"))
    K = torch.from_numpy(np.triu(g["Kxx"].astype(np.float64)))
    classify_gp.diag_add(K, 0.0)
    with pytest.raises(AssertionError):
        classify_gp.solve_system(K.float(), torch.from_numpy(g["Y"]))
    A = classify_gp.solve_system(K, torch.from_numpy(g["Y"]))
    assert A.device.type == "cpu" and A.dtype == torch.float64
    np.testing.assert_allclose(A.numpy(), g["A"], rtol=0, atol=1e-9 * np.abs(g["A"]).max())
    acc = classify_gp.print_accuracy(A, torch.from_numpy(g["Kxtx"]).to(torch.float64), g["yte"], "test")
    assert acc == pytest.approx(float((g["pred"] == g["yte"]).mean()))
    K2 = np.eye(3)
    classify_gp.diag_add(K2, 0.5)
    assert (np.diag(K2) == 1.5).all()


@pytest.mark.parametrize("N,N2,bs,n", [(1000, None, 64, 8), (333, None, 50, 3), (120, 77, 16, 5), (64, None, 64, 4)])
def test_balanced_split_is_a_contiguous_partition_with_even_pair_counts(N, N2, bs, n):
    """worker_tiles_balanced: same tile order as the reference, every tile owned exactly once,
    pair counts per worker within one tile of the mean (the reference's count split is not)."""
    from cnn_gp.data import worker_tiles, worker_tiles_balanced, tile_pairs
    full = worker_tiles(N, N2, bs, 0, 1)
    parts = [worker_tiles_balanced(N, N2, bs, r, n) for r in range(n)]
    assert sum(parts, []) == full
    same = N2 is None
    cost = [sum(tile_pairs(s, i, j, N, N if same else N2, bs) for s, i, j in p) for p in parts]
    total = N * (N + 1) // 2 if same and N <= bs else sum(cost)
    assert sum(cost) == total
    biggest = max(tile_pairs(s, i, j, N, N if same else N2, bs) for s, i, j in full)
    assert max(cost) - sum(cost) / n <= biggest


@pytest.mark.parametrize("N,bs,world", [(28284, 500, 8), (14142, 500, 2), (1000, 200, 1), (1100, 200, 3), (333, 50, 4), (64, 64, 2)])
@pytest.mark.parametrize("max_rows", [None, 1, 3])
def test_launch_groups_cover_a_workers_tiles_once(N, bs, world, max_rows):
    """tiles.launch_groups / _streaming_order (host logic of the band launches, cnngp_gram_band): every tile of a
    worker's slice of the reference tile list (data.py:11-29) lands in exactly one launch group; bands hold whole
    block rows only and at most ``max_rows`` of them; the streaming order keeps the groups, puts the bottom rows
    first and ends with one whole block row -- the topmost one of the slice."""
    from cnn_gp.data import worker_tiles_balanced
    from cnn_gp.tiles import launch_groups, row_segments, _streaming_order, _band_pairs
    nbx = -(-N // bs)
    total = 0
    for rank in range(world):
        tiles = worker_tiles_balanced(N, None, bs, rank, world)
        groups = launch_groups(row_segments(tiles), nbx, max_rows)
        covered = []
        for g in groups:
            if g[0] == "band":
                assert max_rows is None or g[2] - g[1] + 1 <= max_rows
                for r in range(g[1], g[2] + 1):
                    covered += [(True, r, r)] + [(False, r, c) for c in range(r + 1, nbx)]
                total += _band_pairs(N, bs, g[1] * bs, min(N, (g[2] + 1) * bs))
            else:
                r, has_diag, c0, c1 = g[1]
                i0, i1 = r * bs, min(N, (r + 1) * bs)
                if has_diag:
                    covered.append((True, r, r))
                    total += (i1 - i0) * (i1 - i0 + 1) // 2
                if c0 is not None:
                    covered += [(False, r, c) for c in range(c0, c1)]
                    total += (i1 - i0) * (min(N, c1 * bs) - c0 * bs)
        assert sorted(covered) == sorted(tiles)
        order = _streaming_order(groups)
        rows = lambda gs: sorted(r for g in gs for r in (range(g[1], g[2] + 1) if g[0] == "band" else [g[1][0]]))
        assert rows(order) == rows(groups)
        whole = [g for g in groups if g[0] == "band"]
        if whole:
            top = min(g[1] for g in whole)
            assert order[-1] == ("band", top, top)
    assert total == N * (N + 1) // 2


# ----------------------------------------------------- distributed Cholesky orchestration (gloo, CPU)
class _NumpyBackend:
    """The two compute calls of cnn_gp.linalg_dist restated with numpy / scipy (test-only), so that
    ownership, packing, broadcasts and the shifted-base indexing can be checked without a GPU."""

    def panel(self, rows, col0, n, info):
        import scipy.linalg
        a = rows.numpy()
        nb = a.shape[0]
        D = np.triu(a[:, col0:col0 + nb])
        D = D + np.triu(D, 1).T
        try:
            U = np.linalg.cholesky(D).T
        except np.linalg.LinAlgError:
            info[0] = col0 + 1
            return
        a[:, col0:col0 + nb] = np.triu(U) + np.tril(a[:, col0:col0 + nb], -1)
        if col0 + nb < n:
            a[:, col0 + nb:] = scipy.linalg.solve_triangular(U, a[:, col0 + nb:], trans="T", lower=False)

    def syrk(self, X, K, m, rows, col0, t_row0):
        x, a = X.numpy(), rows.numpy()
        for r in range(a.shape[0]):
            i = t_row0 + r
            a[r, col0 + i:col0 + m] -= x[:K, i] @ x[:K, i:m]

    def syrk_strided(self, X, K, m, local, col0, ti0, stride, q0, n_blocks):
        for k in range(n_blocks):
            rows = local[256 * (q0 + k):256 * (q0 + k + 1)]
            t_row0 = 256 * (ti0 + stride * k)
            self.syrk(X, K, m, rows[:max(0, min(256, m - t_row0))], col0, t_row0)

    def potrs(self, U, B):
        import scipy.linalg
        u = np.triu(U.numpy())
        y = scipy.linalg.solve_triangular(u, B.numpy(), trans="T", lower=False)
        return torch.from_numpy(scipy.linalg.solve_triangular(u, y, lower=False))

    # the distributed sweeps (cnngp_trsm_fwd_panel_f64 / cnngp_trsm_bwd_diag_f64 / cnngp_rows_update_f64)
    def fwd_panel(self, rows, col0, n, acc):
        import scipy.linalg
        a, b = rows.numpy(), acc.numpy()
        nb = a.shape[0]
        b[:nb] = scipy.linalg.solve_triangular(np.triu(a[:, col0:col0 + nb]), b[:nb], trans="T", lower=False)
        b[nb:] -= a[:, col0 + nb:n].T @ b[:nb]

    def bwd_diag(self, rows, col0, yb):
        import scipy.linalg
        a, y = rows.numpy(), yb.numpy()
        y[:] = scipy.linalg.solve_triangular(np.triu(a[:, col0:col0 + a.shape[0]]), y, lower=False)

    def rows_update(self, local, nrows, col0, nb, xb, y_local):
        if nrows > 0:
            y_local.numpy()[:nrows] -= local.numpy()[:nrows, col0:col0 + nb] @ xb.numpy()


def _dist_solve_worker(rank, world, port, n, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path[:0] = [os.path.join(ROOT, "cnn-gp_b200"), ROOT, os.path.join(ROOT, "tests")]
    import torch.distributed as dist
    from cnn_gp import linalg_dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    K = Y = None
    if rank == 0:
        rng = np.random.default_rng(n)
        B = rng.standard_normal((n, n + 20))
        Kfull = B @ B.T / n + 0.1 * np.eye(n)
        Y = torch.from_numpy(rng.standard_normal((n, 3)))
        K = torch.from_numpy(np.triu(Kfull) + np.tril(np.full((n, n), np.nan), -1))  # lower triangle never read
        ret["K"], ret["Y"] = Kfull, Y.numpy()
    A = linalg_dist.solve_pos_upper_distributed(K, Y, n, torch.device("cpu"), backend=_NumpyBackend(), lookahead=False,
                                                jitter=0.25)
    ret[f"A{rank}"] = A.numpy()  # the solution comes back on every rank; no rank ever held all of U
    dist.destroy_process_group()


@pytest.mark.parametrize("n,world", [(700, 2), (513, 3), (200, 2)])
def test_distributed_cholesky_orchestration_over_gloo(n, world):
    """Block-row-cyclic ownership, panel broadcast and trailing updates of cnn_gp.linalg_dist with
    a numpy compute backend: the solution must equal scipy's on the same matrix."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_dist_solve_worker, args=(r, world, port, n, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    want = np.linalg.solve(ret["K"] + 0.25 * np.eye(n), ret["Y"])  # the jitter is added in float64 on the owners
    for r in range(world):
        np.testing.assert_allclose(ret[f"A{r}"], want, rtol=0, atol=1e-9 * np.abs(want).max())
        np.testing.assert_array_equal(ret[f"A{r}"], ret["A0"])


def _exchange_worker(rank, world, port, N, bs, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path[:0] = [os.path.join(ROOT, "cnn-gp_b200"), ROOT, os.path.join(ROOT, "tests")]
    import torch.distributed as dist
    from cnn_gp import linalg_dist
    from cnn_gp.tiles import RowShard, exchange_rows
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(3)
    B = rng.standard_normal((N, N + 10))
    truth = (B @ B.T / N + 0.5 * np.eye(N)).astype(np.float32)
    shard = RowShard(N, N, bs, rank, world, True, torch.device("cpu"))
    for r, has_diag, c0, c1 in shard.segments:  # stand-in for shard.compute(job): this worker's tiles only
        i0, i1 = r * bs, min(N, (r + 1) * bs)
        if has_diag:
            shard.data[i0 - shard.row_lo:i1 - shard.row_lo, i0:i1] = torch.from_numpy(truth[i0:i1, i0:i1])
        if c0 is not None:
            j0, j1 = c0 * bs, min(N, c1 * bs)
            shard.data[i0 - shard.row_lo:i1 - shard.row_lo, j0:j1] = torch.from_numpy(truth[i0:i1, j0:j1])
    ret[f"rows{rank}"] = (shard.row_lo, shard.row_hi)
    Y = torch.from_numpy(rng.standard_normal((N, 2))) if rank == 0 else None
    A = linalg_dist.solve_pos_upper_distributed(
        None, Y, N, torch.device("cpu"), backend=_NumpyBackend(), lookahead=False,
        fill=lambda ch: exchange_rows(shard, ch.fill_rows, wanted=ch.wants_rows))
    full = torch.full((N, N), float("nan"))
    exchange_rows(shard, lambda i0, i1, panel: full[i0:i1].copy_(panel))
    if rank == 0:
        ret["A"], ret["Y"], ret["truth"], ret["full"] = A.numpy(), Y.numpy(), truth, full.numpy()
    dist.destroy_process_group()


@pytest.mark.parametrize("N,bs,world", [(600, 64, 3), (300, 50, 2)])
def test_row_shards_feed_the_distributed_cholesky_without_a_gather(N, bs, world):
    """Every worker keeps only the block rows its tiles touch (RowShard); exchange_rows moves finished
    rows from their owners straight into the block-cyclic rows of the distributed Cholesky (the Gram
    matrix never exists on one rank), or to a rank that wants the whole matrix: same NaN pattern as
    merge_h5_files (strictly lower block triangle), same values."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_exchange_worker, args=(r, world, port, N, bs, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    truth = ret["truth"]
    want = np.linalg.solve(truth.astype(np.float64), ret["Y"])
    np.testing.assert_allclose(ret["A"], want, rtol=0, atol=1e-8 * np.abs(want).max())
    bi = np.arange(N) // bs
    owned = bi[None, :] >= bi[:, None]
    np.testing.assert_array_equal(ret["full"][owned], truth[owned])
    assert np.isnan(ret["full"][~owned]).all()
    spans = [ret[f"rows{r}"] for r in range(world)]
    assert max(hi - lo for lo, hi in spans) < N, "no worker holds every row"
