"""The native HDF5 block store (cnn-gp_b200/csrc/h5store.cpp, include/cnngp_h5.h) -- no GPU.

There is no libhdf5 / h5py in the image, so parity is pinned in two steps:
  1. oracle/h5_oracle.py -- an independent pure-Python reader written from the format
     specification -- and the native reader both decode tests/golden/libhdf5_matlab73.mat, a file
     the real HDF5 library wrote (MATLAB 7.4, 512-byte user block; values known from scipy's own
     test-suite: 0, pi/4, ..., 2 pi);
  2. every file the native writer produces is decoded by the oracle reader, structure checks
     included (B-tree key order and capacity, heap free list, link-name order, chunk alignment),
     and compared with a numpy model of the reference's h5py calls
     (cnn_gp/kernel_save_tools.py:7-58, exp_mnist_resnet/merge_h5_files.py:15-30,
     exp_mnist_resnet/classify_gp.py:45-48).
"""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import cnn_gp
from cnn_gp import block_store, data, h5store
from oracle.h5_oracle import OracleFile, H5FormatError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
REAL = os.path.join(GOLD, "libhdf5_matlab73.mat")


def test_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "cnngp_h5.h")).read()
    declared = set(re.findall(r"\b(cnngp_h5_[a-z0-9_]+)\s*\(", hdr))
    L = ctypes.CDLL(h5store.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/cnngp_h5.h but not exported"
    assert declared == set(h5store.EXPORTS)


def test_oracle_reader_decodes_a_libhdf5_file():
    f = OracleFile(REAL)
    assert (f.sb_version, f.base, f.leaf_k, f.internal_k) == (0, 512, 4, 16)
    assert list(f.datasets) == ["testdouble"]
    d = f.datasets["testdouble"]
    assert d.shape == (9, 1) and d.dtype == np.dtype("<f8") and d.layout[0] == "contiguous"
    np.testing.assert_allclose(d.read().ravel(), np.arange(9) * np.pi / 4, rtol=0, atol=1e-15)


def test_native_reader_decodes_a_libhdf5_file():
    with h5store.File(REAL, "r") as f:
        assert f.keys() == ["testdouble"] and "testdouble" in f and "nope" not in f
        d = f["testdouble"]
        assert d.shape == (9, 1) and d.dtype == np.float64 and d.chunks is None and d.maxshape == (9, 1)
        np.testing.assert_allclose(d[...].ravel(), np.arange(9) * np.pi / 4, rtol=0, atol=1e-15)
        np.testing.assert_allclose(d[2:5, 0], np.arange(2, 5) * np.pi / 4, rtol=0, atol=1e-15)
        with pytest.raises(OSError):
            d[0, 0] = 1.0  # read-only
    with pytest.raises(OSError):
        h5store.File(REAL, "r+")  # user block: not writable here


def test_known_bytes_of_a_new_file(tmp_path):
    """Superblock and datatype message, byte for byte as the specification (and every file h5py
    writes with default settings) has them."""
    p = str(tmp_path / "k.h5")
    with h5store.File(p, "w") as f:
        f.create_dataset("a", shape=(1, 3), dtype=np.float32, fillvalue=np.nan, chunks=(1, 2), maxshape=(None, 3))
    raw = open(p, "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n"
    assert raw[8:16] == bytes([0, 0, 0, 0, 0, 8, 8, 0])           # versions, 8-byte offsets / lengths
    assert raw[16:24] == bytes([4, 0, 16, 0, 0, 0, 0, 0])        # group leaf K = 4, internal K = 16, flags
    assert raw[24:32] == bytes(8) and raw[32:40] == b"\xff" * 8  # base address 0, no free-space info
    assert int.from_bytes(raw[40:48], "little") == len(raw)      # end-of-file address
    assert raw[48:56] == b"\xff" * 8                             # no driver info
    assert int.from_bytes(raw[72:76], "little") == 1             # root entry caches the symbol table
    f32 = bytes.fromhex("11201f00040000000000200017080017 7f000000".replace(" ", ""))
    assert raw.count(f32) == 1, "IEEE little-endian float32 datatype message"
    nan = np.array([np.nan], np.float32).tobytes()
    assert bytes([2, 3, 2, 1, 4, 0, 0, 0]) + nan in raw, "fill-value message v2: incremental, if-set, defined, NaN"


def _model_write(ds, ref, key, value):
    ds[key] = value
    ref[key] = value


def test_save_k_layout_is_real_hdf5(tmp_path):
    """save_K through open_store('*.h5'): the file is HDF5 with the reference's dataset layout."""
    rng = np.random.default_rng(0)
    Xs = data.ResidentDataset(torch.rand(11, 1, 4, 4))
    Xt = data.ResidentDataset(torch.rand(5, 1, 4, 4))

    def kern(x, x2, same, diag):
        if diag:
            return rng.standard_normal(len(x)).astype(np.float32)
        return rng.standard_normal((len(x), len(x2))).astype(np.float32)

    paths = []
    for r in range(3):
        p = str(tmp_path / f"w{r}.h5")
        paths.append(p)
        with block_store.open_store(p, "w") as f:
            assert isinstance(f, h5store.File)
            cnn_gp.save_K(f, kern, "Kxx", Xs, None, diag=False, batch_size=4, worker_rank=r, n_workers=3, print_interval=1e9)
            cnn_gp.save_K(f, kern, "Kxtx", Xt, Xs, diag=False, batch_size=4, worker_rank=r, n_workers=3, print_interval=1e9)
        if r == 0:  # save_kernel.py:32-36 reopens with "a" for the diagonals
            with block_store.open_store(p, "a") as f:
                cnn_gp.save_K(f, kern, "Kt_diag", Xt, None, diag=True, batch_size=4, print_interval=1e9)
    o = OracleFile(paths[0])
    assert sorted(o.datasets) == ["Kt_diag", "Kxtx", "Kxx"]
    kxx = o.datasets["Kxx"]
    assert kxx.shape == (1, 11, 11) and kxx.maxshape == (None, 11, 11) and kxx.chunks == (1, 4, 4)
    assert kxx.dtype == np.dtype("<f4") and np.isnan(kxx.fillvalue)
    assert o.datasets["Kt_diag"].shape == (1, 5) and o.datasets["Kt_diag"].chunks == (1, 4)
    # worker files together cover exactly the upper block triangle; the rest stays NaN
    seen = np.zeros((11, 11), int)
    for p in paths:
        k = OracleFile(p).datasets["Kxx"].read()[0]
        seen += ~np.isnan(k)
    blocks = np.arange(11) // 4
    np.testing.assert_array_equal(seen, (blocks[:, None] <= blocks[None, :]).astype(int))
    # merge the worker files (native chunk-wise path) == the reference's element-wise rule
    want = {n: OracleFile(paths[0]).datasets[n].read() for n in ("Kxx", "Kxtx")}
    for p in paths[1:]:
        src = OracleFile(p)
        for n in want:
            s = src.datasets[n].read()
            todo = np.isnan(want[n])
            want[n][todo] = s[todo]
    with block_store.open_store(paths[0], "a") as dest:
        for p in paths[1:]:
            with block_store.open_store(p, "r") as src:
                block_store.merge_into(dest, src)
    merged = OracleFile(paths[0])
    for n in want:
        np.testing.assert_array_equal(merged.datasets[n].read(), want[n])
    assert not np.isnan(merged.datasets["Kxtx"].read()).any()
    # classify_gp.load_kern (classify_gp.py:45-48)
    with block_store.open_store(paths[0], "r") as f:
        A = np.empty((5, 11), np.float32)
        f["Kxtx"].read_direct(A, source_sel=np.s_[0, :, :])
        np.testing.assert_array_equal(A, want["Kxtx"][0])


def test_random_hyperslabs_against_numpy(tmp_path):
    rng = np.random.default_rng(1)
    p = str(tmp_path / "r.h5")
    shape, chunks = (2, 37, 53), (1, 8, 16)
    ref32 = np.full(shape, np.nan, np.float32)
    ref64 = np.full((19, 7), -1.5, np.float64)
    refc = np.zeros((6, 5), np.float32)
    with h5store.File(p, "w") as f:
        a = f.create_dataset("a", shape=shape, dtype=np.float32, fillvalue=np.nan, chunks=chunks, maxshape=(None, 37, 53))
        b = f.create_dataset("b", shape=ref64.shape, dtype="f8", fillvalue=-1.5, chunks=(4, 7))
        c = f.create_dataset("c", shape=refc.shape, dtype=np.float32)  # contiguous, no fill value
        for _ in range(60):
            lo = [rng.integers(0, s) for s in shape]
            hi = [rng.integers(l + 1, s + 1) for l, s in zip(lo, shape)]
            key = tuple(slice(l, h) for l, h in zip(lo, hi))
            _model_write(a, ref32, key, rng.standard_normal([h - l for l, h in zip(lo, hi)]).astype(np.float32))
            i, j = rng.integers(0, 19), rng.integers(0, 7)
            _model_write(b, ref64, (slice(i, 19), j), rng.standard_normal(19 - i))
            _model_write(c, refc, (rng.integers(0, 6), slice(None)), float(rng.integers(1, 9)))
            lo = [rng.integers(0, s) for s in shape]
            key = tuple(slice(l, rng.integers(l, s + 1)) for l, s in zip(lo, shape))
            np.testing.assert_array_equal(a[key], ref32[key])
        np.testing.assert_array_equal(a[1], ref32[1])
        np.testing.assert_array_equal(a[-1, 3, ...], ref32[-1, 3])
        assert a[0, 0, 0].shape == () and len(a) == 2 and a.ndim == 3
        f.flush()
        for name, ref in (("a", ref32), ("b", ref64), ("c", refc)):  # valid after a flush, still open
            np.testing.assert_array_equal(OracleFile(p).datasets[name].read(), ref)
        _model_write(a, ref32, (0, slice(0, 37), slice(0, 53)), 7.0)
    o = OracleFile(p)
    assert o.eof_matches_size and OracleFile(REAL).eof_matches_size
    for name, ref in (("a", ref32), ("b", ref64), ("c", refc)):
        np.testing.assert_array_equal(o.datasets[name].read(), ref)
    assert o.datasets["c"].layout[0] == "contiguous" and not o.datasets["c"].fill_defined
    with h5store.File(p, "r+") as f:  # modify an existing file in place
        _model_write(f["a"], ref32, (1, slice(5, 9), slice(0, 53)), 3.0)
        np.testing.assert_array_equal(f["b"][...], ref64)
    np.testing.assert_array_equal(OracleFile(p).datasets["a"].read(), ref32)


def test_strided_views_are_transferred_in_place(tmp_path):
    """save_K_resident hands over column ranges of a wider pinned row buffer; read_direct may
    target a window of a larger array: no packing copy, same bytes."""
    rng = np.random.default_rng(3)
    p = str(tmp_path / "s.h5")
    wide = rng.standard_normal((9, 40)).astype(np.float32)
    ref = np.full((1, 9, 23), np.nan, np.float32)
    with h5store.File(p, "w") as f:
        d = f.create_dataset("K", shape=(1, 9, 23), dtype=np.float32, fillvalue=np.nan, chunks=(1, 4, 4), maxshape=(None, 9, 23))
        view = wide[:, 7:30]
        assert d._strides_of(view, [False, True, True], (9, 23)) == [0, 160, 4]
        _model_write(d, ref, (0, slice(0, 9), slice(0, 23)), view)
        _model_write(d, ref, (0, slice(2, 9, None), slice(5, 17)), wide[1:8, 20:32])
        assert d._strides_of(wide[:, ::2], [False, True, True], (9, 20)) is None  # falls back to the packing copy
        _model_write(d, ref, (0, slice(0, 9), slice(0, 20)), wide[:, ::2])
        big = np.zeros((3, 12, 30), np.float32)
        d.read_direct(big, source_sel=np.s_[0, 1:8, 3:20], dest_sel=np.s_[1, 2:9, 5:22])
        np.testing.assert_array_equal(big[1, 2:9, 5:22], ref[0, 1:8, 3:20])
        assert np.count_nonzero(big) == np.count_nonzero(big[1, 2:9, 5:22]) == 7 * 17  # nothing outside the window
    np.testing.assert_array_equal(OracleFile(p).datasets["K"].read(), ref)


def test_deep_chunk_index(tmp_path):
    """More than 64 x 64 chunks: a three-level version-1 B-tree, bulk-written, in key order."""
    p = str(tmp_path / "deep.h5")
    n = 142
    ref = np.full((1, n, n), np.nan, np.float32)
    with h5store.File(p, "w") as f:
        d = f.create_dataset("K", shape=(1, n, n), dtype=np.float32, fillvalue=np.nan, chunks=(1, 2, 2), maxshape=(None, n, n))
        row = np.arange(n * n, dtype=np.float32).reshape(n, n)
        iu = np.triu_indices(n // 2)
        for bi, bj in zip(*iu):  # upper block triangle only, like Kxx
            _model_write(d, ref, (0, slice(2 * bi, 2 * bi + 2), slice(2 * bj, 2 * bj + 2)), row[2 * bi:2 * bi + 2, 2 * bj:2 * bj + 2])
        assert d.n_chunks_stored == len(iu[0]) == 2556
    o = OracleFile(p)
    k = o.datasets["K"]
    idx = k.chunk_index()  # walks the tree and checks key order, levels and capacity
    assert len(idx) == 2556
    root_level = o.at(k.layout[1], 6)[5]
    assert root_level == 1 and 2556 > 64  # 40 leaves under one root
    np.testing.assert_array_equal(k.read(), ref)
    with h5store.File(p, "a") as f:  # reload the index, extend it, write it again
        d = f["K"]
        np.testing.assert_array_equal(d[0, 100:, :50], ref[0, 100:, :50])
        for bi in range(n // 2):
            for bj in range(bi):
                _model_write(d, ref, (0, slice(2 * bi, 2 * bi + 2), slice(2 * bj, 2 * bj + 2)), -1.0)
        assert d.n_chunks_stored == (n // 2) ** 2 == 5041
    o = OracleFile(p)
    k = o.datasets["K"]
    assert len(k.chunk_index()) == 5041 and o.at(k.layout[1], 6)[5] == 2  # 79 leaves, 2 nodes, root
    np.testing.assert_array_equal(k.read(), ref)


def test_many_datasets_and_append_mode(tmp_path):
    p = str(tmp_path / "many.h5")
    names = [f"d{i:02d}" for i in (7, 3, 11, 0, 5, 9, 1, 10, 2, 8, 4, 6)] + ["a_rather_long_dataset_name_" * 4]
    with h5store.File(p, "w") as f:
        for n in names[:5]:
            f.create_dataset(n, data=np.full((3, 2), float(len(n)), np.float32))
        with pytest.raises(OSError):
            f.create_dataset(names[0], shape=(1,), dtype=np.float32)  # exists
        with pytest.raises(TypeError):
            f.create_dataset("ints", shape=(1,), dtype=np.int32)
    for n in names[5:]:  # one reopen per dataset: symbol-table nodes and the heap grow across sessions
        with h5store.File(p, "a") as f:
            f.create_dataset(n, shape=(1, 4), dtype=np.float64, fillvalue=2.5, chunks=(1, 4), maxshape=(None, 4))
    o = OracleFile(p)  # checks link order against the B-tree keys and walks the heap free list
    assert sorted(o.datasets) == sorted(names) and len(o.groups) == 1
    with h5store.File(p, "r") as f:
        assert f.keys() == sorted(names) and len(f) == len(names)
        np.testing.assert_array_equal(f[names[0]][...], np.full((3, 2), 3.0, np.float32))
        np.testing.assert_array_equal(f[names[-1]][...], np.full((1, 4), 2.5))
        with pytest.raises(KeyError):
            f["missing"]


def test_resize_leading_dimension(tmp_path):
    """maxshape=(None, N, N2): the leading dimension the reference leaves open for more kernels."""
    p = str(tmp_path / "rs.h5")
    with h5store.File(p, "w") as f:
        d = cnn_gp.create_h5py_dataset(f, 4, "K", False, 6, 5)
        d[0, :, :] = 1.0
        with pytest.raises(IndexError):
            d[1, 0, 0]
        d.resize(3, axis=0)
        d[2, 0:4, 0:4] = 2.0
        with pytest.raises(OSError):
            d.resize((3, 7, 5))  # beyond maxshape
    k = OracleFile(p).datasets["K"]
    assert k.shape == (3, 6, 5) and k.maxshape == (None, 6, 5)
    got = k.read()
    assert (got[0] == 1).all() and np.isnan(got[1]).all() and (got[2, :4, :4] == 2).all() and np.isnan(got[2, 4:]).all()
    with h5store.File(p, "a") as f:
        f["K"].resize((1, 6, 5))
    k = OracleFile(p).datasets["K"]
    assert k.shape == (1, 6, 5) and len(k.chunk_index()) == 4


def test_modes_and_errors(tmp_path):
    p = str(tmp_path / "m.h5")
    with pytest.raises(FileNotFoundError):
        h5store.File(p, "r")
    with h5store.File(p, "w-") as f:
        f.create_dataset("x", shape=(2, 2), dtype=np.float32, fillvalue=1.0, chunks=(1, 2))
    with pytest.raises(FileExistsError):
        h5store.File(p, "x")
    with h5store.File(p, "r") as f:
        with pytest.raises(OSError):
            f.create_dataset("y", shape=(1,), dtype=np.float32)
        np.testing.assert_array_equal(f["x"][...], np.ones((2, 2), np.float32))  # nothing stored: fill value
    f = h5store.File(p, "r")
    f.close()
    with pytest.raises(ValueError):
        f.keys()
    with h5store.File(p, "w") as f:  # "w" truncates
        assert f.keys() == []
    raw = open(p, "rb").read()
    open(p, "wb").write(raw[:-8])  # shorter than the end-of-file address
    with pytest.raises(OSError):
        h5store.File(p, "r")
    with pytest.raises(H5FormatError):
        OracleFile(p)
    open(p, "wb").write(b"not hdf5" * 100)
    with pytest.raises(OSError):
        h5store.File(p, "r")


def test_merge_semantics_on_nan_only(tmp_path):
    a = block_store.open_store(str(tmp_path / "a.h5"), "w")
    b = block_store.open_store(str(tmp_path / "b.h5"), "w")
    da = a.create_dataset("K", shape=(1, 5, 5), dtype=np.float32, fillvalue=np.nan, chunks=(1, 2, 2), maxshape=(None, 5, 5))
    db = b.create_dataset("K", shape=(1, 5, 5), dtype=np.float32, fillvalue=np.nan, chunks=(1, 2, 2), maxshape=(None, 5, 5))
    b.create_dataset("only_b", shape=(1, 2), dtype=np.float32, fillvalue=np.nan, chunks=(1, 2))
    ra, rb = np.full((1, 5, 5), np.nan, np.float32), np.full((1, 5, 5), np.nan, np.float32)
    _model_write(da, ra, (0, slice(0, 3), slice(0, 5)), np.arange(15, dtype=np.float32).reshape(3, 5))
    _model_write(da, ra, (0, 1, slice(1, 3)), np.nan)             # NaN holes inside a stored chunk
    _model_write(db, rb, (0, slice(1, 5), slice(1, 5)), 100.0)
    block_store.merge_into(a, b)
    want = ra.copy()
    want[np.isnan(ra)] = rb[np.isnan(ra)]
    np.testing.assert_array_equal(a["K"][...], want)
    assert "only_b" not in a
    a.close()
    b.close()
    np.testing.assert_array_equal(OracleFile(str(tmp_path / "a.h5")).datasets["K"].read(), want)


def test_plain_c_client(tmp_path):
    """include/cnngp_h5.h is a C ABI: a C99 program (tests/c_abi/h5_roundtrip.c) writes worker
    files block by block, merges and reads them; the oracle reader then decodes the result."""
    import subprocess
    exe = str(tmp_path / "h5_roundtrip")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c_abi", "h5_roundtrip.c"), "-o", exe,
                    h5store.LIB_PATH, "-lm", "-Wl,-rpath," + os.path.dirname(h5store.LIB_PATH)], check=True)
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", (r.returncode, r.stdout, r.stderr)
    K = OracleFile(str(tmp_path / "w0.h5")).datasets["Kxx"].read()[0]
    i, j = np.indices((23, 23))
    upper = j // 5 >= i // 5
    np.testing.assert_array_equal(K[upper], (1000 * i + j)[upper].astype(np.float32))
    assert np.isnan(K[~upper]).all()

