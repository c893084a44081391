"""Pin the CPU oracle (oracle/) against the golden vectors produced by the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import glob
import json
import os

import numpy as np
import pytest

from oracle import oracle
from models import golden_models

GOLD = os.path.join(os.path.dirname(__file__), "golden")
MODELS = golden_models()

# The oracle sums conv taps in a different order than ATen's conv; the reference's own
# fp32-vs-fp64 noise is ~1e-6, so fp32 agreement is checked at 5e-6.
TOL = {"f32": 5e-6, "f64": 1e-12}


def rel_err(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), np.finfo(np.float64).tiny)))


@pytest.mark.parametrize("name", sorted(MODELS))
@pytest.mark.parametrize("tag,dt", [("f32", np.float32), ("f64", np.float64)])
def test_gram_matches_reference(name, tag, dt):
    g = np.load(os.path.join(GOLD, f"gram_{name}.npz"))
    m = MODELS[name]
    X, Z = g["X"].astype(dt), g["Z"].astype(dt)
    n = min(len(X), len(Z))
    calls = {
        "Kxx": lambda: oracle.gram(m, X),
        "Kxz": lambda: oracle.gram(m, X, Z),
        "Kxx_diag": lambda: oracle.gram(m, X, diag=True),
        "Kxz_diag": lambda: oracle.gram(m, X[:n], Z[:n], diag=True),
        "Kxz_same": lambda: oracle.gram(m, X[:n], Z[:n], same=True),
        "Kxz_same_diag": lambda: oracle.gram(m, X[:n], Z[:n], same=True, diag=True),
    }
    for key, fn in calls.items():
        got, want = fn(), g[f"{key}_{tag}"]
        assert got.shape == want.shape and got.dtype == dt
        assert rel_err(got, want) < TOL[tag], (name, key, tag)


def test_readme_known_answers():
    """The smoke vector quoted in SURVEY.md section 4 (config 1, seed 0)."""
    g = np.load(os.path.join(GOLD, "gram_readme.npz"))
    assert abs(g["X"][0, 0, 0, 0] - (-1.1258398294)) < 1e-6
    K = oracle.gram(MODELS["readme"], g["X"].astype(np.float64))
    np.testing.assert_allclose(K[0, 0], 0.23610295140249518, rtol=1e-12)
    np.testing.assert_allclose(K[0, 1], 0.11221043270367054, rtol=1e-12)
    np.testing.assert_allclose(K[1, 1], 0.21488433768972875, rtol=1e-12)
    np.testing.assert_array_equal(K, K.T)


def test_zero_image_is_tiny_not_nan():
    """All-zero image, zero bias: ~1.7e-20, neither 0 nor NaN (f32_tiny, kernels.py:133,146)."""
    m = MODELS["readme"]
    X = np.zeros((1, 3, 28, 28), np.float32)
    k = oracle.gram(m, X, X.copy(), same=False)
    assert np.isfinite(k).all() and 0 < k[0, 0] < 1e-15


def test_conv_nonsquare_even_kernel():
    g = np.load(os.path.join(GOLD, "conv_nonsquare.npz"))
    from cnn_gp import Conv2d
    conv = Conv2d(4, stride=2, var_weight=0.9, var_bias=0.2)
    out = oracle._conv_maps(g["xy"][:, 0], conv, np.dtype(np.float64))
    np.testing.assert_allclose(out, g["out"][:, 0], rtol=1e-13)


def test_same_needs_equal_lengths():
    m = MODELS["readme"]
    X = np.random.rand(3, 3, 28, 28).astype(np.float32)
    with pytest.raises(RuntimeError):
        oracle.gram(m, X, X[:2].copy(), same=True)
    with pytest.raises(AssertionError):
        oracle.gram(m, X, X[:2].copy(), diag=True)


def test_tiles_match_reference():
    with open(os.path.join(GOLD, "tiles.json")) as f:
        cases = json.load(f)
    assert len(cases) > 50
    for c in cases:
        got = oracle.worker_tiles(c["N"], c["N2"], c["bs"], c["rank"], c["n_workers"])
        assert [list(map(int, t)) for t in got] == c["tiles"], c


def test_save_k_layout_matches_reference():
    g = np.load(os.path.join(GOLD, "save_k_layout.npz"))
    m = MODELS["readme"]
    Xs, Xt = g["Xs"], g["Xt"]
    for nw in (1, 3):
        for r in range(nw):
            for name, (A, B, diag) in {"Kxx": (Xs, None, False), "Kxtx": (Xt, Xs, False)}.items():
                want = g[f"{name}_nw{nw}_r{r}"]
                got = oracle.save_k_blocks(m, A, B, diag, 4, r, nw)
                assert got.shape == want.shape
                np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
                np.testing.assert_allclose(got[~np.isnan(got)], want[~np.isnan(want)], rtol=5e-6)
    want = g["Kt_diag_nw1_r0"]
    got = oracle.save_k_blocks(m, Xt, None, True, 4)
    np.testing.assert_allclose(got, want, rtol=5e-6)


def test_solve_matches_reference():
    g = np.load(os.path.join(GOLD, "solve.npz"))
    K = g["Kxx"].astype(np.float64)
    A = oracle.solve_system(np.triu(K), g["Y"])
    np.testing.assert_allclose(A, g["A"], rtol=1e-9, atol=1e-12 * np.abs(g["A"]).max())
    pred = oracle.predict(g["Kxtx"].astype(np.float64), A)
    np.testing.assert_array_equal(pred, g["pred"])
