"""GPU parity of the dense fp64 stage (cnngp_potrf_upper_f64 / cnngp_potrs_upper_f64 /
cnngp_predict_argmax, through the C ABI) against the oracle's scipy path
(reference exp_mnist_resnet/classify_gp.py:17-42) and the committed golden fixture."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _spd(n, seed, cond=1e4):
    rng = np.random.default_rng(seed)
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    ev = np.geomspace(1.0, cond, n)
    return (Q * ev) @ Q.T


@pytest.mark.parametrize("n", [1, 2, 5, 64, 127, 128, 129, 255, 300, 513, 1000, 1537])
def test_potrf_matches_lapack(n):
    from cnn_gp import linalg
    K = _spd(n, n)
    K = (K + K.T) / 2
    want = np.linalg.cholesky(K).T  # upper
    A = torch.from_numpy(np.triu(K) + np.tril(np.full((n, n), np.nan), -1)).cuda()  # lower triangle is never read
    assert linalg.potrf_upper_(A) == 0
    got = A.cpu().numpy()
    iu = np.triu_indices(n)
    np.testing.assert_allclose(got[iu], want[iu], rtol=0, atol=1e-11 * np.abs(want).max())
    il = np.tril_indices(n, -1)
    assert np.isnan(got[il]).all(), "the strictly lower triangle must not be touched"


@pytest.mark.parametrize("n,nrhs", [(1, 1), (7, 3), (128, 10), (300, 10), (777, 1), (1000, 16), (1025, 37)])
def test_solve_matches_scipy(n, nrhs):
    from cnn_gp import linalg
    from oracle import oracle
    K = _spd(n, 100 + n)
    K = (K + K.T) / 2
    rng = np.random.default_rng(n)
    Y = rng.standard_normal((n, nrhs))
    want = oracle.solve_system(np.triu(K), Y)
    got = linalg.solve_pos_upper(torch.from_numpy(np.triu(K)).cuda(), torch.from_numpy(Y).cuda()).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-10 * np.abs(want).max())
    # backward error of our own solution
    assert np.abs(K @ got - Y).max() <= 1e-9 * np.abs(K).max() * np.abs(got).max()


def test_strided_view_and_jitter():
    """A sub-block of a larger allocation (lda > n, odd lda: the unvectorised load path)."""
    from cnn_gp import linalg
    n = 333
    K = _spd(n, 5)
    K = (K + K.T) / 2
    big = torch.full((n + 3, n + 4), float("nan"), dtype=torch.float64, device="cuda")
    view = big[1:n + 1, 2:n + 2]
    view.copy_(torch.from_numpy(K))
    assert linalg.potrf_upper_(view) == 0
    want = np.linalg.cholesky(K).T
    iu = np.triu_indices(n)
    np.testing.assert_allclose(view.cpu().numpy()[iu], want[iu], rtol=0, atol=1e-11 * np.abs(want).max())
    assert torch.isnan(big[0]).all() and torch.isnan(big[:, :2]).all() and torch.isnan(big[:, n + 2:]).all()


def test_not_positive_definite_reports_lapack_info():
    from cnn_gp import linalg
    n = 400
    K = _spd(n, 9)
    K = (K + K.T) / 2
    K[250, 250] = -1.0
    with pytest.raises(linalg.NotPositiveDefiniteError) as ei:
        linalg.potrf_upper_(torch.from_numpy(K).cuda())
    want_info = None
    try:
        np.linalg.cholesky(K)
    except np.linalg.LinAlgError:
        import scipy.linalg
        _, want_info = scipy.linalg.lapack.dpotrf(K, lower=0)
    assert ei.value.info == want_info == 251


def test_golden_solve_and_decisions():
    """Same Kxx bytes as the reference run: weights to 1e-9, decisions identical."""
    from cnn_gp import linalg
    g = np.load(os.path.join(GOLD, "solve.npz"))
    K = torch.from_numpy(np.triu(g["Kxx"].astype(np.float64))).cuda()
    A = linalg.solve_pos_upper(K, torch.from_numpy(g["Y"]).cuda())
    np.testing.assert_allclose(A.cpu().numpy(), g["A"], rtol=0, atol=1e-9 * np.abs(g["A"]).max())
    pred, scores = linalg.predict_argmax(torch.from_numpy(g["Kxtx"]).cuda(), A, return_scores=True)
    np.testing.assert_allclose(scores.cpu().numpy(), g["F"], rtol=0, atol=1e-9 * np.abs(g["F"]).max())
    np.testing.assert_array_equal(pred.cpu().numpy(), g["pred"])


@pytest.mark.parametrize("R,n,c", [(1, 1, 1), (37, 1000, 10), (100, 333, 17), (9, 4097, 40)])
def test_predict_argmax(R, n, c):
    from cnn_gp import linalg
    rng = np.random.default_rng(R * n)
    K = rng.standard_normal((R, n)).astype(np.float32)
    A = rng.standard_normal((n, c))
    want = K.astype(np.float64) @ A
    pred, scores = linalg.predict_argmax(torch.from_numpy(K).cuda(), torch.from_numpy(A).cuda(), return_scores=True)
    np.testing.assert_allclose(scores.cpu().numpy(), want, rtol=0, atol=1e-12 * np.abs(want).max() * np.sqrt(n))
    np.testing.assert_array_equal(pred.cpu().numpy(), want.argmax(1))
    # a genuinely float64 kernel block goes through the same kernel (no library matmul anywhere on this path)
    K64 = K.astype(np.float64) + 1e-9 * rng.standard_normal(K.shape)
    pred64, scores64 = linalg.predict_argmax(torch.from_numpy(K64).cuda(), torch.from_numpy(A).cuda(), return_scores=True)
    np.testing.assert_allclose(scores64.cpu().numpy(), K64 @ A, rtol=0, atol=1e-12 * np.abs(K64 @ A).max() * np.sqrt(n))
    np.testing.assert_array_equal(pred64.cpu().numpy(), (K64 @ A).argmax(1))
    np.testing.assert_array_equal(linalg.predict_argmax(torch.from_numpy(K).cuda(), torch.from_numpy(A).cuda()).cpu().numpy(),
                                  want.argmax(1))


def test_large_factorisation_residual():
    """n = 6000: many block columns, look-ahead on the side stream; checked by the residual
    ||U^T U - K|| and against a float64 torch solve."""
    from cnn_gp import linalg
    n = 6000
    g = torch.Generator(device="cuda").manual_seed(3)
    B = torch.randn(n, n + 50, generator=g, device="cuda", dtype=torch.float64)
    K = B @ B.T / n + 0.1 * torch.eye(n, device="cuda", dtype=torch.float64)
    U = torch.triu(K)
    assert linalg.potrf_upper_(U) == 0
    U = torch.triu(U)
    res = (U.T @ U - K).abs().max().item() / K.abs().max().item()
    assert res < 1e-13, res
    Y = torch.randn(n, 10, generator=g, device="cuda", dtype=torch.float64)
    X = linalg.potrs_upper_(U, Y.clone())
    assert ((K @ X - Y).abs().max() / (K.abs().max() * X.abs().max())).item() < 1e-13


def test_cpu_tensors_are_rejected():
    from cnn_gp import linalg
    with pytest.raises(RuntimeError):
        linalg.potrf_upper_(torch.eye(3, dtype=torch.float64))


def test_distributed_driver_on_one_gpu_matches_single_gpu_factor():
    """cnn_gp.linalg_dist with world size 1 (NCCL): the building-block entry points
    cnngp_potrf_panel_f64 / cnngp_syrk_upper_f64 must reproduce cnngp_potrf_upper_f64 bit for bit,
    NaN below the diagonal untouched, LAPACK info on failure."""
    import socket
    import torch.distributed as dist
    from cnn_gp import linalg, linalg_dist
    created = False
    if not dist.is_initialized():
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1,
                                device_id=torch.device("cuda", 0))
        created = True
    try:
        dev = torch.device("cuda", 0)
        for n in (100, 257, 1000, 1537):
            K = _spd(n, 40 + n)
            K = (K + K.T) / 2
            Kd = torch.from_numpy(np.triu(K) + np.tril(np.full((n, n), np.nan), -1)).to(dev)
            Y = torch.from_numpy(np.random.default_rng(n).standard_normal((n, 4))).to(dev)
            A = linalg_dist.solve_pos_upper_distributed(Kd, Y, n, dev)
            ch = linalg_dist.DistributedCholesky(n, dev)
            ch.scatter_from(Kd)
            assert ch.factorize() == 0
            U = ch.gather_to(0)
            U1 = Kd.clone()
            linalg.potrf_upper_(U1)
            iu = np.triu_indices(n)
            np.testing.assert_array_equal(U.cpu().numpy()[iu], U1.cpu().numpy()[iu])
            want = np.linalg.solve(K, Y.cpu().numpy())
            np.testing.assert_allclose(A.cpu().numpy(), want, rtol=0, atol=1e-10 * np.abs(want).max())
        K = _spd(600, 3)
        K = (K + K.T) / 2
        K[300, 300] = -2.0
        with pytest.raises(linalg.NotPositiveDefiniteError) as ei:
            linalg_dist.solve_pos_upper_distributed(torch.from_numpy(K).to(dev), torch.zeros(600, 1, dtype=torch.float64, device=dev), 600, dev)
        assert ei.value.info == 301
    finally:
        if created:
            dist.destroy_process_group()
