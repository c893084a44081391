"""Throughput of the blocked fp64 Cholesky / solve / prediction on the GPU box.
usage: python scripts/bench_solve.py [--cpu] [n ...]   -> one JSON line per n
--cpu also times the reference's solve, scipy.linalg.solve(assume_a='pos', lower=False)
(exp_mnist_resnet/classify_gp.py:24-26) through the oracle on the box's host cores, for n <= 8192
(larger sizes are extrapolated with n^3 from the largest one timed, and say so)."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "cnn-gp_b200"), ROOT]
import torch  # noqa: E402
from cnn_gp import linalg  # noqa: E402


def dmma_peak_tflops():
    L = ctypes.CDLL(os.path.join(ROOT, "cnn-gp_b200", "libcnngp_bench.so"))
    L.mb_probe.restype = ctypes.c_double
    L.mb_probe.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
    return 2.0 * max(L.mb_probe(20, 4, 2000), L.mb_probe(20, 8, 2000)) / 1e12, 2.0 * L.mb_probe(21, 8, 2000) / 1e12


def cpu_solve_seconds(K, Y):
    """The reference's CPU solve on the same matrix (oracle.solve_system = scipy posv, upper)."""
    import time
    import numpy as np
    from oracle import oracle
    Kh, Yh = np.triu(K.cpu().numpy()), Y.cpu().numpy()
    t0 = time.perf_counter()
    oracle.solve_system(Kh, Yh)
    return time.perf_counter() - t0


def main():
    args = [a for a in sys.argv[1:] if a != "--cpu"]
    with_cpu = "--cpu" in sys.argv[1:]
    ns = [int(a) for a in args] or [8192, 16384, 32768]
    peak, dfma = dmma_peak_tflops()
    cpu_ref = None  # (n, seconds) of the largest CPU solve timed
    for n in ns:
        g = torch.Generator(device="cuda").manual_seed(n)
        K = torch.empty((n, n), dtype=torch.float64, device="cuda")
        # SPD by diagonal dominance-ish low-rank + ridge, built without an n^3 product
        B = torch.randn(n, 64, generator=g, device="cuda", dtype=torch.float64)
        torch.mm(B, B.T, out=K)
        K.diagonal().add_(1.0)
        Y = torch.randn(n, 10, generator=g, device="cuda", dtype=torch.float64)
        times = {}
        for rep in range(2):
            U = K.clone()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            info = linalg.potrf_upper_(U, check=False)
            e[1].record()
            X = linalg.potrs_upper_(U, Y.clone())
            e[2].record()
            torch.cuda.synchronize()
            times = {"potrf_ms": e[0].elapsed_time(e[1]), "potrs_ms": e[1].elapsed_time(e[2])}
        assert int(info.item()) == 0
        r = K @ X - Y
        Kp = torch.randn(4096, n, generator=g, device="cuda", dtype=torch.float32)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        linalg.predict_argmax(Kp, X)
        e0.record()
        linalg.predict_argmax(Kp, X)
        e1.record()
        torch.cuda.synchronize()
        tf = n ** 3 / 3 / (times["potrf_ms"] * 1e-3) / 1e12
        if with_cpu:
            if n <= 8192:
                cpu_ref = (n, cpu_solve_seconds(K, Y))
                times["cpu_scipy_solve_s"] = cpu_ref[1]
                times["cpu_cores"] = os.cpu_count()
            elif cpu_ref:
                times["cpu_scipy_solve_s_extrapolated"] = cpu_ref[1] * (n / cpu_ref[0]) ** 3
                times["cpu_extrapolated_from_n"] = cpu_ref[0]
                times["cpu_cores"] = os.cpu_count()
        print(json.dumps({"n": n, **times, "potrf_tflops": tf, "dmma_peak_tflops": peak, "dfma_peak_tflops": dfma,
                          "frac_of_dmma_peak": tf / peak, "residual": float(r.abs().max() / (X.abs().max() * K.abs().max())),
                          "predict_ms_4096rows": e0.elapsed_time(e1),
                          "predict_GBs": 4096 * n * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9}))
        del K, U, B, Kp


if __name__ == "__main__":
    main()
