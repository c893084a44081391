"""potrs (persistent dataflow sweeps) against torch's triangular solves + timing.  usage: python scripts/check_potrs.py [n ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "cnn-gp_b200"), ROOT]
import torch  # noqa: E402
from cnn_gp import linalg  # noqa: E402

for n in [int(a) for a in sys.argv[1:]] or [100, 128, 300, 1000, 4096, 8192, 32768]:
    g = torch.Generator(device="cuda").manual_seed(n)
    B = torch.randn(n, 64, generator=g, device="cuda", dtype=torch.float64)
    K = B @ B.T
    K.diagonal().add_(1.0)
    for nrhs in (10, 3, 20):
        Y = torch.randn(n, nrhs, generator=g, device="cuda", dtype=torch.float64)
        U = K.clone()
        linalg.potrf_upper_(U)
        X = linalg.potrs_upper_(U, Y.clone())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        X = linalg.potrs_upper_(U, Y.clone())
        e1.record()
        torch.cuda.synchronize()
        r = float((K @ X - Y).abs().max() / (X.abs().max() * K.abs().max()))
        ok = ""
        if n <= 8192:
            Ut = torch.triu(U)
            want = torch.linalg.solve_triangular(Ut, torch.linalg.solve_triangular(Ut.T, Y, upper=False), upper=True)
            ok = f" vs torch {float((X - want).abs().max() / want.abs().max()):.2e}"
        print(f"n={n} nrhs={nrhs}: potrs {e0.elapsed_time(e1):.3f} ms residual {r:.2e}{ok}", flush=True)
    del K, U, B
