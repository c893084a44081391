"""Per-call times of the triangular sweeps (potrs) on an SPD factor: python scripts/potrs_times.py [n ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "cnn-gp_b200"), ROOT]
import torch  # noqa: E402
from cnn_gp import linalg  # noqa: E402

for n in [int(a) for a in sys.argv[1:]] or [4096, 32768]:
    g = torch.Generator(device="cuda").manual_seed(n)
    B = torch.randn(n, 64, generator=g, device="cuda", dtype=torch.float64)
    K = B @ B.T
    K.diagonal().add_(1.0)
    U = K.clone()
    linalg.potrf_upper_(U)
    for nrhs in (10, 16, 20):
        Y = torch.randn(n, nrhs, generator=g, device="cuda", dtype=torch.float64)
        ms = []
        for _ in range(6):
            Yc = Y.clone()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            X = linalg.potrs_upper_(U, Yc)
            e1.record()
            torch.cuda.synchronize()
            ms.append(round(e0.elapsed_time(e1), 3))
        r = float((K @ X - Y).abs().max() / (X.abs().max() * K.abs().max()))
        print(f"{os.environ.get('CNNGP_LIB', 'default')} n={n} nrhs={nrhs}: ms {ms} residual {r:.2e}", flush=True)
    del K, U, B
