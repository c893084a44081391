"""Per-call times of predict_argmax (float32 kernel block x float64 weights): python scripts/predict_times.py [R n ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "cnn-gp_b200"), ROOT]
import torch  # noqa: E402
from cnn_gp import linalg  # noqa: E402

args = [int(a) for a in sys.argv[1:]] or [4096, 32768, 10000, 60000]
for R, n in zip(args[0::2], args[1::2]):
    g = torch.Generator(device="cuda").manual_seed(n)
    K = torch.randn(R, n, generator=g, device="cuda", dtype=torch.float32)
    A = torch.randn(n, 10, generator=g, device="cuda", dtype=torch.float64)
    want = (K.double() @ A)
    ms = []
    for _ in range(6):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pred, scores = linalg.predict_argmax(K, A, return_scores=True)
        e1.record()
        torch.cuda.synchronize()
        ms.append(round(e0.elapsed_time(e1), 3))
    err = float((scores - want).abs().max() / want.abs().max())
    same = bool((pred == want.argmax(1)).all())
    print(f"R={R} n={n}: ms {ms} -> {R * n * 4 / min(ms) / 1e6:.0f} GB/s of K, scores vs torch {err:.1e}, argmax identical {same}", flush=True)
