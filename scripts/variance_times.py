"""Time of cnngp_variances (per-image variance rows) for one config, streaming kernel vs interpreter:
python scripts/variance_times.py CONFIG N [REPS]   (CUDA events; bytes = rows written + images read)"""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "cnn-gp_b200"), ROOT]
import torch  # noqa: E402
from cnn_gp import engine  # noqa: E402

cfg, n = sys.argv[1], int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
C, S = (3, 32) if cfg == "cifar10" else (1, 28)
model = importlib.import_module("configs." + cfg).initial_model.cuda()
X = torch.rand(n, C, S, S, device="cuda")
plan = engine.plan_for(model, S, S, torch.float32)
res = {"config": cfg, "n": n, "row_bytes": plan.aux_elems * 4}
for label, env in (("streaming", None), ("interpreter", "1")):
    if env:
        os.environ["CNNGP_VARIANCE_GENERIC"] = env
    ms = []
    for _ in range(reps + 2):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        engine.variances(plan, X)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    best = min(ms[2:])
    res[label] = {"ms": round(best, 4), "GBps": round((plan.aux_elems * 4 + C * S * S * 4) * n / best / 1e6, 1)}
print(json.dumps(res))
