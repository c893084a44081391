"""Per-call times of the device-resident Gram launch(es) of one config: python scripts/gram_times.py CONFIG N [REPS]
(variance rows computed once; every call timed by its own CUDA events)"""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "cnn-gp_b200"), ROOT]
import torch  # noqa: E402
from cnn_gp import engine  # noqa: E402

cfg, n = sys.argv[1], int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
C, S = (3, 32) if cfg == "cifar10" else (1, 28)
model = importlib.import_module("configs." + cfg).initial_model.cuda()
X = torch.rand(n, C, S, S, device="cuda")
plan = engine.plan_for(model, S, S, torch.float32)
aux, _, kd = engine.variances(plan, X)
out = torch.empty((n, n), device="cuda")
ms = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    engine.gram_with_aux(plan, X, X, aux, aux, True, False, True, out=out, kdiag=kd)
    e1.record()
    torch.cuda.synchronize()
    ms.append(round(e0.elapsed_time(e1), 2))
best = min(ms)
print(json.dumps({"config": cfg, "n": n, "plan": plan.describe()[:48], "ms": ms, "best_Mpairs_per_s": round(n * (n + 1) / 2 / best / 1e3, 2)}))
