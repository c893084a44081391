"""Sweep the super-tile edge (CNNGP_SUPER_EDGE) of the fused Gram kernels: pairs/s per setting."""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "cnn-gp_b200"), ROOT]
import torch  # noqa: E402
from cnn_gp import engine  # noqa: E402

cfg, n = sys.argv[1], int(sys.argv[2])
C, S = (3, 32) if cfg == "cifar10" else (1, 28)
model = importlib.import_module("configs." + cfg).initial_model.cuda()
X = torch.rand(n, C, S, S, device="cuda")
plan = engine.plan_for(model, S, S, torch.float32)
aux, _, kd = engine.variances(plan, X)
out = torch.empty((n, n), device="cuda")
res = {}
for edge in [int(a) for a in sys.argv[3:]]:
    os.environ["CNNGP_SUPER_EDGE"] = str(edge)
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        engine.gram_with_aux(plan, X, X, aux, aux, True, False, True, out=out, kdiag=kd)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res[edge] = {"ms": best, "Mpairs_per_s": n * (n + 1) / 2 / best / 1e3}
print(json.dumps({"config": cfg, "n": n, "path": engine.last_path(), "sweep": res}))
