"""Time and cross-check the fused-kernel variants (CNNGP_FUSED_VARIANT=nw,nsplit,nst) on the headline program."""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "cnn-gp_b200"), ROOT]
import torch  # noqa: E402
from cnn_gp import engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
variants = sys.argv[2:] or ["8,1,2", "8,2,4", "8,4,8", "12,2,2", "12,4,4"]
model = importlib.import_module("configs.mnist_paper_convnet_gp").initial_model.cuda()
g = torch.Generator(device="cuda").manual_seed(1)
X = torch.rand(n, 1, 28, 28, device="cuda", generator=g)
plan = engine.plan_for(model, 28, 28, torch.float32)
aux, _, kd = engine.variances(plan, X)
ref = None
res = {}
for v in variants:
    os.environ["CNNGP_FUSED_VARIANT"] = v
    out = torch.full((n, n), float("nan"), device="cuda")
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        engine.gram_with_aux(plan, X, X, aux, aux, True, False, True, out=out, kdiag=kd)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    # rectangular, ragged call too
    Kr = torch.full((777, 1001), float("nan"), device="cuda")
    engine.gram_with_aux(plan, X[:777], X[1000:2001], aux[:777], aux[1000:2001], False, False, False, out=Kr)
    torch.cuda.synchronize()
    if ref is None:
        ref = (out.clone(), Kr.clone())
        err = 0.0
    else:
        err = max(float(((out - ref[0]).abs() / ref[0].abs()).max()), float(((Kr - ref[1]).abs() / ref[1].abs()).max()))
    assert torch.isfinite(out).all() and torch.isfinite(Kr).all()
    res[v] = {"ms": best, "Mpairs_per_s": n * (n + 1) / 2 / best / 1e3, "max_rel_diff_vs_first": err}
    del out
print(json.dumps({"n": n, "variants": res}, indent=1))
