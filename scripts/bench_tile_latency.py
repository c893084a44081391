"""Per-call cost of the drop-in API on the reference's default tile (200 x 200 images, same=False and
the diagonal same=True tile): what a user of the literal save_K loop sees per iteration."""
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "cnn-gp_b200"), ROOT]
import torch  # noqa: E402

out = {}
for cfg, C, S in (("mnist_paper_convnet_gp", 1, 28), ("mnist_as_tf", 1, 28), ("cifar10", 3, 32)):
    model = importlib.import_module("configs." + cfg).initial_model.cuda()
    x = torch.rand(200, C, S, S, device="cuda")
    z = torch.rand(200, C, S, S, device="cuda")
    for name, call in (("rect", lambda: model(x, z)), ("diag_tile", lambda: model(x, x, same=True))):
        for _ in range(5):
            call()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(50):
            call()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 50
        pairs = 200 * 200 if name == "rect" else 200 * 201 // 2
        out[f"{cfg}:{name}"] = {"ms_per_call": dt * 1e3, "Mpairs_per_s": pairs / dt / 1e6}
print(json.dumps(out, indent=1))
