"""Accuracy of the float32 kernels against the float64 generic kernel (run on the GPU box)."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "cnn-gp_b200"), ROOT]
import torch  # noqa: E402
from cnn_gp import engine  # noqa: E402

for cfg, C, S in (("mnist_paper_convnet_gp", 1, 28), ("mnist_paper_residual_cnn_gp", 1, 28), ("mnist_as_tf", 1, 28),
                  ("cifar10", 3, 32)):
    model = importlib.import_module("configs." + cfg).initial_model
    gen = torch.Generator().manual_seed(1)
    for kind in ("rand", "randn", "near-dup"):
        X = torch.rand(96, C, S, S, generator=gen) if kind != "randn" else torch.randn(96, C, S, S, generator=gen)
        Z = torch.rand(80, C, S, S, generator=gen) if kind != "randn" else torch.randn(80, C, S, S, generator=gen)
        if kind == "near-dup":
            Z[:40] = X[:40] * (1 + 1e-3 * torch.randn(40, 1, 1, 1, generator=gen)) + 1e-3 * torch.rand(40, C, S, S, generator=gen)
        ref = model.double().cuda()(X.double().cuda(), Z.double().cuda())
        m32 = model.float().cuda()
        out = {}
        for path in ("generic", "auto"):
            engine.set_path(path)
            K = m32(X.cuda(), Z.cuda()).double()
            out[path + ":" + engine.last_path()] = float(((K - ref).abs() / ref.abs()).max())
        engine.set_path("auto")
        print(cfg, kind, {k: f"{v:.2e}" for k, v in out.items()})
