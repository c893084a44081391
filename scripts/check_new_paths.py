"""Small invocations of the paths added late in round 2 -- variance kernel on an odd image count, band launches
(straight-line and fused-net programs), the streamed host call -- with bit-equality checks against the per-row /
device-resident forms.  Sized for compute-sanitizer (`--tool memcheck`), which is closed on the gpurun pool: run
plainly there.                                          usage: python scripts/sanitize_new_paths.py"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "cnn-gp_b200"), ROOT]
import torch  # noqa: E402
from cnn_gp import engine  # noqa: E402
from cnn_gp.tiles import GramJob, compute_worker_blocks  # noqa: E402

for cfg, C, S, n, bs in (("mnist_paper_convnet_gp", 1, 28, 131, 24), ("mnist_as_tf", 1, 28, 70, 12)):
    model = importlib.import_module("configs." + cfg).initial_model.cuda()
    X = torch.rand(n, C, S, S, generator=torch.Generator().manual_seed(1)).cuda()
    job = GramJob(model, X)                       # variance rows (odd image count)
    for rank in range(2):
        a = torch.full((n, n), float("nan"), device="cuda")
        b = torch.full((n, n), float("nan"), device="cuda")
        compute_worker_blocks(job, a, bs, rank, 2, balanced=True)
        compute_worker_blocks(job, b, bs, rank, 2, balanced=True, rows_per_launch=1)
        torch.cuda.synchronize()
        assert torch.equal(a.view(torch.int32), b.view(torch.int32))
    Kh = model(X.cpu().pin_memory())              # streamed launch with progress counters
    Kd = model(X)
    torch.cuda.synchronize()
    assert torch.equal(Kh, Kd.cpu())
    print(cfg, "ok", engine.last_path(), flush=True)
