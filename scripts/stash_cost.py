"""What the skip connection of an identity block costs inside the fused-net kernel: blocks of
ReLU conv ReLU conv with a Sum, 4 and 8 of them -> ns per pair and block.  With CNNGP_LIB pointing at a build
with -DCNNGP_EXP_SKIP=1/2/3 (scripts/build_variant.sh) the block leaves out the stash / the add / both
(wrong results, timing only).   usage: python scripts/stash_cost.py [N]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "cnn-gp_b200"), ROOT]
import torch  # noqa: E402
from cnn_gp import Conv2d, ReLU, Sequential, Sum, engine  # noqa: E402


def rate(m, X, reps=3):
    m = m.cuda()
    n = X.shape[0]
    for _ in range(2):
        m(X)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        m(X)
    e1.record()
    torch.cuda.synchronize()
    return 1e9 / (n * (n + 1) / 2 * reps / (e0.elapsed_time(e1) * 1e-3))


n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
X = torch.rand(n, 1, 28, 28, generator=torch.Generator().manual_seed(1)).cuda()
r = {}
for nb in (4, 8):
    m = Sequential(Conv2d(3), *[Sum([Sequential(), Sequential(ReLU(), Conv2d(3), ReLU(), Conv2d(3))]) for _ in range(nb)], Conv2d(28, padding=0))
    r[nb] = rate(m, X)
print(os.environ.get("CNNGP_LIB", "default"), engine.plan_for(m, 28, 28, torch.float32).describe()[:40], "ns per pair and block: %.3f" % ((r[8] - r[4]) / 4))
