"""Per-layer cost of the two register-resident kernels on the same arithmetic: a straight-line stack of
ReLU + 3x3 convolutions on the straight-line kernel (default) or the fused-net kernel (CNNGP_NO_FUSED=1),
and the same stack with identity shortcuts (Sum).   usage: python scripts/fnet_vs_fused.py [N]"""
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
    return 1e9 / (n * (n + 1) / 2 * reps / (e0.elapsed_time(e1) * 1e-3)), engine.plan_for(m, 28, 28, torch.float32).describe()[:44]


def body():
    return [ReLU(), Conv2d(3), ReLU(), Conv2d(3)]


n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
X = torch.rand(n, 1, 28, 28, generator=torch.Generator().manual_seed(1)).cuda()
res = {}
for nb in (4, 8):
    plain = Sequential(Conv2d(3), *[l for _ in range(nb) for l in body()], Conv2d(28, padding=0))
    with_sum = Sequential(Conv2d(3), *[Sum([Sequential(), Sequential(*body())]) for _ in range(nb)], Conv2d(28, padding=0))
    res[nb] = (rate(plain, X), rate(with_sum, X))
    print(nb, "blocks: plain", res[nb][0], " with Sum", res[nb][1], flush=True)
print("ns per pair and block: plain %.3f, with Sum %.3f" % ((res[8][0][0] - res[4][0][0]) / 4, (res[8][1][0] - res[4][1][0]) / 4))
