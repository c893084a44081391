import os, sys
ROOT = "/root/repo"
sys.path[:0] = [os.path.join(ROOT, "cnn-gp_b200"), ROOT]
import torch
from cnn_gp import Conv2d, ReLU, Sequential, Sum, engine
def rate(m, X, reps=3):
    m = m.cuda(); n = X.shape[0]
    for _ in range(2): K = m(X)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): K = m(X)
    e1.record(); torch.cuda.synchronize()
    return 1e9 / (n * (n + 1) / 2 * reps / (e0.elapsed_time(e1) * 1e-3)), engine.plan_for(m, 28, 28, torch.float32).describe()[:60]
X = torch.rand(4000, 1, 28, 28, generator=torch.Generator().manual_seed(1)).cuda()
def body(): return [ReLU(), Conv2d(3), ReLU(), Conv2d(3)]
for nb in (4, 8):
    with_sum = Sequential(Conv2d(3), *[Sum([Sequential(), Sequential(*body())]) for _ in range(nb)], Conv2d(28, padding=0))
    plain = Sequential(Conv2d(3), *[l for _ in range(nb) for l in body()], Conv2d(28, padding=0))
    print(nb, "blocks  with Sum:", rate(with_sum, X), " without:", rate(plain, X), flush=True)
