"""Where the fused-net kernel's time goes: rates of truncated ResNet GPs (stage 1 only, stages 1-2,
all three stages) -> ns per pair of each stage.   usage: python scripts/fnet_phases.py [N] [S0]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "cnn-gp_b200"), ROOT]
import torch  # noqa: E402
from cnn_gp import Conv2d, ReLU, Sequential, resnet_block, engine  # noqa: E402


def model(stages, S0, blocks=5):
    mods, size = [Conv2d(kernel_size=3)], S0
    for stage, stride in enumerate((1, 2, 2)[:stages]):
        mods.append(resnet_block(stride=stride, projection_shortcut=True))
        mods += [resnet_block(stride=1) for _ in range(blocks - 1)]
        size //= stride
    mods.append(Conv2d(kernel_size=size, padding=0))
    return Sequential(*mods)


def rate(m, X, reps=3):
    m = m.cuda()
    n = X.shape[0]
    for _ in range(2):
        K = m(X)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        K = m(X)
    e1.record()
    torch.cuda.synchronize()
    assert bool(torch.isfinite(K).all())
    return n * (n + 1) / 2 * reps / (e0.elapsed_time(e1) * 1e-3), engine.last_path()


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    S0 = int(sys.argv[2]) if len(sys.argv) > 2 else 28
    C = 3 if S0 == 32 else 1
    X = torch.rand(n, C, S0, S0, generator=torch.Generator().manual_seed(1)).cuda()
    out, prev = {}, 0.0
    for name, m in (("stem+dense", Sequential(Conv2d(kernel_size=3), Conv2d(kernel_size=S0, padding=0))),
                    ("stage1", model(1, S0)), ("stage1-2", model(2, S0)), ("stage1-3", model(3, S0)),
                    ("stage1 x2 blocks", model(1, S0, blocks=9))):
        r, path = rate(m, X)
        ns = 1e9 / r
        out[name] = {"Mpairs_s": round(r / 1e6, 2), "ns_per_pair": round(ns, 3), "path": path}
        print(name, out[name], flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
