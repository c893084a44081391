#!/usr/bin/env python
"""Times the UNMODIFIED reference on the host cores -- TEST / BENCH INFRASTRUCTURE ONLY.

    python oracle/ref_cpu.py CONFIG --steps K --warmup W --seconds S [--edge E]

Runs in its own process because the reference package is also called ``cnn_gp``: only
oracle/_ref (the byte-for-byte copy made by oracle/make_ref.sh) is put on sys.path here, never
this repository's drop-in package.  One step is the reference's own hot call
(exp_mnist_resnet/save_kernel.py:21-24 without the .cuda()):

    with torch.no_grad(): config.initial_model(X_tile, Z_tile, same=False)

in float32 with torch.set_num_threads(all host cores), on a tile of at most 200 x 200 images (the
reference's default tile, save_kernel.py:43) of the same synthetic workload bench.py uses
(torch.rand, seed 1234).  Prints one JSON line: pairs/s per step, cores, the sample.
"""
import argparse
import importlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.isdir(os.path.join(REF, "cnn_gp")) and os.path.isdir(os.path.join(REF, "configs"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--seconds", type=float, default=4.0, help="target time of one step")
    ap.add_argument("--edge", type=int, default=0, help="tile edge (0: from --seconds, at most 200)")
    a = ap.parse_args()
    if not available():
        print(json.dumps({"unavailable": "oracle/_ref is missing (oracle/make_ref.sh needs /root/reference)"}))
        return
    # only the reference's packages: drop every path that could serve this repository's cnn_gp / configs
    root = os.path.dirname(HERE)
    sys.path[:] = [REF] + [p for p in sys.path if p and os.path.abspath(p) not in (root, os.path.join(root, "cnn-gp_b200"), HERE)]
    import numpy as np
    if not hasattr(np, "int"):
        np.int = int  # cnn_gp/data.py:12 (removed alias); the reference file itself stays untouched
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)  # torchrun exports OMP_NUM_THREADS=1
    cfg = importlib.import_module("configs." + a.config)
    assert os.path.abspath(cfg.__file__).startswith(REF), cfg.__file__
    import cnn_gp
    assert os.path.abspath(cnn_gp.__file__).startswith(REF), cnn_gp.__file__
    model = cfg.initial_model
    c, s = (3, 32) if a.config == "cifar10" else (1, 28)
    gen = torch.Generator().manual_seed(1234)
    X = torch.rand(400, c, s, s, generator=gen)

    def call(e):
        t0 = time.perf_counter()
        with torch.no_grad():
            K = model(X[:e], X[400 - e:], same=False)
        dt = time.perf_counter() - t0
        assert K.shape == (e, e) and bool(torch.isfinite(K).all())
        return dt

    edge = a.edge
    if edge <= 0:
        call(16)  # first-call set-up
        per_pair = call(32) / (32 * 32)
        edge = int(max(32, min(200, (a.seconds / max(per_pair, 1e-9)) ** 0.5)))
        edge -= edge % 8
    rates = []
    for k in range(a.warmup + a.steps):
        dt = call(edge)
        if k >= a.warmup:
            rates.append(edge * edge / dt)
    print(json.dumps({
        "value": sum(rates) / len(rates), "best": max(rates), "unit": "pairs/s", "cores": cores,
        "torch_threads": torch.get_num_threads(), "kind": "reference", "edge": edge,
        "sample": f"{edge}x{edge} tile of {a.config} through the unmodified reference "
                  f"(oracle/_ref: initial_model(X, Z, same=False), no_grad, float32, torch {torch.__version__} CPU, "
                  f"{torch.get_num_threads()} threads), mean of {a.steps} after {a.warmup} warm-up"}))


if __name__ == "__main__":
    main()
