#!/usr/bin/env python
"""Golden fixtures for the two documented deviations of the fused float32 kernels, from the
UNMODIFIED reference (run in the authoring container only: python tests/golden/make_golden_degenerate.py).

  degenerate_<model>.npz   X, Z with near-collinear / duplicate / all-zero images and the reference's
                           Kxz (same=False) in float32 AND float64, plus Kxx in both

(a) near-collinear pairs in a same=False tile: the reference's float32 formula cancels in
    `xx*yy - xy**2` (cnn_gp/kernels.py:150) and drifts from its own float64 answer; the fused kernels
    use a cancellation-free form.  (b) an all-zero image with zero bias: entries are regularisation
    artefacts of `+ f32_tiny` (kernels.py:133,146), ~1e-20-sized.  tests/test_gpu_gram.py holds the
    CUDA paths to these numbers with explicit bounds.
"""
import importlib
import os
import sys

import numpy as np

np.int = int  # cnn_gp/data.py:12
REF = "/root/reference"
sys.path.insert(0, REF)

import torch  # noqa: E402
import cnn_gp as ref  # noqa: E402

assert os.path.realpath(ref.__file__).startswith(REF), ref.__file__
HERE = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(8)


def inputs(C, S, seed):
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(8, C, S, S, generator=g)
    X[1] = 0.0                     # all-zero image
    X[3] = X[2] * (1 + 1e-4)       # almost collinear: cos(theta) -> 1
    X[4] = X[2]                    # exact duplicate in another slot
    X[5] = X[2] + 1e-3 * torch.rand(C, S, S, generator=g)   # near duplicate, not collinear
    X[6] = 0.5 * X[2]              # exactly collinear, different norm
    Z = X.flip(0).contiguous()
    return X, Z


def main():
    readme = ref.Sequential(ref.Conv2d(kernel_size=3), ref.ReLU(), ref.Conv2d(kernel_size=3, stride=2), ref.ReLU(),
                            ref.Conv2d(kernel_size=14, padding=0))
    cases = {"readme": (readme, 3, 28)}
    for name, C, S in (("mnist_paper_convnet_gp", 1, 28), ("mnist_as_tf", 1, 28), ("mnist_paper_residual_cnn_gp", 1, 28)):
        cases[name] = (importlib.import_module("configs." + name).initial_model, C, S)
    for k, (name, (model, C, S)) in enumerate(cases.items()):
        X, Z = inputs(C, S, 40 + k)
        out = {"X": X.numpy(), "Z": Z.numpy()}
        for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
            m = model.to(dt)
            with torch.no_grad():
                out[f"Kxz_{tag}"] = m(X.to(dt), Z.to(dt)).numpy()
                out[f"Kxx_{tag}"] = m(X.to(dt)).numpy()
        model.to(torch.float32)
        np.savez_compressed(os.path.join(HERE, f"degenerate_{name}.npz"), **out)
        d = np.abs(out["Kxz_f32"] - out["Kxz_f64"]) / np.abs(out["Kxz_f64"])
        print(name, "reference f32 vs its own f64, same=False tile: max rel", d.max(), "at", np.unravel_index(d.argmax(), d.shape),
              "| zero-image entry", out["Kxz_f32"][1, 0], out["Kxz_f64"][1, 0])


if __name__ == "__main__":
    main()
