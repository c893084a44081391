#!/usr/bin/env python
"""Third batch of golden fixtures from the UNMODIFIED reference: dilated windows and nested Sums
(three live maps) at the 28 x 28 and 32 x 32 sizes the fused-net kernel covers (SURVEY.md 8f rank 3;
reference cnn_gp/kernels.py:61,95-96 dilation, :246-254 Sum).

    python tests/golden/make_golden_f3.py

Writes gram_edge3_<case>.npz with the keys of make_golden.gram_case.  Nothing from this repository is
imported; the same module trees are rebuilt with our classes in tests/models.py::edge3_models.
"""
import os
import sys

import numpy as np

np.int = int  # cnn_gp/data.py:12
REF = "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import torch  # noqa: E402
import cnn_gp as ref  # noqa: E402
from make_golden import gram_case  # noqa: E402

assert os.path.realpath(ref.__file__).startswith(REF), ref.__file__


def cases():
    from cnn_gp import Conv2d, ReLU, Sequential, Sum, Mixture
    g = torch.Generator().manual_seed(777)

    def rnd(n, c, h, w, kind="rand"):
        return torch.rand(n, c, h, w, generator=g) if kind == "rand" else torch.randn(n, c, h, w, generator=g)

    out = {}
    # 3 x 3 windows with dilation 2 ("same": padding 2), alone and inside a residual branch, with a bias
    out["28_dilated"] = (Sequential(
        Conv2d(3, dilation=2, var_weight=1.3, var_bias=0.1), ReLU(), Conv2d(3), ReLU(),
        Sum([Sequential(), Sequential(Conv2d(3, dilation=2, var_bias=0.05), ReLU(), Conv2d(3, var_weight=0.8))]),
        ReLU(), Conv2d(28, padding=0, var_bias=0.02)), rnd(4, 1, 28, 28), rnd(3, 1, 28, 28))
    # a Sum inside a Sum: the outer skip, the inner skip and the working map are live at once
    out["28_nested"] = (Sequential(
        Conv2d(3),
        Sum([Sequential(), Sequential(ReLU(), Sum([Sequential(), Conv2d(3, var_weight=1.2)]), ReLU(), Conv2d(3))]),
        Mixture([Conv2d(1), Sequential(ReLU(), Conv2d(5, var_bias=0.2))], logit_proportions=torch.tensor([0.3, -0.7])),
        ReLU(), Conv2d(28, padding=0)), rnd(3, 2, 28, 28, "randn"), rnd(4, 2, 28, 28, "randn"))
    # both on the CIFAR geometry, followed by a strided stage
    out["32_dilated_nested"] = (Sequential(
        Conv2d(3, var_bias=0.3),
        Sum([Sequential(), Sequential(ReLU(), Sum([Sequential(), Conv2d(3, dilation=2)]), ReLU(), Conv2d(3))]),
        ReLU(), Conv2d(3, stride=2), ReLU(), Conv2d(16, padding=0, var_bias=0.1)), rnd(3, 3, 32, 32), rnd(3, 3, 32, 32))
    return out


if __name__ == "__main__":
    torch.set_num_threads(8)
    for name, (model, X, Z) in cases().items():
        gram_case("edge3_" + name, model, X, Z)
