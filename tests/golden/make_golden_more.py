#!/usr/bin/env python
"""More golden fixtures from the UNMODIFIED reference (second batch; tests/golden/make_golden.py
wrote the first and is left untouched so that its random streams and outputs stay as committed).

    python tests/golden/make_golden_more.py

Writes gram_edge2_<case>.npz with the same keys as make_golden.gram_case (all six call forms of
cnn_gp/kernels.py:18-57 in f32 and f64).  Nothing from this repository is imported.  The same
module trees are rebuilt with our classes in tests/models.py::edge2_models.
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
    from cnn_gp import Conv2d, ReLU, Sequential, Sum, Mixture, resnet_block
    g = torch.Generator().manual_seed(4321)

    def rnd(n, c, h, w, kind="rand"):
        return torch.rand(n, c, h, w, generator=g) if kind == "rand" else torch.randn(n, c, h, w, generator=g)

    out = {}
    # the program starts with a ReLU on the raw input covariance (kernels.py:43-49 feed it directly)
    out["relu_first"] = (Sequential(ReLU(), Conv2d(3, var_weight=2.0, var_bias=0.1), ReLU(), Conv2d(12, padding=0)),
                         rnd(4, 2, 12, 12), rnd(3, 2, 12, 12))
    # three-way Mixture with an identity branch, inside a Sum (kernels.py:203-225, 246-254)
    out["mixture3"] = (Sequential(
        Conv2d(3, var_bias=0.2), ReLU(),
        Sum([Sequential(), Mixture([Sequential(), Conv2d(3), Sequential(ReLU(), Conv2d(5, var_weight=0.7))],
                                   logit_proportions=torch.tensor([0.5, -1.0, 0.25]))]),
        ReLU(), Conv2d(12, padding=0, var_bias=0.01)), rnd(3, 1, 12, 12, "randn"), rnd(4, 1, 12, 12, "randn"))
    # the resnet_block factory on small maps: identity block, strided projection block (kernels.py:274-296)
    out["resnet_blocks"] = (Sequential(
        Conv2d(3, var_weight=1.3, var_bias=0.05), resnet_block(stride=1), resnet_block(stride=2, projection_shortcut=True),
        resnet_block(stride=1, projection_shortcut=True), ReLU(), Conv2d(6, padding=0)),
        rnd(4, 3, 12, 12), rnd(2, 3, 12, 12))
    # 28 x 28 straight-line program with every window the fused kernel knows, large biases, k = 1
    out["28_windows"] = (Sequential(
        Conv2d(5, var_weight=1.9, var_bias=3.0), ReLU(), Conv2d(7, var_weight=0.8, var_bias=0.0), ReLU(),
        Conv2d(1, var_weight=0.4, var_bias=0.7), Conv2d(4, var_weight=2.2, var_bias=1e-3), ReLU(),
        Conv2d(3, var_weight=30.0, var_bias=5.0), ReLU(), Conv2d(28, padding=0, var_weight=0.9, var_bias=0.02)),
        rnd(5, 1, 28, 28), rnd(4, 1, 28, 28))
    # 28 x 28 with Sum, stride 2 and 5 x 5 / 7 x 7 windows
    out["28_sum_stride"] = (Sequential(
        Conv2d(7, var_bias=0.1), ReLU(),
        Sum([Sequential(), Sequential(Conv2d(5, var_weight=1.5), ReLU(), Conv2d(3))]),
        resnet_block(stride=2, projection_shortcut=True), ReLU(), Conv2d(14, padding=0, var_bias=0.3)),
        rnd(4, 2, 28, 28, "randn"), rnd(3, 2, 28, 28, "randn"))
    # a single image on each side
    out["single"] = (Sequential(Conv2d(3), ReLU(), Conv2d(3, stride=2), ReLU(), Conv2d(14, padding=0)),
                     rnd(1, 3, 28, 28), rnd(1, 3, 28, 28))
    # 32 x 32 x 3 without any Sum (not a ResNet): straight-line on the CIFAR geometry
    out["32_plain"] = (Sequential(Conv2d(3, var_weight=2.0, var_bias=0.5), ReLU(), Conv2d(5, var_weight=1.2), ReLU(),
                                  Conv2d(3, stride=2), ReLU(), Conv2d(16, padding=0)),
                       rnd(3, 3, 32, 32), rnd(4, 3, 32, 32))
    # dilation together with stride and with an even "same" kernel (zero first row / column, kernels.py:71-84)
    out["dil_stride"] = (Sequential(
        Conv2d(4, dilation=2, var_bias=0.2), ReLU(), Conv2d(3, stride=2, dilation=2, var_weight=1.4), ReLU(),
        Conv2d(2, dilation=3, padding=0), ReLU(), Conv2d(3, padding=0)), rnd(3, 2, 12, 12), rnd(3, 2, 12, 12))
    return out


if __name__ == "__main__":
    torch.set_num_threads(8)
    for name, (model, X, Z) in cases().items():
        gram_case("edge2_" + name, model, X, Z)
