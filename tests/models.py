"""The models behind tests/golden/gram_*.npz, rebuilt with this repository's cnn_gp package
(tests/golden/make_golden.py builds the same trees with the reference's classes)."""
import importlib

import torch

from cnn_gp import Conv2d, ReLU, Sequential, Sum, Mixture

CONFIGS = ("mnist_paper_convnet_gp", "mnist_paper_residual_cnn_gp", "mnist_as_tf", "mnist", "cifar10")


def readme_model():
    return Sequential(Conv2d(kernel_size=3), ReLU(), Conv2d(kernel_size=3, stride=2), ReLU(),
                      Conv2d(kernel_size=14, padding=0))


def edge_models():
    return {
        "edge_evenk_sum": Sequential(
            Conv2d(4, var_weight=1.7, var_bias=0.3), ReLU(),
            Sum([Conv2d(1, var_weight=0.5), Sequential(Conv2d(3), ReLU(), Conv2d(2))]),
            ReLU(), Conv2d(12, padding=0)),
        "edge_dilated": Sequential(
            Conv2d(3, dilation=2), ReLU(), Conv2d(3, stride=3, padding=1, var_bias=0.1),
            ReLU(), Conv2d(2, padding=0, dilation=3)),
        "edge_nested": Sequential(
            Conv2d(3),
            Sum([Sequential(), Sequential(ReLU(), Sum([Sequential(), Conv2d(3)]), ReLU(), Conv2d(3))]),
            Mixture([Conv2d(1), Sequential(ReLU(), Conv2d(5, var_bias=0.2))],
                    logit_proportions=torch.tensor([0.3, -0.7])),
            ReLU(), Conv2d(12, padding=0)),
        "edge_linear": Sequential(Conv2d(3, var_bias=0.5), Conv2d(12, padding=0)),
        "edge_nonsquare": Sequential(Conv2d(3), ReLU(), Conv2d(10, stride=4, padding=0, var_bias=0.05), ReLU(),
                                     Conv2d(3, stride=3, padding=1)),
    }


def golden_models():
    """name -> model for every gram_<name>.npz fixture."""
    out = {"readme": readme_model()}
    for c in CONFIGS:
        m = importlib.import_module("configs." + c).initial_model
        out[c] = m
        out[c + "_randn"] = m
    out.update(edge_models())
    return out
