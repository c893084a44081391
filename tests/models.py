"""The models behind tests/golden/gram_*.npz, rebuilt with this repository's cnn_gp package
(tests/golden/make_golden.py builds the same trees with the reference's classes)."""
import importlib

import torch

from cnn_gp import Conv2d, ReLU, Sequential, Sum, Mixture, resnet_block

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


def edge2_models():
    """The second batch (tests/golden/make_golden_more.py)."""
    return {
        "edge2_relu_first": Sequential(ReLU(), Conv2d(3, var_weight=2.0, var_bias=0.1), ReLU(), Conv2d(12, padding=0)),
        "edge2_mixture3": Sequential(
            Conv2d(3, var_bias=0.2), ReLU(),
            Sum([Sequential(), Mixture([Sequential(), Conv2d(3), Sequential(ReLU(), Conv2d(5, var_weight=0.7))],
                                       logit_proportions=torch.tensor([0.5, -1.0, 0.25]))]),
            ReLU(), Conv2d(12, padding=0, var_bias=0.01)),
        "edge2_resnet_blocks": Sequential(
            Conv2d(3, var_weight=1.3, var_bias=0.05), resnet_block(stride=1),
            resnet_block(stride=2, projection_shortcut=True), resnet_block(stride=1, projection_shortcut=True),
            ReLU(), Conv2d(6, padding=0)),
        "edge2_28_windows": Sequential(
            Conv2d(5, var_weight=1.9, var_bias=3.0), ReLU(), Conv2d(7, var_weight=0.8, var_bias=0.0), ReLU(),
            Conv2d(1, var_weight=0.4, var_bias=0.7), Conv2d(4, var_weight=2.2, var_bias=1e-3), ReLU(),
            Conv2d(3, var_weight=30.0, var_bias=5.0), ReLU(), Conv2d(28, padding=0, var_weight=0.9, var_bias=0.02)),
        "edge2_28_sum_stride": Sequential(
            Conv2d(7, var_bias=0.1), ReLU(),
            Sum([Sequential(), Sequential(Conv2d(5, var_weight=1.5), ReLU(), Conv2d(3))]),
            resnet_block(stride=2, projection_shortcut=True), ReLU(), Conv2d(14, padding=0, var_bias=0.3)),
        "edge2_single": Sequential(Conv2d(3), ReLU(), Conv2d(3, stride=2), ReLU(), Conv2d(14, padding=0)),
        "edge2_32_plain": Sequential(Conv2d(3, var_weight=2.0, var_bias=0.5), ReLU(), Conv2d(5, var_weight=1.2), ReLU(),
                                     Conv2d(3, stride=2), ReLU(), Conv2d(16, padding=0)),
        "edge2_dil_stride": Sequential(
            Conv2d(4, dilation=2, var_bias=0.2), ReLU(), Conv2d(3, stride=2, dilation=2, var_weight=1.4), ReLU(),
            Conv2d(2, dilation=3, padding=0), ReLU(), Conv2d(3, padding=0)),
    }


def edge3_models():
    """The third batch (tests/golden/make_golden_f3.py): dilated windows and nested Sums at fused-net sizes."""
    return {
        "edge3_28_dilated": Sequential(
            Conv2d(3, dilation=2, var_weight=1.3, var_bias=0.1), ReLU(), Conv2d(3), ReLU(),
            Sum([Sequential(), Sequential(Conv2d(3, dilation=2, var_bias=0.05), ReLU(), Conv2d(3, var_weight=0.8))]),
            ReLU(), Conv2d(28, padding=0, var_bias=0.02)),
        "edge3_28_nested": Sequential(
            Conv2d(3),
            Sum([Sequential(), Sequential(ReLU(), Sum([Sequential(), Conv2d(3, var_weight=1.2)]), ReLU(), Conv2d(3))]),
            Mixture([Conv2d(1), Sequential(ReLU(), Conv2d(5, var_bias=0.2))], logit_proportions=torch.tensor([0.3, -0.7])),
            ReLU(), Conv2d(28, padding=0)),
        "edge3_32_dilated_nested": Sequential(
            Conv2d(3, var_bias=0.3),
            Sum([Sequential(), Sequential(ReLU(), Sum([Sequential(), Conv2d(3, dilation=2)]), ReLU(), Conv2d(3))]),
            ReLU(), Conv2d(3, stride=2), ReLU(), Conv2d(16, padding=0, var_bias=0.1)),
    }


def golden_models():
    """name -> model for every gram_<name>.npz fixture."""
    out = {"readme": readme_model()}
    for c in CONFIGS:
        m = importlib.import_module("configs." + c).initial_model
        out[c] = m
        out[c + "_randn"] = m
    out.update(edge_models())
    out.update(edge2_models())
    out.update(edge3_models())
    return out
