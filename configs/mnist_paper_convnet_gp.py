"""The 7-layer ConvNet GP of the paper (reference configs/mnist_paper_convnet_gp.py:16-30):
seven [7x7 "same" conv, ReLU] layers and a 28x28 valid convolution as the dense layer, with the
randomly-searched variances var_weight = 2.79 (per tap sum, hence the k^2 factor) and
var_bias = 7.86.  This is the program BASELINE.json's headline metric is quoted on."""
from cnn_gp import Conv2d, ReLU
from ._common import lazy_dataset, stacked, tf_mnist_split

train_range, validation_range, test_range = tf_mnist_split()
dataset_name, model_name = "MNIST", "ResNet"
in_channels, out_channels, epochs, transforms = 1, 10, 0, []

var_weight, var_bias, n_layers, window = 2.79, 7.86, 7, 7


def _hidden():
    return [Conv2d(window, padding="same", var_weight=var_weight * window ** 2, var_bias=var_bias), ReLU()]


initial_model = stacked(n_layers, _hidden, Conv2d(28, padding=0, var_weight=var_weight, var_bias=var_bias))
__getattr__ = lazy_dataset(globals())
