"""The 7-layer ConvNet GP of the paper (reference configs/mnist_paper_convnet_gp.py:16-30):
seven [7x7 "same" conv, ReLU] layers and a 28x28 valid convolution as the dense layer, with the
randomly-searched variances var_weight = 2.79 (per tap sum, hence the k^2 factor) and
var_bias = 7.86.  This is the program BASELINE.json's headline metric is quoted on."""
from cnn_gp import Conv2d, ReLU, Sequential
from ._common import dataset_class

train_range = range(5000, 55000)
validation_range = list(range(55000, 60000)) + list(range(0, 5000))
test_range = range(60000, 70000)

dataset_name = "MNIST"
model_name = "ResNet"
transforms = []
epochs = 0
in_channels = 1
out_channels = 10

var_bias = 7.86
var_weight = 2.79
n_layers = 7


def _hidden_layer():
    return [Conv2d(kernel_size=7, padding="same", var_weight=var_weight * 7 ** 2, var_bias=var_bias), ReLU()]


initial_model = Sequential(
    *[m for _ in range(n_layers) for m in _hidden_layer()],
    Conv2d(kernel_size=28, padding=0, var_weight=var_weight, var_bias=var_bias))


def __getattr__(name):
    if name == "dataset":
        return dataset_class(dataset_name)
    raise AttributeError(name)
