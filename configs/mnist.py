"""MNIST ResNet GP, default split (reference configs/mnist.py)."""
from cnn_gp import Conv2d, ReLU
from ._common import dataset_class, resnet_gp

train_range = range(50000)
validation_range = range(50000, 60000)
test_range = range(60000, 70000)

dataset_name = "MNIST"
model_name = "ResNet"
transforms = []
epochs = 0
in_channels = 1
out_channels = 10
initial_model = resnet_gp(final_pool=7, tail=(ReLU(), Conv2d(kernel_size=1, padding=0, in_channel_multiplier=4)))


def __getattr__(name):
    if name == "dataset":
        return dataset_class(dataset_name)
    raise AttributeError(name)
