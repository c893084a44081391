"""CIFAR-10 ResNet GP (reference configs/cifar10.py): 3x32x32 inputs, stages at 32/16/8 pixels,
an 8x8 valid convolution as pooling, then 1x1 conv, ReLU, 1x1 conv."""
from cnn_gp import Conv2d, ReLU
from ._common import dataset_class, resnet_gp

train_range = range(40000)
validation_range = range(40000, 50000)
test_range = range(50000, 60000)

kernel_batch_size = 350

dataset_name = "CIFAR10"
model_name = "ResNet"
in_channels = 3
transforms = []
epochs = 0
initial_model = resnet_gp(final_pool=8, tail=(
    Conv2d(kernel_size=1, padding=0, in_channel_multiplier=4, out_channel_multiplier=4),
    ReLU(),
    Conv2d(kernel_size=1, padding=0, in_channel_multiplier=4)))


def __getattr__(name):
    if name == "dataset":
        return dataset_class(dataset_name)
    raise AttributeError(name)
