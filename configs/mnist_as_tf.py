"""MNIST ResNet GP with the train / validation / test split of the paper's TensorFlow
experiments (reference configs/mnist_as_tf.py:9-11); same architecture as ``mnist``."""
from cnn_gp import Conv2d, ReLU
from ._common import dataset_class, resnet_gp

train_range = range(5000, 55000)
validation_range = list(range(55000, 60000)) + list(range(0, 5000))
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
