"""A machine-checkable stand-in for the MNIST configs when no dataset files exist: the paper's
7-layer ConvNet GP (configs/mnist_paper_convnet_gp.py) on synthetic 28x28x1 class-template
images (cnn_gp/synthetic.py).  Sizes follow the environment variables CNNGP_SYNTH_TRAIN /
CNNGP_SYNTH_VAL / CNNGP_SYNTH_TEST (defaults 1000 / 200 / 300) so that the same
save_kernel -> merge -> classify_gp pipeline runs from a unit test up to 60k images."""
import importlib
import os

from cnn_gp.synthetic import synthetic_dataset

# CNNGP_SYNTH_MODEL names the config whose architecture is used (default: the 7-layer ConvNet GP)
initial_model = importlib.import_module(
    "configs." + os.environ.get("CNNGP_SYNTH_MODEL", "mnist_paper_convnet_gp")).initial_model

_n_train = int(os.environ.get("CNNGP_SYNTH_TRAIN", "1000"))
_n_val = int(os.environ.get("CNNGP_SYNTH_VAL", "200"))
_n_test = int(os.environ.get("CNNGP_SYNTH_TEST", "300"))

train_range = range(0, _n_train)
validation_range = range(_n_train, _n_train + _n_val)
test_range = range(_n_train + _n_val, _n_train + _n_val + _n_test)

dataset_name = "SYNTHETIC"
model_name = "ConvNet"
transforms = []
epochs = 0
# CNNGP_SYNTH_SHAPE = "C,H,W" of the images (default 1,28,28; "3,32,32" for the CIFAR-10 architecture)
_shape = tuple(int(v) for v in os.environ.get("CNNGP_SYNTH_SHAPE", "1,28,28").split(","))
in_channels = _shape[0]
out_channels = 10
dataset = synthetic_dataset(_n_train + _n_val, _n_test, shape=_shape)
