"""Shared pieces of the shipped configurations."""
from cnn_gp import Conv2d, ReLU, Sequential, resnet_block


def dataset_class(name):
    """torchvision dataset class by name, resolved lazily so that importing a config does not
    need torchvision unless the dataset is actually used."""
    import torchvision
    return getattr(torchvision.datasets, name)


def resnet_gp(final_pool, stage_blocks=5, tail=()):
    """The 32-layer ResNet GP of the reference (configs/mnist.py:15-45, configs/cifar10.py:16-47):
    a 3x3 stem, three stages of ``stage_blocks`` pre-activation blocks (the first of each stage
    has a projection shortcut; stages two and three halve the map), then a ``final_pool``-sized
    valid convolution standing in for average pooling, and ``tail``."""
    mods = [Conv2d(kernel_size=3)]
    for stage, (stride, mult) in enumerate(((1, 1), (2, 2), (2, 4))):
        mods.append(resnet_block(stride=stride, projection_shortcut=True, multiplier=mult))
        mods += [resnet_block(stride=1, projection_shortcut=False, multiplier=mult)
                 for _ in range(stage_blocks - 1)]
    mods.append(Conv2d(kernel_size=final_pool, padding=0, in_channel_multiplier=4,
                       out_channel_multiplier=4))
    mods += list(tail)
    return Sequential(*mods)


def tf_mnist_split():
    """(train, validation, test) index ranges of the paper's TensorFlow-style MNIST split: the
    middle 50 000 training images, the 10 000 around them for validation, the official test set."""
    lo, hi, n_train, n_all = 5000, 55000, 60000, 70000
    return range(lo, hi), [*range(hi, n_train), *range(lo)], range(n_train, n_all)


def lazy_dataset(module_globals):
    """Module-level ``__getattr__`` that resolves ``dataset`` to the torchvision class named by the
    module's ``dataset_name`` on first use (importing a config must not need torchvision)."""
    def __getattr__(name):
        if name == "dataset":
            return dataset_class(module_globals["dataset_name"])
        raise AttributeError(name)
    return __getattr__


def stacked(n, make_layer, *tail):
    """Sequential of ``n`` repetitions of ``make_layer()`` (a list of modules) followed by ``tail``."""
    mods = []
    for _ in range(n):
        mods.extend(make_layer())
    return Sequential(*mods, *tail)
