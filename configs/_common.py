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
