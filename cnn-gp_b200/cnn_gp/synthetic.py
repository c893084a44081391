"""Synthetic labelled images with the torchvision dataset constructor signature
``cls(root, train=True, download=False, transform=None)`` that ``DatasetFromConfig`` expects
(reference cnn_gp/data.py:149-151), for machines with no MNIST / CIFAR files and no network.

Images are class templates plus noise, ``x = 0.6 * T[y] + 0.4 * u`` with ``T`` and ``u`` uniform
on [0, 1) -- the value range ``ToTensor`` produces -- so a GP classifier separates them and
prediction margins are meaningful.
"""
import torch
from torch.utils.data import Dataset


class SyntheticImages(Dataset):
    shape = (1, 28, 28)
    n_classes = 10
    n_train, n_test = 1200, 300
    seed = 1234
    signal = 0.6

    def __init__(self, root=None, train=True, download=False, transform=None, target_transform=None):
        del root, download  # nothing is read or fetched
        g = torch.Generator().manual_seed(self.seed)
        templates = torch.rand((self.n_classes,) + tuple(self.shape), generator=g)
        n_all = self.n_train + self.n_test
        labels = torch.randint(self.n_classes, (n_all,), generator=g)
        noise_seed = torch.Generator().manual_seed(self.seed + 1)
        lo, hi = (0, self.n_train) if train else (self.n_train, n_all)
        # generate the whole stream so that train and test never share noise
        noise = torch.rand((n_all,) + tuple(self.shape), generator=noise_seed)[lo:hi]
        self.labels = labels[lo:hi].clone()
        self.images = self.signal * templates[self.labels] + (1.0 - self.signal) * noise
        self.targets = self.labels
        # images are already float tensors in ToTensor's range; ``transform`` (ToTensor, possibly
        # composed with config.transforms) is therefore only applied when it accepts tensors
        self.transform, self.target_transform = None, target_transform

    def __len__(self):
        return self.images.shape[0]

    def __getitem__(self, i):
        y = int(self.labels[i])
        if self.target_transform is not None:
            y = self.target_transform(y)
        return self.images[i], y


def synthetic_dataset(n_train, n_test, shape=(1, 28, 28), n_classes=10, seed=1234, signal=0.6):
    """A ``SyntheticImages`` subclass with the given sizes, usable as ``config.dataset``."""
    return type("SyntheticImages_%d_%d" % (n_train, n_test), (SyntheticImages,),
                dict(n_train=n_train, n_test=n_test, shape=tuple(shape), n_classes=n_classes, seed=seed,
                     signal=signal))
