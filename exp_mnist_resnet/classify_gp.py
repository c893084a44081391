"""Given a pre-computed kernel and a data set, compute validation / test accuracy -- the
reference's exp_mnist_resnet/classify_gp.py:17-102 with the dense float64 work on the GPU:

    solve_system   scipy.linalg.solve(assume_a='pos', lower=False) (classify_gp.py:24-26)
                   -> blocked Cholesky on the FP64 tensor pipe + triangular solves (cnn_gp.linalg)
    print_accuracy (Kxvx @ A).argmax(1) (classify_gp.py:40) -> cnngp_predict_argmax

Function names, arguments and flags are the reference's."""
import importlib
import os
import sys

import absl.app
import numpy as np
import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [p for p in (os.path.join(_ROOT, "cnn-gp_b200"), _ROOT) if p not in sys.path]

from cnn_gp import DatasetFromConfig  # noqa: E402
from cnn_gp import linalg  # noqa: E402
from cnn_gp.block_store import open_store  # noqa: E402

FLAGS = absl.app.flags.FLAGS
DEVICE = "cuda"


def solve_system(Kxx, Y):
    """A = Kxx^{-1} Y from the upper triangle of ``Kxx`` (float64, overwritten like the
    reference's ``overwrite_a=True``).  CPU tensors are moved to the GPU and the result comes
    back on the input's device."""
    print("Running B200 Cholesky solve Kxx^-1 Y routine")
    assert Kxx.dtype == torch.float64 and Y.dtype == torch.float64, """
    It is important that `Kxx` and `Y` are `float64`s for the inversion,
    even if they were `float32` when being calculated. This makes the
    inversion much less likely to complain about the matrix being singular.
    """
    where = Y.device
    A = linalg.solve_pos_upper(Kxx.to(DEVICE), Y.to(DEVICE), overwrite_a=True)
    return A.to(where)


def diag_add(K, diag):
    if isinstance(K, torch.Tensor):
        K.view(K.numel())[::K.shape[-1] + 1] += diag
    elif isinstance(K, np.ndarray):
        K.flat[::K.shape[-1] + 1] += diag
    else:
        raise TypeError("What do I do with a `{}`, K={}?".format(type(K), K))


def predict(A, Kxvx):
    """argmax_c (Kxvx @ A)[:, c] accumulated in float64 (classify_gp.py:40); ``Kxvx`` float32 (as stored)
    or float64 -- both go through the library's prediction kernel, never through a library matmul."""
    K = Kxvx.to(DEVICE)
    if K.dtype not in (torch.float32, torch.float64):
        K = K.to(torch.float64)
    return linalg.predict_argmax(K.contiguous(), A.to(DEVICE)).cpu()


def print_accuracy(A, Kxvx, Y, key):
    Ypred = predict(A, Kxvx)
    acc = float((torch.as_tensor(Y).cpu() == Ypred).double().mean())
    print(f"{key} accuracy: {acc*100}%")
    return acc


def load_kern(dset, i, dtype=torch.float64, device=DEVICE, block_bytes=256 << 20):
    """Slab ``i`` of a stored kernel as a ``dtype`` tensor on ``device`` (reference: float64 on the
    CPU, classify_gp.py:45-48).  Row blocks go file -> pinned buffer -> HBM through two buffers, so
    reading the next block overlaps the upload of the previous one, and the widening happens in
    HBM; the host never holds more than two blocks."""
    n, m = dset.shape[1:]
    out = torch.empty((n, m), dtype=dtype, device=device)
    rows = max(1, min(n, block_bytes // max(1, 4 * m)))
    bufs = [torch.empty((rows, m), dtype=torch.float32).pin_memory() for _ in range(min(2, -(-n // rows)))]
    done = [None] * len(bufs)
    side = torch.cuda.Stream(device)
    side.wait_stream(torch.cuda.current_stream(device))
    for k, r0 in enumerate(range(0, n, rows)):
        r1 = min(n, r0 + rows)
        b = bufs[k % len(bufs)]
        if done[k % len(bufs)] is not None:
            done[k % len(bufs)].synchronize()  # the upload that last used this buffer
        dset.read_direct(b.numpy()[:r1 - r0], source_sel=np.s_[i, r0:r1, :])
        with torch.cuda.stream(side):
            out[r0:r1].copy_(b[:r1 - r0].to(device, non_blocking=True))
            done[k % len(bufs)] = torch.cuda.Event()
            done[k % len(bufs)].record(side)
    torch.cuda.current_stream(device).wait_stream(side)
    return out


def classify(config, dataset, in_path, jitter=0.0):
    print("Reading training labels")
    _, Y = dataset.load_full(dataset.train)
    n_classes = int(Y.max()) + 1
    Y_1hot = torch.ones((len(Y), n_classes), dtype=torch.float64).neg_()  # all -1
    Y_1hot[torch.arange(len(Y)), Y] = 1.
    out = {}
    with open_store(in_path, "r") as f:
        print("Loading kernel")
        Kxx = load_kern(f["Kxx"], 0)
        diag_add(Kxx, jitter)
        print("Solving Kxx^{-1} Y")
        A = solve_system(Kxx, Y_1hot.to(DEVICE))
        del Kxx
        for key, name, subset in (("validation", "Kxvx", dataset.validation), ("test", "Kxtx", dataset.test)):
            _, Ys = dataset.load_full(subset)
            K = load_kern(f[name], 0, dtype=torch.float32)
            out[key] = print_accuracy(A, K, Ys, key)
            del K
    out["A"] = A
    return out


def main(_):
    config = importlib.import_module(f"configs.{FLAGS.config}")
    dataset = DatasetFromConfig(FLAGS.datasets_path, config)
    classify(config, dataset, FLAGS.in_path, FLAGS.jitter)


if __name__ == '__main__':
    f = absl.app.flags
    f.DEFINE_string("datasets_path", "/scratch/ag919/datasets/", "where to save datasets")
    f.DEFINE_string("config", "mnist", "which config to load from `configs`")
    f.DEFINE_string('in_path', "/scratch/ag919/grams_pytorch/mnist/dest.h5", "path of h5 file to load kernels from")
    f.DEFINE_float("jitter", 0.0, "add to the diagonal")
    absl.app.run(main)
