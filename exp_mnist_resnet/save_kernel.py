"""Save a kernel matrix to disk -- the reference's exp_mnist_resnet/save_kernel.py:14-50 with the
same flags, dataset names (Kxx, Kxvx, Kxtx; Kv_diag, Kt_diag on worker 0) and block layout.

``--resident`` (default) keeps the datasets in HBM and streams finished block rows to the store
(cnn_gp.kernel_save_tools.save_K_resident); ``--noresident`` runs the reference's loop literally:
``save_K`` with a ``kern`` closure that copies every tile's images to the GPU and the result back
(save_kernel.py:21-24).  Both write identical files."""
import importlib
import os
import sys

import absl.app
import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [p for p in (os.path.join(_ROOT, "cnn-gp_b200"), _ROOT) if p not in sys.path]

from cnn_gp import DatasetFromConfig, save_K  # noqa: E402
from cnn_gp.block_store import open_store  # noqa: E402
from cnn_gp.kernel_save_tools import save_K_resident  # noqa: E402

FLAGS = absl.app.flags.FLAGS


def compute_all(config, dataset, out_path, batch_size=200, n_workers=1, worker_rank=0, resident=True,
                device="cuda"):
    """Everything ``main`` does after flag parsing; importable for tests and the torchrun driver."""
    model = config.initial_model.to(device)

    def kern(x, x2, same, diag):
        with torch.no_grad():
            xd = x.to(device)
            x2d = xd if x2 is x else x2.to(device)  # diagonal tiles hand over one batch twice
            return model(xd, x2d, same, diag).detach().cpu().numpy()

    kwargs = dict(worker_rank=worker_rank, n_workers=n_workers, batch_size=batch_size, print_interval=2.)

    def save(f, name, X, X2, diag):
        if resident:
            save_K_resident(f, model, name=name, X=X, X2=X2, diag=diag, device=torch.device(device), **kwargs)
        else:
            save_K(f, kern, name=name, X=X, X2=X2, diag=diag, **kwargs)

    with open_store(out_path, "w") as f:
        save(f, "Kxx", dataset.train, None, False)
        save(f, "Kxvx", dataset.validation, dataset.train, False)
        save(f, "Kxtx", dataset.test, dataset.train, False)
    if worker_rank == 0:
        with open_store(out_path, "a") as f:
            save(f, "Kv_diag", dataset.validation, None, True)
            save(f, "Kt_diag", dataset.test, None, True)


def main(_):
    print(f"CUDA_VISIBLE_DEVICES={os.environ.get('CUDA_VISIBLE_DEVICES', '(unset)')}")
    config = importlib.import_module(f"configs.{FLAGS.config}")
    dataset = DatasetFromConfig(FLAGS.datasets_path, config)
    compute_all(config, dataset, FLAGS.out_path, batch_size=FLAGS.batch_size, n_workers=FLAGS.n_workers,
                worker_rank=FLAGS.worker_rank, resident=FLAGS.resident)


if __name__ == '__main__':
    f = absl.app.flags
    f.DEFINE_string("datasets_path", "/scratch/ag919/datasets/", "where to save datasets")
    f.DEFINE_integer('batch_size', 200, "max number of examples to simultaneously compute the kernel of")
    f.DEFINE_string("config", "mnist", "which config to load from `configs`")
    f.DEFINE_integer("n_workers", 1, "num of workers")
    f.DEFINE_integer("worker_rank", 0, "rank of worker")
    f.DEFINE_string('out_path', None, "path of h5 file (or .npy store directory) to save kernels in")
    f.DEFINE_boolean("resident", True, "keep datasets in HBM and stream block rows (default) instead of "
                                       "per-tile host round trips")
    absl.app.run(main)
