"""One process per GPU: the reference's exp_mnist_resnet/run.bash:24-49 (fan out save_kernel
workers, wait, merge, classify) as a torchrun program.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \\
        -m exp_mnist_resnet.run --config=synthetic --batch_size=200 [--out_path=DIR]

Every rank holds the (small) datasets in HBM, evaluates its contiguous slice of the reference's
tile lists (cnn_gp/data.py:11-29) with no communication, and the NaN-marked partial matrices
meet on rank 0 in one NCCL reduction over NVLink (cnn_gp.tiles.gather_blocks) instead of through
per-worker files; rank 0 then runs the float64 Cholesky solve and prints the accuracies.  With
``--out_path`` rank 0 also writes the merged store in the save_K layout."""
import importlib
import os
import sys
import time

import absl.app
import torch
import torch.distributed as dist

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [p for p in (os.path.join(_ROOT, "cnn-gp_b200"), _ROOT) if p not in sys.path]

from cnn_gp import DatasetFromConfig  # noqa: E402
from cnn_gp import linalg, linalg_dist  # noqa: E402
from cnn_gp.block_store import open_store  # noqa: E402
from cnn_gp.kernel_save_tools import create_h5py_dataset  # noqa: E402
from cnn_gp.tiles import GramJob, compute_worker_blocks, gather_blocks  # noqa: E402

FLAGS = absl.app.flags.FLAGS


def gram_sharded(model, X, X2, batch_size, rank, world):
    """This rank's tiles of K(X, X2) gathered on rank 0 (None elsewhere)."""
    N, N2 = X.shape[0], (X if X2 is None else X2).shape[0]
    K = torch.full((N, N2), float("nan"), dtype=torch.float32, device=X.device)
    compute_worker_blocks(GramJob(model, X, X2), K, batch_size, rank, world, balanced=True)
    if world == 1:
        return K
    return gather_blocks(K, dst=0)


def main(_):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    config = importlib.import_module(f"configs.{FLAGS.config}")
    dataset = DatasetFromConfig(FLAGS.datasets_path, config)
    model = config.initial_model.to(dev)
    sets = {k: DatasetFromConfig.resident(getattr(dataset, k), dev) for k in ("train", "validation", "test")}
    t0 = time.perf_counter()
    with torch.no_grad():
        Kxx = gram_sharded(model, sets["train"].images, None, FLAGS.batch_size, rank, world)
        Kxvx = gram_sharded(model, sets["validation"].images, sets["train"].images, FLAGS.batch_size, rank, world)
        Kxtx = gram_sharded(model, sets["test"].images, sets["train"].images, FLAGS.batch_size, rank, world)
    torch.cuda.synchronize()
    if rank == 0:
        print(f"kernels: {time.perf_counter() - t0:.3f} s on {world} GPU(s)")
        if FLAGS.out_path:
            with open_store(FLAGS.out_path, "w") as f:
                for name, K in (("Kxx", Kxx), ("Kxvx", Kxvx), ("Kxtx", Kxtx)):
                    ds = create_h5py_dataset(f, FLAGS.batch_size, name, False, K.shape[0], K.shape[1])
                    ds[0, :, :] = K.cpu().numpy()
    Y = sets["train"].labels
    n_classes = int(Y.max()) + 1
    Y_1hot = torch.ones((len(Y), n_classes), dtype=torch.float64).neg_()
    Y_1hot[torch.arange(len(Y)), Y] = 1.
    t1 = time.perf_counter()
    if world > 1 and FLAGS.dist_solve:
        # block rows of Kxx dealt out to all GPUs, panel broadcasts over NVLink (cnn_gp.linalg_dist)
        if rank == 0 and FLAGS.jitter:
            Kxx.diagonal().add_(FLAGS.jitter)
        A = linalg_dist.solve_pos_upper_distributed(Kxx, Y_1hot.to(dev) if rank == 0 else None, len(Y), dev)
    elif rank == 0:
        K64 = Kxx.to(torch.float64)
        K64.diagonal().add_(FLAGS.jitter)
        A = linalg.solve_pos_upper(K64, Y_1hot.to(dev), overwrite_a=True)
        del K64
    if rank == 0:
        torch.cuda.synchronize()
        print(f"solve: {time.perf_counter() - t1:.3f} s (n = {len(Y)}, "
              f"{'distributed over ' + str(world) + ' GPUs' if world > 1 and FLAGS.dist_solve else 'one GPU'})")
        for key, K in (("validation", Kxvx), ("test", Kxtx)):
            pred = linalg.predict_argmax(K, A).cpu()
            acc = float((pred == sets[key].labels).double().mean())
            print(f"{key} accuracy: {acc*100}%")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    f = absl.app.flags
    f.DEFINE_string("datasets_path", "/tmp/datasets/", "where datasets live")
    f.DEFINE_integer('batch_size', 200, "tile edge of the reference's tile list")
    f.DEFINE_string("config", "synthetic", "which config to load from `configs`")
    f.DEFINE_string('out_path', None, "optional: store to write the merged kernels to")
    f.DEFINE_float("jitter", 0.0, "add to the diagonal")
    f.DEFINE_boolean("dist_solve", True, "with several GPUs: factorise Kxx across all of them")
    absl.app.run(main)
