"""One process per GPU: the reference's exp_mnist_resnet/run.bash:24-49 (fan out save_kernel
workers, wait, merge, classify) as a torchrun program.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \\
        -m exp_mnist_resnet.run --config=synthetic --batch_size=200 [--out_path=DIR]

Every rank holds the (small) datasets in HBM, evaluates its contiguous slice of the reference's
tile lists (cnn_gp/data.py:11-29) with no communication and keeps only the block rows its tiles
touch (cnn_gp.tiles.RowShard).  Nothing is gathered to one GPU: finished block rows of Kxx travel
from their owners straight into the block-cyclic rows of the distributed float64 Cholesky (NCCL
broadcasts of row panels over NVLink), both triangular sweeps run on the distributed factor, and
every rank scores the validation / test entries it computed itself (partial K* A summed by one
small all-reduce).  With ``--out_path`` rank 0 additionally streams the merged matrices, row panel by
row panel, into a store in the save_K layout -- the job of merge_h5_files.py."""
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
from cnn_gp.tiles import GramJob, RowShard, exchange_rows  # noqa: E402

FLAGS = absl.app.flags.FLAGS


def gram_shard(model, X, X2, batch_size, rank, world):
    """This rank's tiles of K(X, X2), in the block rows they touch."""
    N, N2 = X.shape[0], (X if X2 is None else X2).shape[0]
    shard = RowShard(N, N2, batch_size, rank, world, X2 is None, X.device)
    shard.compute(GramJob(model, X, X2))
    return shard


def partial_scores(shard, A):
    """sum over the entries this rank computed of K*[i, j] A[j, :]  ->  [N*, classes] (classify_gp.py:40)."""
    scores = torch.zeros((shard.N, A.shape[1]), dtype=torch.float64, device=A.device)
    for r, has_diag, c0, c1 in shard.segments:
        i0, i1 = r * shard.bs, min(shard.N, (r + 1) * shard.bs)
        j0, j1 = c0 * shard.bs, min(shard.N2, c1 * shard.bs)
        K = shard.data[i0 - shard.row_lo:i1 - shard.row_lo, j0:j1]
        scores[i0:i1] += linalg.predict_argmax(K, A[j0:j1].contiguous(), return_scores=True)[1]
    return scores


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
        Kxx = gram_shard(model, sets["train"].images, None, FLAGS.batch_size, rank, world)
        Kxvx = gram_shard(model, sets["validation"].images, sets["train"].images, FLAGS.batch_size, rank, world)
        Kxtx = gram_shard(model, sets["test"].images, sets["train"].images, FLAGS.batch_size, rank, world)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    n = Kxx.N
    if rank == 0:
        print(f"kernels: {time.perf_counter() - t0:.3f} s on {world} GPU(s); rank 0 holds rows "
              f"[{Kxx.row_lo}, {Kxx.row_hi}) of Kxx ({Kxx.data.numel() * 4 / 2**30:.2f} GiB of {n * n * 4 / 2**30:.2f})")
    if FLAGS.out_path:
        # the merge: row panels from their owners to rank 0, which writes them as they arrive
        t2 = time.perf_counter()
        store = open_store(FLAGS.out_path, "w") if rank == 0 else None
        for name, sh in (("Kxx", Kxx), ("Kxvx", Kxvx), ("Kxtx", Kxtx)):
            ds = create_h5py_dataset(store, FLAGS.batch_size, name, False, sh.N, sh.N2) if rank == 0 else None

            def write(i0, i1, panel, ds=ds):
                ds[0, i0:i1, :] = panel.cpu().numpy()
            exchange_rows(sh, write, wanted=lambda i0, i1: rank == 0)
        if rank == 0:
            store.close()
            print(f"store written: {time.perf_counter() - t2:.3f} s")
    Y = sets["train"].labels
    n_classes = int(Y.max()) + 1
    Y_1hot = torch.ones((len(Y), n_classes), dtype=torch.float64).neg_()
    Y_1hot[torch.arange(len(Y)), Y] = 1.
    t1 = time.perf_counter()
    if world > 1 and FLAGS.dist_solve:
        # block rows of Kxx dealt out to all GPUs, panel broadcasts over NVLink; the jitter is added in
        # float64 on the owners (classify_gp.py:64), the sweeps run on the distributed factor
        A = linalg_dist.solve_pos_upper_distributed(
            None, Y_1hot.to(dev) if rank == 0 else None, n, dev, jitter=FLAGS.jitter,
            fill=lambda ch: exchange_rows(Kxx, ch.fill_rows, wanted=ch.wants_rows))
    else:
        K64 = torch.empty((n, n), dtype=torch.float64, device=dev) if rank == 0 else None
        exchange_rows(Kxx, lambda i0, i1, panel: K64[i0:i1].copy_(panel), wanted=lambda i0, i1: rank == 0)
        A = torch.empty((n, n_classes), dtype=torch.float64, device=dev)
        if rank == 0:
            K64.diagonal().add_(FLAGS.jitter)
            A.copy_(linalg.solve_pos_upper(K64, Y_1hot.to(dev), overwrite_a=True))
            del K64
        if world > 1:
            dist.broadcast(A, src=0)
    torch.cuda.synchronize()
    if rank == 0:
        print(f"solve: {time.perf_counter() - t1:.3f} s (n = {n}, "
              f"{'distributed over ' + str(world) + ' GPUs' if world > 1 and FLAGS.dist_solve else 'one GPU'}); "
              f"peak HBM on rank 0: {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")
    for key, sh in (("validation", Kxvx), ("test", Kxtx)):
        scores = partial_scores(sh, A)
        if world > 1:
            dist.all_reduce(scores)
        if rank == 0:
            pred = scores.argmax(dim=1).cpu()
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
