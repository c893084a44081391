"""End-to-end timing of exp_mnist_resnet.save_kernel (resident path, store on local disk) +
classify_gp on a synthetic dataset.
usage: python scripts/bench_pipeline.py N_TRAIN [MODEL_CONFIG] [h5|npy]   (store kind; default h5 = native HDF5)"""
import importlib
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "cnn-gp_b200"), ROOT]
n = int(sys.argv[1])
os.environ.update(CNNGP_SYNTH_TRAIN=str(n), CNNGP_SYNTH_VAL=str(n // 10), CNNGP_SYNTH_TEST=str(n // 5))
if len(sys.argv) > 2 and sys.argv[2] not in ("h5", "npy"):
    os.environ["CNNGP_SYNTH_MODEL"] = sys.argv[2]
    if sys.argv[2] == "cifar10":
        os.environ["CNNGP_SYNTH_SHAPE"] = "3,32,32"
store = "npy" if sys.argv[-1] == "npy" else "h5"
import torch  # noqa: E402
from cnn_gp import DatasetFromConfig  # noqa: E402
from exp_mnist_resnet import classify_gp, save_kernel  # noqa: E402

cfg = importlib.import_module("configs.synthetic")
ds = DatasetFromConfig("/nonexistent", cfg)
tmp = tempfile.mkdtemp(prefix="cnngp_store_")
try:
    path = os.path.join(tmp, "k.h5" if store == "h5" else "k")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    save_kernel.compute_all(cfg, ds, path, batch_size=200, resident=True)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    res = classify_gp.classify(cfg, ds, path)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    pairs = n * (n + 1) // 2 + (n // 10 + n // 5) * n
    size = sum(os.path.getsize(os.path.join(d, f)) for d, _, fs in os.walk(tmp) for f in fs)
    print(json.dumps({"n_train": n, "store": store, "store_bytes": size, "model": os.environ.get("CNNGP_SYNTH_MODEL", "mnist_paper_convnet_gp"),
                      "save_kernel_s": t1 - t0, "pairs": pairs, "save_kernel_Mpairs_per_s": pairs / (t1 - t0) / 1e6,
                      "classify_s": t2 - t1, "val_acc": res["validation"], "test_acc": res["test"]}))
finally:
    shutil.rmtree(tmp, ignore_errors=True)
