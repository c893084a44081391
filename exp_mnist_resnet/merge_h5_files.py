"""Merge the workers' stores: copy ``src`` into ``dest`` wherever ``dest`` is NaN, for every
dataset present in both -- the reference's exp_mnist_resnet/merge_h5_files.py:13-30 (NaN is the
"block not computed by this worker" marker of cnn_gp/kernel_save_tools.py:21-23).

usage: python -m exp_mnist_resnet.merge_h5_files dest_file [source_file1 source_file2 ...]"""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [p for p in (os.path.join(_ROOT, "cnn-gp_b200"), _ROOT) if p not in sys.path]

from cnn_gp.block_store import merge_into, open_store  # noqa: E402


def merge(dest_file, src_files):
    with open_store(dest_file, "a") as dest_f:
        for path in src_files:
            with open_store(path, "r") as src_f:
                merge_into(dest_f, src_f)


if __name__ == '__main__':
    if len(sys.argv) < 3:
        print(f"Usage: {sys.argv[0]} dest_file [source_file1 source_file2 ...]")
        sys.exit(1)
    merge(sys.argv[1], sys.argv[2:])
