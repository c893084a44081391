#!/usr/bin/env bash
# One save_kernel worker per visible GPU writing its own HDF5 file, then merge_h5_files into the
# first worker's file, then classify_gp on it: the per-worker-file flow of the reference's
# exp_mnist_resnet/run.bash:24-49 (file names NN_nwMM.h5 kept).  exp_mnist_resnet/run.py is the
# same pipeline as one torchrun program that exchanges blocks over NCCL instead of files.
#
#   CONFIG=synthetic OUT=/tmp/grams BATCH=200 DATASETS=/tmp/datasets GPUS="0 1" bash exp_mnist_resnet/run.bash
set -euo pipefail

CONFIG="${CONFIG:-synthetic}"
OUT="${OUT:-/tmp/cnngp_grams_$$}"
BATCH="${BATCH:-200}"
DATASETS="${DATASETS:-/tmp/datasets/}"
GPUS="${GPUS:-$(nvidia-smi --query-gpu=index --format=csv,noheader | tr '\n' ' ')}"

read -r -a gpu_list <<< "$GPUS"
workers=${#gpu_list[@]}
[ "$workers" -gt 0 ] || { echo "no GPUs: set GPUS=\"0 1 ...\"" >&2; exit 1; }
[ ! -e "$OUT" ] || { echo "refusing to overwrite $OUT" >&2; exit 1; }
mkdir -p "$OUT"
cd "$(dirname "$0")/.."

file_of() { printf '%s/%02d_nw%02d.h5' "$OUT" "$1" "$workers"; }

echo "[$(date +%T)] $workers save_kernel worker(s), config $CONFIG, tiles of $BATCH"
pids=()
for rank in "${!gpu_list[@]}"; do
    CUDA_VISIBLE_DEVICES="${gpu_list[$rank]}" python -m exp_mnist_resnet.save_kernel \
        --config="$CONFIG" --datasets_path="$DATASETS" --batch_size="$BATCH" \
        --n_workers="$workers" --worker_rank="$rank" --out_path="$(file_of "$rank")" &
    pids+=($!)
done
failed=0
for pid in "${pids[@]}"; do wait "$pid" || failed=1; done
[ "$failed" -eq 0 ] || { echo "a worker failed" >&2; exit 1; }

dest="$(file_of 0)"
if [ "$workers" -gt 1 ]; then
    echo "[$(date +%T)] merging into $dest"
    sources=()
    for rank in $(seq 1 $((workers - 1))); do sources+=("$(file_of "$rank")"); done
    python -m exp_mnist_resnet.merge_h5_files "$dest" "${sources[@]}"
fi

echo "[$(date +%T)] classify_gp on $dest"
python -m exp_mnist_resnet.classify_gp --config="$CONFIG" --datasets_path="$DATASETS" --in_path="$dest"
