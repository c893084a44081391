#!/bin/bash
# 8 x B200: the torchrun bench line (with extra), then BASELINE config 3 at full size through the torchrun
# pipeline (60 000 + 2 000 + 10 000 synthetic images, mnist_paper_residual_cnn_gp): row shards, no gather,
# distributed Cholesky + sweeps + predictions.   usage: r02_8gpu.sh N OUTDIR
n=${1:-8}; out=${2:-gpurun_out/r02q}; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR bench.py --gpus $n --steps 5 --warmup 3 > $out/bench_${n}gpu.json 2> $out/bench_${n}gpu.err; tail -c 3000 $out/bench_${n}gpu.json; tail -3 $out/bench_${n}gpu.err
CNNGP_SYNTH_MODEL=mnist_paper_residual_cnn_gp CNNGP_SYNTH_TRAIN=60000 CNNGP_SYNTH_VAL=2000 CNNGP_SYNTH_TEST=10000 CNNGP_DIST_VERBOSE=1 \
  timeout 900 $TR -m exp_mnist_resnet.run --config=synthetic --batch_size=200 2>&1 | grep -v "^W\|^\*\*\|OMP_NUM" | tail -14 | tee $out/run60k_residual_${n}gpu.log
