#!/bin/bash
# N-GPU box: distributed solve check, the torchrun pipeline on the synthetic config, the bench line.  usage: r02_multi.sh N OUTDIR
n=${1:-2}; out=${2:-gpurun_out/r02k}; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511"
timeout 300 python scripts/check_potrs.py 4096 32768 2>&1 | tail -6 | tee $out/potrs.log
timeout 600 $TR scripts/check_dist_solve.py 8192 32768 2>&1 | grep -v "^W\|^\*\*" | tail -4 | tee $out/dist_solve.log
CNNGP_SYNTH_TRAIN=6000 CNNGP_SYNTH_VAL=1000 CNNGP_SYNTH_TEST=2000 CNNGP_DIST_VERBOSE=1 timeout 600 $TR -m exp_mnist_resnet.run --config=synthetic --batch_size=200 2>&1 | grep -v "^W\|^\*\*" | tail -12 | tee $out/run_dist.log
CNNGP_SYNTH_TRAIN=6000 CNNGP_SYNTH_VAL=1000 CNNGP_SYNTH_TEST=2000 timeout 600 $TR -m exp_mnist_resnet.run --config=synthetic --batch_size=200 --nodist_solve --out_path=/tmp/k_$n.h5 2>&1 | grep -v "^W\|^\*\*" | tail -8 | tee $out/run_gather.log
CNNGP_SYNTH_TRAIN=6000 CNNGP_SYNTH_VAL=1000 CNNGP_SYNTH_TEST=2000 timeout 600 python -m exp_mnist_resnet.run --config=synthetic --batch_size=200 2>&1 | tail -5 | tee $out/run_1gpu.log
timeout 900 $TR bench.py --gpus $n --steps 5 --warmup 3 > $out/bench_${n}gpu.json 2> $out/bench_${n}gpu.err; tail -c 2500 $out/bench_${n}gpu.json; tail -3 $out/bench_${n}gpu.err
