#!/bin/bash
# round 2, first box: where the fused-net kernel stands before any change.  Plain bench lines of the
# three fused-net configs, then one `ncu --set full` capture (with source) of each.
out=gpurun_out/r02a
mkdir -p $out
for cfg in mnist_as_tf cifar10 mnist_paper_residual_cnn_gp; do
  timeout 300 python bench.py --config $cfg --n-images 6000 --steps 3 --warmup 3 --no-cpu-baseline > $out/bench_$cfg.json 2> $out/bench_$cfg.err
  cat $out/bench_$cfg.json
done
for cfg in mnist_as_tf cifar10; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:fnet_kernel -c 1 -f \
     -o $out/fnet_$cfg python bench.py --config $cfg --n-images 4000 --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu_$cfg.log 2>&1
  python scripts/ncu_summary.py $out/fnet_$cfg.ncu-rep $out/fnet_${cfg}_summary.json > /dev/null
done
ls -la $out
