#!/bin/bash
# Everything the round's evidence needs from ONE box, in the order the profiling recipe asks for (plain
# runs first, ncu afterwards): GPU tests, bench line, reference arm, launch list of the bench command,
# DRAM bytes of every Gram call at the size bench.py runs it (one-pass metrics, all launches of the call),
# one `--set full` capture per kernel (ncu saves and restores device memory between its ~40 passes:
# small sizes).  The .ncu-rep files are summarised on the box and deleted (gpurun brings back at most
# 64 MiB).                                                          usage: scripts/round_capture.sh TAG [quick]
tag=${1:-r02}
out=gpurun_out/$tag
mkdir -p $out
if [ "$2" != "quick" ]; then
  timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -3 | tee $out/${tag}_pytest_gpu.log
  timeout 500 python bench.py > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench.err; cat $out/${tag}_bench_1gpu.json
  timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_ref.err
  cat $out/${tag}_bench_reference_arm.json
  for cfg in mnist_as_tf cifar10 mnist_paper_residual_cnn_gp; do
    timeout 200 python bench.py --config $cfg --n-images 6000 --steps 4 --warmup 3 --no-cpu-baseline --no-extra > $out/${tag}_bench_$cfg.json 2>> $out/${tag}_bench.err
  done
fi
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $out/${tag}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $out/${tag}_ncu_launches.log 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
timeout 300 ncu --metrics $M --clock-control none -k regex:fused_kernel -c 1 --csv --log-file "$out/${tag}_dram_mnist_paper_convnet_gp@10000.csv" \
    python scripts/gram_times.py mnist_paper_convnet_gp 10000 1 > /dev/null 2>&1
for cfg in mnist_as_tf cifar10 mnist_paper_residual_cnn_gp; do
  timeout 300 ncu --metrics $M --clock-control none -k regex:fnet_kernel --csv --log-file "$out/${tag}_dram_$cfg@6000.csv" \
      python scripts/gram_times.py $cfg 6000 1 > /dev/null 2>&1
done
export CNNGP_FNET_HANDOFF_MB=256
scripts/ncu_one.sh $out ${tag}_fused fused_kernel 1 python scripts/gram_times.py mnist_paper_convnet_gp 4000 2
scripts/ncu_one.sh $out ${tag}_fnetA_mnist_as_tf fnet_kernel 2 python scripts/gram_times.py mnist_as_tf 2000 2
scripts/ncu_one.sh $out ${tag}_fnetB_mnist_as_tf fnet_kernel 3 python scripts/gram_times.py mnist_as_tf 2000 2
scripts/ncu_one.sh $out ${tag}_fnetA_cifar10 fnet_kernel 2 python scripts/gram_times.py cifar10 2000 2
scripts/ncu_one.sh $out ${tag}_fnetB_cifar10 fnet_kernel 3 python scripts/gram_times.py cifar10 2000 2
scripts/ncu_one.sh $out ${tag}_fnet_mnist_paper_residual_cnn_gp fnet_kernel 1 python scripts/gram_times.py mnist_paper_residual_cnn_gp 4000 2
scripts/ncu_one.sh $out ${tag}_sweep sweep_kernel 0 python scripts/check_potrs.py 16384
du -sh gpurun_out; ls -la $out
