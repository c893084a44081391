#!/bin/bash
# Everything the round's evidence needs from ONE box, in the order the profiling recipe asks for (plain
# runs first, ncu afterwards): GPU tests, bench line, reference arm, launch list of the bench command,
# DRAM bytes of every Gram kernel at the size bench.py runs it (one-pass metrics), one `--set full`
# capture per kernel at 4 000 images (ncu saves and restores device memory between its ~40 passes:
# minutes per capture at the bench sizes).  The .ncu-rep files are summarised on the box and deleted
# (gpurun brings back at most 64 MiB).                               usage: scripts/round_capture.sh TAG [quick]
tag=${1:-r02}
out=gpurun_out/$tag
mkdir -p $out
if [ "$2" != "quick" ]; then
  timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $out/${tag}_pytest_gpu.log
  timeout 400 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; cat $out/${tag}_bench.json
  timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_ref.err
  cat $out/${tag}_bench_reference_arm.json
fi
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $out/${tag}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $out/${tag}_ncu_launches.log 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
timeout 300 ncu --metrics $M --clock-control none -k regex:fused_kernel -c 1 --csv --log-file $out/${tag}_dram_mnist_paper_convnet_gp@10000.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > /dev/null 2>&1
for cfg in mnist_as_tf cifar10 mnist_paper_residual_cnn_gp; do
  timeout 300 ncu --metrics $M --clock-control none -k regex:fnet_kernel -c 1 --csv --log-file $out/${tag}_dram_$cfg@6000.csv \
      python bench.py --config $cfg --n-images 6000 --steps 1 --warmup 3 --no-cpu-baseline --no-extra > /dev/null 2>&1
done
full() {  # name, kernel regex, command...
  name=$1; k=$2; shift 2
  timeout 500 ncu --set full --import-source on --clock-control none -k regex:$k -c 1 -f -o $out/${tag}_$name "$@" > $out/${tag}_ncu_$name.log 2>&1
  python scripts/ncu_summary.py $out/${tag}_$name.ncu-rep $out/${tag}_${name}_ncu_summary.json > /dev/null
  ncu -i $out/${tag}_$name.ncu-rep --page source --csv > /tmp/src_$name.csv 2>/dev/null
  python scripts/ncu_opcodes.py /tmp/src_$name.csv > $out/${tag}_${name}_opcodes.txt 2>&1
  rm -f $out/${tag}_$name.ncu-rep
}
full fused fused_kernel python bench.py --n-images 4000 --steps 1 --warmup 3 --no-cpu-baseline --no-extra
for cfg in mnist_as_tf cifar10 mnist_paper_residual_cnn_gp; do
  full fnet_$cfg fnet_kernel python bench.py --config $cfg --n-images 4000 --steps 1 --warmup 3 --no-cpu-baseline --no-extra
done
full sweep sweep_kernel python scripts/check_potrs.py 16384
du -sh gpurun_out; ls -la $out
