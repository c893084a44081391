#!/bin/bash
# Everything the round's evidence needs from ONE box, in the order the profiling recipe asks for
# (plain runs first, ncu afterwards): GPU tests, bench line, reference arm, launch list of the bench
# command, one `--set full` capture of the headline launch.   usage: scripts/round_capture.sh TAG
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
timeout 420 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $out/${tag}_pytest_gpu.log
timeout 300 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; cat $out/${tag}_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_ref.err
cat $out/${tag}_bench_reference_arm.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $out/${tag}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_launches.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:fused_kernel -c 1 -f \
    -o $out/${tag}_fused python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_full.log 2>&1
ls -la $out/${tag}_*
