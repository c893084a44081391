#!/bin/bash
# final evidence of round 2 on one box: GPU tests, the bench line and the reference arm, the launch list of the
# bench command, --set full captures of the kernels that changed after the first capture (variance rows,
# triangular sweeps, predict)                               usage: scripts/r02_final_capture.sh [TAG]
tag=${1:-r02f}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 2>&1 | tail -3 | tee $out/${tag}_pytest_gpu.log
timeout 600 python bench.py > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench.err; cat $out/${tag}_bench_1gpu.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_ref.err
cat $out/${tag}_bench_reference_arm.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $out/${tag}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $out/${tag}_ncu_launches.log 2>&1
scripts/ncu_one.sh $out ${tag}_variance variance_kernel 1 python scripts/variance_times.py mnist_paper_convnet_gp 10000 1
scripts/ncu_one.sh $out ${tag}_sweep sweep_kernel 0 python scripts/check_potrs.py 16384
scripts/ncu_one.sh $out ${tag}_predict predict_dmma_kernel 1 python scripts/predict_times.py 4096 32768
python scripts/predict_times.py > $out/${tag}_predict_times.log 2>&1
python scripts/potrs_times.py > $out/${tag}_potrs_times.log 2>&1
ls -la $out
