#!/bin/bash
# fused-net kernel A/B on one box: parity tests, then the three fused-net configs with the
# measurement switches of gram_fnet.cu (CNNGP_FNET_NOBLOCKS, CNNGP_FNET_NOCARRY).   usage: r02_fnet_ab.sh OUTDIR [N]
out=${1:-gpurun_out/r02c}
n=${2:-6000}
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $out/pytest_gpu.log
for cfg in mnist_as_tf cifar10 mnist_paper_residual_cnn_gp; do
  for variant in default noblocks nocarry; do
    env=""
    [ $variant = noblocks ] && env="CNNGP_FNET_NOBLOCKS=1"
    [ $variant = nocarry ] && env="CNNGP_FNET_NOCARRY=1"
    env $env timeout 300 python bench.py --config $cfg --n-images $n --steps 3 --warmup 3 --no-cpu-baseline --no-extra \
        > $out/bench_${cfg}_$variant.json 2> $out/bench_${cfg}_$variant.err
    python - <<PY
import json
try:
    d = json.load(open("$out/bench_${cfg}_$variant.json"))
    print("$cfg $variant", round(d["value"] / 1e6, 2), "M pairs/s  frac", round(d["roofline"]["frac"], 4))
except Exception as e:
    print("$cfg $variant FAILED", e)
PY
  done
done
