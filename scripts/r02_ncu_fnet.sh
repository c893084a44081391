#!/bin/bash
# one `ncu --set full` capture (with source) of the fused-net kernel per config.  usage: r02_ncu_fnet.sh OUTDIR N CONFIG...
out=$1; n=$2; shift 2
mkdir -p $out
for cfg in "$@"; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:fnet_kernel -c 1 -f \
     -o $out/fnet_$cfg python bench.py --config $cfg --n-images $n --steps 1 --warmup 3 --no-cpu-baseline --no-extra > $out/ncu_$cfg.log 2>&1
  python scripts/ncu_summary.py $out/fnet_$cfg.ncu-rep $out/fnet_${cfg}_summary.json > /dev/null
done
ls -la $out
