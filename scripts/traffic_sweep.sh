#!/bin/bash
# DRAM traffic of one headline Gram launch per super-tile edge (ncu, three metrics, one pass).
# usage: scripts/traffic_sweep.sh CONFIG N EDGE...   -> gpurun_out/traffic_<config>_<edge>.csv
cfg=$1; n=$2; shift 2
mkdir -p gpurun_out
for e in "$@"; do
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
      --clock-control none -k regex:'fused_kernel|fnet_kernel' -c 1 --csv \
      --log-file gpurun_out/traffic_${cfg}_${e}.csv python scripts/sweep_super.py $cfg $n $e > gpurun_out/traffic_${cfg}_${e}.log 2>&1
  grep -E "dram__|gpu__time|lts__" gpurun_out/traffic_${cfg}_${e}.csv | awk -F'","' -v e=$e '{print e, $(NF-2), $(NF-1), $NF}'
done
