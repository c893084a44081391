#!/bin/bash
# A/B of the tile hand-out order (CNNGP_TILE_ORDER=static | default counter): rate and DRAM traffic.
# usage: scripts/tile_order_ab.sh CONFIG N
cfg=$1; n=$2
mkdir -p gpurun_out
for order in static counter; do
  export CNNGP_TILE_ORDER=$order
  echo "== $order"; timeout 300 python scripts/sweep_super.py $cfg $n 504
  timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
      --clock-control none -k regex:'fused_kernel|fnet_kernel' -c 1 --csv \
      --log-file gpurun_out/traffic_${cfg}_${order}.csv python scripts/sweep_super.py $cfg $n 504 > gpurun_out/traffic_${cfg}_${order}.log 2>&1
  grep -E "dram__|gpu__time|lts__" gpurun_out/traffic_${cfg}_${order}.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done
