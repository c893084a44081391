#!/bin/bash
# Same-box comparison of several builds of libcnngp.so: scripts/ab_libs.sh OUTLOG N "CFG ..." LIB.so [LIB.so ...]   (two rounds, alternating)
log=$1; n=$2; cfgs=$3; shift 3
mkdir -p $(dirname $log)
for cfg in $cfgs; do
  for i in 1 2; do
    for lib in "$@"; do
      echo -n "$cfg $(basename $lib) "; CNNGP_LIB=$lib timeout 200 python scripts/sweep_super.py $cfg $n 504 | grep -o '"Mpairs_per_s": [0-9.]*'
    done
  done
done 2>&1 | tee $log
