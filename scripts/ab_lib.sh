#!/bin/bash
# Same-box A/B of two builds of libcnngp.so: scripts/ab_lib.sh BASE_SO [CONFIG N]   (rates alternate base / current)
base=$1; cfg=${2:-mnist_paper_convnet_gp}; n=${3:-10000}
for i in 1 2; do
  echo -n "base    "; CNNGP_LIB=$base timeout 150 python scripts/sweep_super.py $cfg $n 504 | grep -o '"Mpairs_per_s": [0-9.]*'
  echo -n "current "; timeout 150 python scripts/sweep_super.py $cfg $n 504 | grep -o '"Mpairs_per_s": [0-9.]*'
done
