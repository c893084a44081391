#!/bin/bash
# Build a variant of libcnngp.so with extra nvcc flags on one source: scripts/build_variant.sh OUT.so SOURCE.cu [-DFOO=1 ...]
# (the other objects come from cnn-gp_b200/build; run cnn-gp_b200/build.py first)
set -e
out=$1; src=$2; shift 2
root=$(dirname $(dirname $(readlink -f $0)))/cnn-gp_b200
obj=/tmp/variant_$$.o
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -ccbin /usr/bin/g++ "$@" -c $root/csrc/$src -o $obj
others=$(ls $root/build/*.o | grep -v "/${src%.cu}.o")
nvcc -shared -o $out $obj $others -ccbin /usr/bin/g++ -lcudart
rm -f $obj
echo built $out
