#!/bin/bash
# one `ncu --set full` capture of one kernel (regex) of a command, summarised on the box (summary JSON + opcode histogram);
# the .ncu-rep is deleted (gpurun brings back at most 64 MiB).      usage: scripts/ncu_one.sh OUTDIR NAME KERNEL_REGEX SKIP command...
# (SKIP matching launches are passed over first: the split fused-net launches alternate phase A / phase B)
out=$1; name=$2; k=$3; skip=$4; shift 4
mkdir -p $out
timeout 400 ncu --set full --import-source on --clock-control none -k "regex:$k" --launch-skip $skip -c 1 -f -o $out/$name "$@" > $out/ncu_$name.log 2>&1
python scripts/ncu_summary.py $out/$name.ncu-rep $out/${name}_ncu_summary.json > /dev/null
ncu -i $out/$name.ncu-rep --page source --csv > /tmp/src_$name.csv 2>/dev/null
python scripts/ncu_opcodes.py /tmp/src_$name.csv > $out/${name}_opcodes.txt 2>&1
rm -f $out/$name.ncu-rep
