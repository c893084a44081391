#!/bin/bash
# headline kernel: ring depth and a deliberate phase offset between the two halves of the CTA
out=gpurun_out/r02g; mkdir -p $out
run() { env "$@" timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2), 'M pairs/s')"; }
echo "default 12,2,2:"; run A=1
echo "12,2,3:"; run CNNGP_FUSED_VARIANT=12,2,3
for ns in 1000 2000 4000; do
  echo "12,2,2 skew $ns:"; run CNNGP_FUSED_SKEW_NS=$ns
  echo "12,2,3 skew $ns:"; run CNNGP_FUSED_VARIANT=12,2,3 CNNGP_FUSED_SKEW_NS=$ns
done
