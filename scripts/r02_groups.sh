#!/bin/bash
out=gpurun_out/r02h; mkdir -p $out
run() { env "$@" timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra 2>$out/err.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2), 'M pairs/s')"; tail -2 $out/err.log; }
echo "default 12,2,3:"; run A=1
echo "12,2,2:"; run CNNGP_FUSED_VARIANT=12,2,2
echo "two groups (122):"; run CNNGP_FUSED_VARIANT=122,2,2
timeout 300 python -m pytest tests/test_gpu_gram.py -x -q 2>&1 | tail -3
CNNGP_FUSED_VARIANT=122,2,2 timeout 300 python -m pytest tests/test_gpu_gram.py -x -q 2>&1 | tail -3
for cfg in mnist_as_tf mnist_paper_residual_cnn_gp; do
 for v in A=1 CNNGP_FNET_NST2=1; do echo $cfg $v; env $v timeout 200 python bench.py --config $cfg --n-images 6000 --steps 3 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2), 'M pairs/s')"; done; done
