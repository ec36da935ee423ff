#!/bin/bash
# quick GPU check after a step-kernel change: decode / parity tests, two bench lines, a short prompt-launch stress
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_decode_gpu.py tests/test_fallback_gpu.py tests/test_turbo_parity_gpu.py tests/test_abi_parity.py -x -q -m gpu 2>&1 | grep -v Warning | tail -2 | cut -c1-300
B="python bench.py --steps 5 --warmup 3 --cpu-baseline 0 --word-timestamps-pass 0"
for v in a b; do
  timeout 600 $B 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(j['value'],1), round(j['roofline']['us_per_window_step'],2), round(j['long_clip']['value'],1))"
done
bash tools/soak3.sh 1 1500 2
