#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
B="python bench.py --model large-v3 --minutes 4 --steps 2 --warmup 1 --cpu-baseline 0 --long-clip 0"
rm -f gpurun_out/r2h_bench.txt
for cfg in "B200_STEP_WARPS=8" "B200_STEP_WARPS=8 B200_BATCH_WINDOWS=4" "B200_STEP_WARPS=8 B200_BATCH_WINDOWS=2" "B200_STEP_IMPL=mega"; do
  echo "== $cfg" >> gpurun_out/r2h_bench.txt
  env $cfg timeout 900 $B 2>gpurun_out/r2h_bench_err.txt | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('rtfx',round(d['value'],1),'dec1 ms',round(d['stage_ms_per_step']['decoder1'],2),'us/winstep',round(d['roofline']['us_per_step'],2),'enc ms',round(d['stage_ms_per_step']['encoder'],2),'d256',round(d['stage_ms_per_step']['decoder256'],2))" >> gpurun_out/r2h_bench.txt 2>&1
  tail -3 gpurun_out/r2h_bench_err.txt >> gpurun_out/r2h_bench.txt
done
cat gpurun_out/r2h_bench.txt
