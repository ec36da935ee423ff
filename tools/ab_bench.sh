#!/bin/bash
# A/B of the current build against tools/_ab/libwhisper_b200_prev.so (B200_LIB), alternating, same box:  bash tools/ab_bench.sh [rounds]
for i in $(seq 1 ${1:-2}); do
  for L in "" "B200_LIB=/root/repo/tools/_ab/libwhisper_b200_prev.so"; do
    env $L python bench.py --steps 3 --warmup 2 --cpu-baseline 0 --long-clip 0 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('${L:-current}'[-8:], round(d['value'],1), 'dec1 ms', round(d['stage_ms_per_step']['decoder1'],2), 'us/step', round(d['roofline']['us_per_step'],2), 'enc ms', round(d['stage_ms_per_step']['encoder'],2))"
  done
done
