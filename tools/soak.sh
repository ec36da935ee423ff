#!/bin/bash
# soak (under gpurun): many full-length decodes in the lane shapes the unit tests only touch briefly; any protocol time-out of
# the step kernel shows up as "gave up in wait" + the per-CTA stage table in the log
cd /root/repo; mkdir -p gpurun_out
run() { echo "== $1 :: ${@:2}"; env $1 timeout 1500 python bench.py ${@:2} 2>gpurun_out/soak.err | python -c "
import json,sys
t=sys.stdin.read().strip().splitlines()
print(round(json.loads(t[-1])['value'],1) if t and t[-1].startswith('{') else 'FAILED')"; grep -c "gave up\|error" gpurun_out/soak.err; grep "gave up" gpurun_out/soak.err | head -2; }
S="--cpu-baseline 0 --word-timestamps-pass 0 --long-clip 0"
run B200_DECODE_LANES=8 --shard-file 60 --steps 4 --warmup 1
run B200_DECODE_LANES=8 --shard-file 60 --model large-v3 --steps 2 --warmup 1
run B200_DECODE_LANES=1 --shard-file 30 --steps 2 --warmup 1
run B200_DECODE_LANES=3 --shard-file 30 --steps 2 --warmup 1
run B200_DECODE_LANES=2 $S --steps 40 --warmup 3
run B200_DECODE_LANES=1 $S --steps 20 --warmup 3
