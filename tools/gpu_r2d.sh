#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
for st in 6 2 7; do
for w in 1 8; do B200_STEP_PROBE=$st timeout 300 python tools/step_timeline.py turbo $w > gpurun_out/r2d_probe_s${st}_w$w.txt 2>&1; done
done
grep -h -A30 "inner marks" gpurun_out/r2d_probe_s6_w8.txt | head -40
timeout 900 python -m pytest tests/test_decode_gpu.py -q -m gpu -k "batched or context_limit" 2>&1 | grep -v Warning | head -150 > gpurun_out/r2d_tests.txt
grep -n "^E " gpurun_out/r2d_tests.txt | head -20
