#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
for u in 13 4; do for w in 1 2; do B200_STEP_COPYU=$u B200_STEP_PROBE=2 timeout 300 python tools/step_timeline.py turbo $w > gpurun_out/r2f_probe_u${u}_w$w.txt 2>&1; echo "== copy U=$u W=$w"; grep -h -A30 "inner marks" gpurun_out/r2f_probe_u${u}_w$w.txt | head -34; grep "step total" gpurun_out/r2f_probe_u${u}_w$w.txt; done; done
for w in 1 2; do B200_STEP_PROBE=6 timeout 300 python tools/step_timeline.py turbo $w > gpurun_out/r2f_probe_ln_w$w.txt 2>&1; echo "== LN+gemv (mlp1) W=$w"; grep -h -A34 "inner marks" gpurun_out/r2f_probe_ln_w$w.txt | head -38; done
