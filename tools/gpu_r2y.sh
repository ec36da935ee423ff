#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
B200_STEP_PROBE=32 B200_DECODE_LANES=1 B200_STEP_CTAS=74 timeout 600 python tools/step_timeline.py turbo 1 2>/dev/null | tail -22 | cut -c1-200
