#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
B200_DECODE_LANES=1 B200_STEP_CTAS=74 timeout 300 python tools/step_timeline.py turbo 1 > gpurun_out/r2j_new74.txt 2>&1
B200_STEP_IMPL=mega B200_MEGA_CTAS=74 B200_DECODE_LANES=1 timeout 300 python tools/step_timeline_mega.py turbo > gpurun_out/r2j_mega74.txt 2>&1
B200_DECODE_LANES=1 timeout 300 python tools/step_timeline.py turbo 1 > gpurun_out/r2j_new148.txt 2>&1
B200_STEP_IMPL=mega B200_DECODE_LANES=1 timeout 300 python tools/step_timeline_mega.py turbo > gpurun_out/r2j_mega148.txt 2>&1
grep -h "step total\|mean us" gpurun_out/r2j_*.txt
