#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 300 python tools/repro_batch.py tiny 8 6 > gpurun_out/r2b_repro.txt 2>&1; tail -5 gpurun_out/r2b_repro.txt
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/repro_batch.py tiny 8 3 > gpurun_out/r2b_sanitizer.txt 2>&1; grep -m 40 -E "Invalid|at |by thread|Address|ERROR SUMMARY|beam" gpurun_out/r2b_sanitizer.txt | head -60
timeout 600 python -m pytest tests/test_turbo_parity_gpu.py -q -m gpu -k "step_fused" 2>&1 | head -80 > gpurun_out/r2b_fused.txt; grep -n "Error\|error\|assert" gpurun_out/r2b_fused.txt | head
