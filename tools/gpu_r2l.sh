#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_fallback_gpu.py tests/test_decode_gpu.py tests/test_configs_gpu.py tests/test_turbo_parity_gpu.py tests/test_mel_gpu.py tests/test_word_times_gpu.py -q -m gpu -x 2>&1 | grep -v Warning | tail -40 > gpurun_out/r2l_tests.txt
tail -25 gpurun_out/r2l_tests.txt
