#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 3000 python -m pytest tests -q -m gpu 2>&1 | grep -v Warning | tail -30 > gpurun_out/r2k_tests.txt
tail -8 gpurun_out/r2k_tests.txt
python bench.py --steps 3 --warmup 2 --cpu-baseline 0 2>gpurun_out/r2k_err.txt | tail -1 > gpurun_out/r2k_bench.json; python -c "
import json; d=json.load(open('gpurun_out/r2k_bench.json')); print('rtfx',round(d['value'],1),'e2e',round(d['e2e']['value'],1),d['stage_ms_per_step'],'long',round(d['long_clip']['value'],1))"
make -s -C . -f /dev/null 2>/dev/null; ./whisper.coreml_b200/build/abi_smoke /tmp/b200_bench_weights_turbo_0 32 4 1280 128 51866 5 0 > gpurun_out/r2k_abi_smoke.txt 2>&1; tail -12 gpurun_out/r2k_abi_smoke.txt
