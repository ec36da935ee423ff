#!/bin/bash
# full GPU suite + default bench + the configs[4] bench line (large-v3, 60-min file) after the ring-protocol fix
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2s_tests.txt
timeout 900 python bench.py > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
timeout 1500 python bench.py --shard-file 60 --model large-v3 --steps 2 --warmup 1 > gpurun_out/r2s_shard_large_v3.json 2> gpurun_out/r2s_shard.err
timeout 900 python bench.py --shard-file 60 --steps 2 --warmup 1 > gpurun_out/r2s_shard_turbo.json 2>> gpurun_out/r2s_shard.err
tail -3 gpurun_out/r2s_tests.txt; cut -c1-400 gpurun_out/r2s_bench.json; cut -c1-200 gpurun_out/r2s_shard_large_v3.json; cut -c1-200 gpurun_out/r2s_shard_turbo.json; grep -c "error" gpurun_out/r2s_shard.err gpurun_out/r2s_bench.err
