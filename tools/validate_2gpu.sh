#!/bin/bash
# N = 2 under `gpurun --gpus 2`: default bench (weak scaling) and the configs[4] file (strong scaling); outputs in gpurun_out/r2x_*
cd /root/repo; mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $T --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2x_bench_n2.json 2> gpurun_out/r2x_bench_n2.err
timeout 1200 $T --master-port 29512 bench.py --gpus 2 --shard-file 60 --model large-v3 --steps 2 --warmup 1 > gpurun_out/r2x_shard_large_v3_n2.json 2> gpurun_out/r2x_shard_n2.err
timeout 900 $T --master-port 29513 bench.py --gpus 2 --shard-file 60 --steps 2 --warmup 1 > gpurun_out/r2x_shard_turbo_n2.json 2>> gpurun_out/r2x_shard_n2.err
for f in gpurun_out/r2x_*_n2.json; do echo $f; tail -1 $f | cut -c1-330; done
grep -c "error" gpurun_out/r2x_bench_n2.err gpurun_out/r2x_shard_n2.err
