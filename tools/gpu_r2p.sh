#!/bin/bash
# BASELINE configs[4]: large-v3 dims, ONE 1-hour synthetic file, fixed windows sharded rank::world_size
cd /root/repo; mkdir -p gpurun_out
N=${1:-1}
if [ "$N" = "1" ]; then
  timeout 2400 python bench.py --shard-file 60 --model large-v3 --steps 1 --warmup 1 > gpurun_out/r2p_shard_n1.json 2> gpurun_out/r2p_shard_n1.err
else
  timeout 2400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --shard-file 60 --model large-v3 --steps 1 --warmup 1 > gpurun_out/r2p_shard_n$N.json 2> gpurun_out/r2p_shard_n$N.err
fi
tail -3 gpurun_out/r2p_shard_n$N.err; cut -c1-700 gpurun_out/r2p_shard_n$N.json
