set -x
B="python bench.py --steps 1 --warmup 1 --cpu-baseline 0 --long-clip 0"
$B --sample-len 24 > gpurun_out/plain_r1f.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r1f.csv $B --sample-len 24 > gpurun_out/ncu1f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:flash_attn -s 36 -c 1 -o gpurun_out/prof_flash_r1f $B --sample-len 2 > gpurun_out/ncu_ff.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 146 -c 4 -o gpurun_out/prof_gemm_r1f $B --sample-len 2 > gpurun_out/ncu_gf.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:decoder_mega -s 30 -c 1 -o gpurun_out/prof_mega_r1f $B --sample-len 24 > gpurun_out/ncu_mf.log 2>&1
tail -1 gpurun_out/plain_r1f.log | cut -c1-300
