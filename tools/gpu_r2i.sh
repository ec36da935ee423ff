#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 2 --cpu-baseline 0 --long-clip 4"
rm -f gpurun_out/r2i_bench.txt
for cfg in ""; do
  echo "== $cfg" >> gpurun_out/r2i_bench.txt
  env $cfg timeout 600 $B 2>gpurun_out/r2i_bench_err.txt | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('rtfx',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'dec1 ms',round(d['stage_ms_per_step']['decoder1'],2),'us/winstep',round(d['roofline']['us_per_step'],2),'enc ms',round(d['stage_ms_per_step']['encoder'],2),'d256',round(d['stage_ms_per_step']['decoder256'],2),'samp',round(d['stage_ms_per_step']['sampling'],2),'long',round(d['long_clip']['value'],1),round(d['long_clip']['decoder1_hbm_frac'],3))" >> gpurun_out/r2i_bench.txt 2>&1
  tail -3 gpurun_out/r2i_bench_err.txt >> gpurun_out/r2i_bench.txt
done
cat gpurun_out/r2i_bench.txt
timeout 1200 python -m pytest tests/test_decode_gpu.py tests/test_abi_parity.py tests/test_audio_ingest.py tests/test_configs_gpu.py -q -m gpu 2>&1 | grep -v Warning | tail -30 > gpurun_out/r2i_tests.txt
tail -8 gpurun_out/r2i_tests.txt
B200_DECODE_LANES=1 B200_STEP_CTAS=74 timeout 300 python tools/step_timeline.py turbo 1 > gpurun_out/r2i_new74.txt 2>&1
B200_DECODE_LANES=1 timeout 300 python tools/step_timeline.py turbo 1 > gpurun_out/r2i_new148.txt 2>&1
B200_DECODE_LANES=1 timeout 300 python tools/step_timeline.py turbo 2 > gpurun_out/r2i_new148_w2.txt 2>&1
grep -h "step total\|mean us" gpurun_out/r2i_new*.txt
