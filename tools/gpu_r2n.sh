#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 3000 python -m pytest tests -q -m gpu 2>&1 | grep -v Warning | tail -30 > gpurun_out/r2n_tests.txt
tail -8 gpurun_out/r2n_tests.txt
for cfg in "" "B200_ENCODER_GRAPH=0"; do
env $cfg timeout 900 python bench.py --steps 5 --warmup 3 --cpu-baseline 0 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; tail -2 gpurun_out/r2n_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2n_bench.json').read().strip().splitlines()[-1])
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'roofline frac',round(d['roofline']['frac'],3),'us',round(d['roofline']['us_per_step'],1),'traffic',d['roofline']['traffic'])
print('enc',round(d['encoder_roofline']['frac'],4),'stages',{k:round(v,2) for k,v in d['stage_ms_per_step'].items()})
w=d.get('word_timestamps'); print('words',round(w['value'],1),round(w['relative_to_value'],3),'align ms',round(w['stage_ms_per_step']['align'],2))
print('long',round(d['long_clip']['value'],1),round(d['long_clip']['decoder1_hbm_frac'],3),round(d['long_clip']['encoder_tflops'],1))
PY
done
