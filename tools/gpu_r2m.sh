#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; tail -2 gpurun_out/r2m_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2m_bench.json').read().strip().splitlines()[-1])
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'roofline frac',round(d['roofline']['frac'],3),'us',round(d['roofline']['us_per_step'],1),'traffic',d['roofline']['traffic'])
print('enc',d['encoder_roofline']['frac'],'stages',{k:round(v,2) for k,v in d['stage_ms_per_step'].items()})
print('words',d.get('word_timestamps'))
print('long',d.get('long_clip',{}).get('value'),d.get('long_clip',{}).get('decoder1_hbm_frac'))
print('cpu',d.get('cpu_baseline',{}).get('value'))
PY
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2m_ref.json 2> gpurun_out/r2m_ref.err; cut -c1-400 gpurun_out/r2m_ref.json
timeout 600 python bench.py --shard-file 4 --steps 2 --warmup 1 > gpurun_out/r2m_shard.json 2> gpurun_out/r2m_shard.err; cut -c1-600 gpurun_out/r2m_shard.json; tail -2 gpurun_out/r2m_shard.err
bash tools/profile_round.sh
