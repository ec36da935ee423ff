#!/bin/bash
# round validation under gpurun: full GPU suite + smoke + default bench + the configs[4] lines + the reference arm (outputs in gpurun_out/r2s_*)
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2s_tests.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2s_smoke.txt 2>&1
timeout 900 python bench.py > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
timeout 1500 python bench.py --shard-file 60 --model large-v3 --steps 2 --warmup 1 > gpurun_out/r2s_shard_large_v3.json 2> gpurun_out/r2s_shard.err
timeout 900 python bench.py --shard-file 60 --steps 2 --warmup 1 > gpurun_out/r2s_shard_turbo.json 2>> gpurun_out/r2s_shard.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2s_ref.json 2> gpurun_out/r2s_ref.err
tail -3 gpurun_out/r2s_tests.txt; tail -1 gpurun_out/r2s_smoke.txt; python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2s_bench.json').read().strip().splitlines()[-1])
print(j['value'], j['e2e']['value'], j['stage_ms_per_step'], j['gpu_launches'])
print('roofline', j['roofline']['frac'], j['roofline']['us_per_window_step'], j['roofline']['traffic'], 'long', j['long_clip']['value'], j['long_clip'].get('decoder1_hbm_frac'), 'words', j['word_timestamps']['value'], j['word_timestamps']['relative_to_value'], j['clocks'], 'enc', j['encoder_roofline']['frac'], 'cpu', j['cpu_baseline'])
for f in ('r2s_shard_large_v3','r2s_shard_turbo','r2s_ref'):
    try:
        k=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, k.get('value'), k.get('ms_per_step'), k.get('impl'), k.get('cpu_baseline'))
    except Exception as e: print(f, 'failed', e)
PY
grep -c "error" gpurun_out/r2s_shard.err gpurun_out/r2s_bench.err
