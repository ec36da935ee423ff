#!/bin/bash
# full GPU suite + default bench (+ optional configs[4] lines)
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2s_tests.txt
timeout 900 python bench.py > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
tail -3 gpurun_out/r2s_tests.txt; python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2s_bench.json').read().strip().splitlines()[-1])
print(j['value'], j['e2e']['value'], j['stage_ms_per_step'])
print('roofline', j['roofline']['frac'], j['roofline']['us_per_window_step'], 'long', j['long_clip']['value'], j['long_clip'].get('encoder_tflops'), 'words', j['word_timestamps']['value'], j['clocks'])
print({k: j[k] for k in j if k.startswith('encoder')})
PY
