#!/bin/bash
# GPU box helper: bf16 parity tests + a short bench with the per-stage table; results under gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py -x -q 2>&1 | tail -5
timeout 600 python bench.py --no-cpu-baseline --no-fp32 --no-latency --no-extra --steps 10 --warmup 3 > gpurun_out/quick.json 2> gpurun_out/quick.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/quick.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'])
print(' '.join('%s=%.1f' % (s['stage'], s['us_per_image']) for s in d.get('stages', [])))
PY
