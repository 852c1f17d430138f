#!/bin/bash
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2c_pytest.log
python tools/small_n_latency.py gpurun_out/r2c_latency.json 2>&1 | tail -8
python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c_bench.json'))
print('value',d['value'],d['run']['steady'],d['run']['launches_per_step'])
print({k:(round(v['ms_per_call'],4),v['calls']) for k,v in d['kernel_families'].items()})
for k,v in d['extra_configs'].items():
    if isinstance(v,dict): print(k,v.get('value'),v.get('steady'),v.get('burst'),v.get('roofline',{}).get('frac'))
PY
