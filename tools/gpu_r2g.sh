#!/bin/bash
# round 2, call G: state of HEAD after the re-entry -- GPU tests, full bench line, ncu --set full of the <float,20> passes (config 5)
set -x
nproc; nvidia-smi --query-gpu=name,memory.total --format=csv
python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2g_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2g_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2g_bench.json'))
print('value',d['value'],d['run']['steady'],d['run']['launches_per_step'], 'e2e', d['e2e']['value'])
print({k:(round(v['ms_per_call'],4),v['calls']) for k,v in d['kernel_families'].items()})
for k,v in d['extra_configs'].items():
    if isinstance(v,dict): print(k,v.get('value'),v.get('steady'),v.get('burst'),v.get('roofline',{}))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_update_classify|k_formk_cmprlb|k_subsm_lsinit" --launch-skip 104 --launch-count 4 -f -o gpurun_out/r2g_c5_full python tools/config5_rate.py > gpurun_out/r2g_c5_ncu.log 2>&1; echo "ncu rc=$?"; tail -5 gpurun_out/r2g_c5_ncu.log
ls -la gpurun_out/
