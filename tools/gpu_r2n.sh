#!/bin/bash
# round 2, call N: full GPU suite, the bench line, ncu launch list + --set full of the headline passes and of the <float,20> passes
set -x
python -m pytest tests -m gpu -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2n_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2n_bench.json'))
print('value',d['value'],d['run']['steady'],d['run']['launches_per_step'], 'e2e', d['e2e']['value'])
print({k:(round(v['ms_per_call'],4),v['calls']) for k,v in d['kernel_families'].items()})
for k,v in d['extra_configs'].items():
    if isinstance(v,dict): print(k,v.get('value'),v.get('steady'),v.get('burst'),v.get('roofline',{}))
PY
python bench.py --steps 2 --warmup 1 --no-extra --no-e2e --no-cpu > gpurun_out/r2n_short.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2n_launches.csv python bench.py --steps 2 --warmup 1 --no-extra --no-e2e --no-cpu > gpurun_out/r2n_ncu_l.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_update_classify|k_formk_cmprlb|k_subsm_lsinit" --launch-skip 60 --launch-count 4 -f -o gpurun_out/r2n_c3_full python bench.py --steps 2 --warmup 1 --no-extra --no-e2e --no-cpu > gpurun_out/r2n_ncu_f.log 2>&1; echo "ncu full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_update_classify|k_formk_cmprlb|k_subsm_lsinit|k_formk_delta" --launch-skip 130 --launch-count 6 -f -o gpurun_out/r2n_c5_full python tools/config5_rate.py > gpurun_out/r2n_c5_ncu.log 2>&1; echo "ncu c5 rc=$?"
