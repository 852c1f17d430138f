#!/bin/bash
# One B200: the GPU test suite, the bench line, and the ncu captures that profiles/ is made from.
#   gpurun --timeout 2400 -- 'bash tools/gpu_single.sh TAG'
# Outputs under gpurun_out/TAG_*: pytest log, bench JSON, launch list (csv), --set full raw csv of the headline passes
# (<double,10> at n = 1e8) and of the <float,20> passes + k_formk_delta (n = 4e8).  The .ncu-rep files stay on the box
# (too large to bring back); tools/ncu_summary.py turns the csv files into profiles/TAG_*.md and profiles/ncu_traffic.json.
set -x
TAG=${1:-cur}
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/${TAG}_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('value',d['value'],d['run']['steady'],d['run']['launches_per_step'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'])
print({k:(round(v['ms_per_call'],4),v['calls']) for k,v in d['kernel_families'].items()})
for k,v in d['extra_configs'].items():
    if isinstance(v,dict): print(k,v.get('value'),v.get('steady'),v.get('burst'),(v.get('roofline') or {}).get('frac'))
    else: print(k, str(v)[:300])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-extra --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_l.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"k_update_classify|k_formk_cmprlb|k_subsm_lsinit" --launch-skip 60 --launch-count 4 -f -o /tmp/${TAG}_c3_full python bench.py --steps 2 --warmup 1 --no-extra --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_f.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/${TAG}_c3_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_c3_full_raw.csv
timeout 900 ncu --set full --clock-control none -k regex:"k_update_classify|k_formk_cmprlb|k_subsm_lsinit|k_formk_delta" --launch-skip 130 --launch-count 6 -f -o /tmp/${TAG}_c5_full python tools/config5_rate.py > gpurun_out/${TAG}_c5_ncu.log 2>&1; echo "ncu c5 rc=$?"
ncu -i /tmp/${TAG}_c5_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_c5_full_raw.csv
ls -la gpurun_out
