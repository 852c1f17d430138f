#!/bin/bash
# round 2, call O: bench line + quadratic per-family profile + ncu (launch list, --set full exported to csv on the box)
set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; echo "bench rc=$?"
python bench.py --workload quadratic --steps 12 --warmup 5 --no-extra --no-e2e --no-cpu --profile-out gpurun_out/r2o_quad_profile.json > gpurun_out/r2o_quad.json 2> gpurun_out/r2o_quad.err; echo "quad rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2o_quad.json'))
print('quad value',d['value'],d['ms_per_step'],d['run'])
print({k:(round(v['ms_per_call'],4),v['calls'],round(v['share'],3)) for k,v in d['kernel_families'].items()})
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2o_launches.csv python bench.py --steps 2 --warmup 1 --no-extra --no-e2e --no-cpu > gpurun_out/r2o_ncu_l.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"k_update_classify|k_formk_cmprlb|k_subsm_lsinit" --launch-skip 60 --launch-count 4 -f -o /tmp/r2o_c3_full python bench.py --steps 2 --warmup 1 --no-extra --no-e2e --no-cpu > gpurun_out/r2o_ncu_f.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/r2o_c3_full.ncu-rep --page raw --csv > gpurun_out/r2o_c3_full_raw.csv
timeout 900 ncu --set full --clock-control none -k regex:"k_update_classify|k_formk_cmprlb|k_subsm_lsinit|k_formk_delta" --launch-skip 130 --launch-count 6 -f -o /tmp/r2o_c5_full python tools/config5_rate.py > gpurun_out/r2o_c5_ncu.log 2>&1; echo "ncu c5 rc=$?"
ncu -i /tmp/r2o_c5_full.ncu-rep --page raw --csv > gpurun_out/r2o_c5_full_raw.csv
ls -la gpurun_out
