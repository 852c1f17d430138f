#!/bin/bash
# round 2, call L (8 GPUs): sharded-vs-single parity tests, bench lines at N=8 and N=4 (strong headline + weak extras + parity key)
set -x
nvidia-smi -L | wc -l
python -m pytest tests/test_gpu_mgpu.py -m gpu -q > gpurun_out/r2l_pytest_mgpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2l_pytest_mgpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2l_bench8.json 2> gpurun_out/r2l_bench8.err; echo "bench8 rc=$?"
tail -c 800 gpurun_out/r2l_bench8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 20 --warmup 5 --no-extra > gpurun_out/r2l_bench4.json 2> gpurun_out/r2l_bench4.err; echo "bench4 rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2l_bench8.json','gpurun_out/r2l_bench4.json'):
    try:
        d=json.load(open(f))
        print(f, 'value',d['value'],'ms',d['ms_per_step'],d['run']['steady'],d['run']['launches_per_step'], 'e2e', d.get('e2e',{}).get('value'), 'parity', d.get('parity'))
        print({k:(round(v['ms_per_call'],4),v['calls']) for k,v in d['kernel_families'].items()})
        for k,v in d.get('extra_configs',{}).items():
            if isinstance(v,dict): print(k,v.get('value'),v.get('ms_per_step'))
    except Exception as e: print(f, 'ERR', e)
PY
