#!/bin/bash
# N B200s of one box: the sharded parity tests (tests/test_gpu_mgpu.py: sharded vs single GPU, tie problems) and the bench
# line at N GPUs (strong headline + weak extras + parity key).
#   gpurun --gpus N --timeout 1500 -- 'bash tools/gpu_multi.sh TAG N [notests]'
set -x
TAG=${1:-cur}; N=${2:-2}
if [ "$3" != "notests" ]; then
python -m pytest tests/test_gpu_mgpu.py -m gpu -q > gpurun_out/${TAG}_pytest_mgpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest_mgpu.log | cut -c1-1500
fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench${N}.json 2> gpurun_out/${TAG}_bench${N}.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench${N}.json'))
print('value',d['value'],'ms',d['ms_per_step'],d['run']['steady'],d['run']['launches_per_step'], 'e2e', d.get('e2e',{}).get('value'), 'parity', d.get('parity'))
print({k:(round(v['ms_per_call'],4),v['calls']) for k,v in d['kernel_families'].items()})
for k,v in d.get('extra_configs',{}).items():
    if isinstance(v,dict): print(k,v.get('value'),v.get('ms_per_step'))
PY
