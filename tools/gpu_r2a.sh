#!/bin/bash
# round 2, call A: GPU tests, the new bench line, host facts
set -x
nproc; free -g | head -2; lscpu | grep -E "Model name|Socket|Thread" 
nvidia-smi --query-gpu=name,memory.total --format=csv
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2a_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2a_bench.err
head -c 3000 gpurun_out/r2a_bench.json
