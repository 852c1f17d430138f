#!/bin/bash
# round 2, call B (2 GPUs): sharded-vs-single parity tests, bench line at N=2 (strong default + weak extras + parity key)
set -x
nvidia-smi -L
python -m pytest tests/test_gpu_mgpu.py -m gpu -x -q > gpurun_out/r2b_pytest_mgpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2b_pytest_mgpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2b_bench2.json 2> gpurun_out/r2b_bench2.err; echo "bench rc=$?"
tail -c 1200 gpurun_out/r2b_bench2.err
head -c 1500 gpurun_out/r2b_bench2.json
