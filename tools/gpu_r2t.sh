#!/bin/bash
# round 2, call T (2 GPUs): sharded tie replay + the sharded parity tests
set -x
python -m pytest tests/test_gpu_mgpu.py -m gpu -q -x > gpurun_out/r2t_pytest_mgpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2t_pytest_mgpu.log | cut -c1-1500
