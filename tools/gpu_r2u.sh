#!/bin/bash
set -x
python -m pytest tests -m gpu -q > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2u_pytest.log
python tools/small_n_latency.py gpurun_out/r2u_latency.json 2>&1 | tail -8
