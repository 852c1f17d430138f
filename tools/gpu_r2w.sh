#!/bin/bash
set -x
python tools/trace_digest.py gpurun_out/r2w_digest.json > gpurun_out/r2w_digest.log 2>&1
python tools/iter_breakdown.py quadratic 125000000 8 2>&1 | tail -9
python -m pytest tests -m gpu -q -x > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2w_pytest.log
