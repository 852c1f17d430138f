#!/bin/bash
set -x
python tools/trace_digest.py gpurun_out/r2m_digest.json > gpurun_out/r2m_digest.log 2>&1
python tools/config5_rate.py > gpurun_out/r2m_c5.log 2>&1; tail -3 gpurun_out/r2m_c5.log

python -m pytest tests -m gpu -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2m_pytest.log
