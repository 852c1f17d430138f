#!/bin/bash
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2f_pytest.log
