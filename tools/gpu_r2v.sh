#!/bin/bash
set -x
python -m pytest tests/test_gpu_driver_api.py tests/test_gpu_rare_paths.py -m gpu -q -x 2>&1 | tail -25 | cut -c1-1200
python tools/small_n_latency.py gpurun_out/r2v_latency.json 2>&1 | tail -10 | cut -c1-600
