#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_batch.py -m gpu -x -q 2>&1 | tail -15
python tools/batch_rate.py 1000 25 5
python tools/batch_rate.py 10000 25 5
python tools/batch_rate.py 2000 1000 10
