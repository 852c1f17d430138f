#!/bin/bash
set -x
python tools/iter_breakdown.py quadratic 125000000 14 2>&1 | tail -16
python tools/iter_breakdown.py rosenbrock 100000000 6 2>&1 | tail -8
