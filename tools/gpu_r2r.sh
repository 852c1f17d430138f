#!/bin/bash
set -x
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1400 -c 400 --csv --log-file gpurun_out/r2r_quad_launches.csv python tools/iter_breakdown.py quadratic 125000000 6 > gpurun_out/r2r_ncu.log 2>&1; echo "rc=$?"
tail -3 gpurun_out/r2r_ncu.log
