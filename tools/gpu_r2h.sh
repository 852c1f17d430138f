#!/bin/bash
# round 2, call H: REAL32 shape VEC=2/UNROLL=8 (all 256 threads on every staged sub-tile): tests, digests, config 5 rate
set -x
python tools/trace_digest.py gpurun_out/r2h_digest.json > gpurun_out/r2h_digest.log 2>&1; python - <<PY
import json
a=json.load(open("tests/golden/gpu_trace_digest.json")); b=json.load(open("gpurun_out/r2h_digest.json"))
print(" ".join(("SAME" if a[k]["sha256"]==b.get(k,{}).get("sha256") else "DIFF:"+k) for k in a))
PY
python tools/config5_rate.py > gpurun_out/r2h_c5.log 2>&1; tail -12 gpurun_out/r2h_c5.log
python -m pytest tests -m gpu -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2h_pytest.log
