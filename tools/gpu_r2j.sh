#!/bin/bash
set -x
python tools/trace_digest.py gpurun_out/r2j_digest.json > gpurun_out/r2j_digest.log 2>&1; python - <<PY
import json
a=json.load(open("gpurun_out/r2h_digest.json")) if __import__("os").path.exists("gpurun_out/r2h_digest.json") else json.load(open("tests/golden/gpu_trace_digest.json")); b=json.load(open("gpurun_out/r2j_digest.json"))
print(" ".join(("SAME" if a[k]["sha256"]==b.get(k,{}).get("sha256") else "DIFF:"+k) for k in a))
PY
python tools/config5_rate.py > gpurun_out/r2j_c5.log 2>&1; tail -14 gpurun_out/r2j_c5.log
