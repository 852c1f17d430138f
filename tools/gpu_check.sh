python tools/trace_digest.py gpurun_out/cur_digest.json > gpurun_out/cur_digest.log 2>&1; python - <<EOF
import json
a=json.load(open("tests/golden/gpu_trace_digest.json")); b=json.load(open("gpurun_out/cur_digest.json"))
print(" ".join(("SAME" if a[k]["sha256"]==b.get(k,{}).get("sha256") else "DIFF:"+k) for k in a))
EOF
python bench.py --no-cpu --no-e2e > gpurun_out/cur_bench.json 2> gpurun_out/cur_bench.err; tail -3 gpurun_out/cur_bench.err; python -c "
import json; d=json.load(open('gpurun_out/cur_bench.json')); print(d['value'], d['ms_per_step'], d['iteration_roofline']['frac_of_peak']); 
print(' '.join('%s=%.3f' % (k, v['ms_per_call']) for k,v in d['kernel_families'].items() if v['ms_per_call']>0.05))"
