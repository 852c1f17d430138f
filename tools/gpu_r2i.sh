#!/bin/bash
# round 2, call I: f32 rare-path cases on the GPU, steady per-kernel times of config 5, ncu --set full of the new <float,20> passes
set -x
python tools/_explore_f32_ascent.py 2>&1 | tail -6
python tools/config5_rate.py > gpurun_out/r2i_c5.log 2>&1; tail -4 gpurun_out/r2i_c5.log
python -m pytest tests/test_gpu_batch.py tests/test_gpu_rare_paths.py tests/test_gpu_configs.py -m gpu -q 2>&1 | tail -8
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_update_classify|k_formk_cmprlb|k_subsm_lsinit" --launch-skip 104 --launch-count 4 -f -o gpurun_out/r2i_c5_full python tools/config5_rate.py > gpurun_out/r2i_c5_ncu.log 2>&1; echo "ncu rc=$?"
