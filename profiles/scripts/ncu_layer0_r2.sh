#!/bin/bash
# ncu --set full of the two first-layer basis kernels at C4 size (small report: two launches, no source import)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 120 python profiles/layer0_time.py 10 2>&1 | tail -1 | tee gpurun_out/layer0_time.log
timeout 200 ncu --set full --clock-control none -k regex:layer0_basis -c 2 -f -o gpurun_out/prof_r2_layer0 python profiles/layer0_time.py 1 > gpurun_out/prof_r2_layer0.log 2>&1
tail -1 gpurun_out/prof_r2_layer0.log
ls -la gpurun_out/
