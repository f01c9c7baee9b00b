#!/bin/bash
# ncu --set full of the round-2 node-side / first-layer kernels inside one bench step
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-parity --cuda-graph 0"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"layer0_basis|layernorm|readout" -c 12 -f -o gpurun_out/prof_r2_node $B > gpurun_out/prof_r2_node.log 2>&1
tail -1 gpurun_out/prof_r2_node.log
timeout 500 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -c 28 -f -o gpurun_out/prof_r2_gemm $B > gpurun_out/prof_r2_gemm.log 2>&1
tail -1 gpurun_out/prof_r2_gemm.log
