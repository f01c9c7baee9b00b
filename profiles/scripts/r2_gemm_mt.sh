#!/bin/bash
# GEMM with 256-row CTA tiles (MT = 2) vs 128-row tiles (A/B library): tests, per-shape timing, bench line
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_fused_node.py -m gpu -x -q 2>&1 | tail -3
echo "=== MT=2" | tee gpurun_out/gemm_mt.log
timeout 200 python profiles/gemm_time.py 10 2>&1 | tail -12 | tee -a gpurun_out/gemm_mt.log
echo "=== MT=1" | tee -a gpurun_out/gemm_mt.log
HERMNET_B200_LIB=$PWD/hermnet_b200/lib/libhn_gemm_mt1.so timeout 200 python profiles/gemm_time.py 10 2>&1 | tail -12 | tee -a gpurun_out/gemm_mt.log
timeout 600 python bench.py --no-cpu-baseline --e2e-steps 3 > gpurun_out/bench_mt2.json 2> gpurun_out/bench_mt2.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_mt2.json"))
print("C4 ms/step", d["ms_per_step"], "eager", d["ms_per_step_eager"], "e2e", d["e2e"]["ms_per_step"], "parity", d["parity"]["rel_dE"], d["parity"]["max_dF"])
print({k: round(v["avg_ms"] * v["launches"] / d["steps"], 2) for k, v in d["kernels"].items()})
PY
