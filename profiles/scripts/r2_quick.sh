#!/bin/bash
# quick 1-GPU check after a plan-builder change: TC kernel tests, plan-build profile, bench line
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_kernels.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python profiles/graph_build_profile2.py 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_quick.json"))
print("C4 ms/step", d["ms_per_step"], "eager", d["ms_per_step_eager"], "e2e", d["e2e"]["ms_per_step"], "parity", d["parity"]["rel_dE"], d["parity"]["max_dF"])
PY
