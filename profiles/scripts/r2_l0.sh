#!/bin/bash
# layer-0 aggregation check: kernel tests, bench line
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_fused_node.py -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --no-cpu-baseline --e2e-steps 3 > gpurun_out/bench_l0.json 2> gpurun_out/bench_l0.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_l0.json"))
print("C4 ms/step", d["ms_per_step"], "eager", d["ms_per_step_eager"], "e2e", d["e2e"]["ms_per_step"], "parity", d["parity"]["rel_dE"], d["parity"]["max_dF"])
print({k: round(v["avg_ms"] * v["launches"] / d["steps"], 2) for k, v in d["kernels"].items()})
PY
