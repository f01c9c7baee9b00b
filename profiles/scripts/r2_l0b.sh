#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_fused_node.py -m gpu -x -q -k "layer0" 2>&1 | tail -1
timeout 60 python profiles/layer0_time.py 10 2>&1 | tail -1 | tee gpurun_out/layer0_time.log
timeout 75 python bench.py --no-cpu-baseline --e2e-steps 2 > gpurun_out/bench_l0b.json 2> gpurun_out/bench_l0b.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_l0b.json"))
print("C4 ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "parity", d["parity"]["rel_dE"], d["parity"]["max_dF"])
print({k: round(v["avg_ms"] * v["launches"] / d["steps"], 2) for k, v in d["kernels"].items() if "layer0" in k})
PY
