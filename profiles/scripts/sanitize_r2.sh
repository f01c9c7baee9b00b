#!/bin/bash
# compute-sanitizer over the round-2 node-side / first-layer kernels (small cases) + a bench line of the same build
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 280 compute-sanitizer --tool memcheck --log-file gpurun_out/sanitize_memcheck.log python -m pytest tests/test_fused_node.py -m gpu -q -k "layer0_basis_kernels or layernorm or readout" -x 2>&1 | tail -1
tail -1 gpurun_out/sanitize_memcheck.log
timeout 200 compute-sanitizer --tool racecheck --log-file gpurun_out/sanitize_racecheck.log python -m pytest tests/test_fused_node.py -m gpu -q -k "layer0_basis_kernels or layernorm" -x 2>&1 | tail -1
tail -1 gpurun_out/sanitize_racecheck.log
timeout 600 python bench.py --no-cpu-baseline --e2e-steps 3 > gpurun_out/bench_l0.json 2> gpurun_out/bench_l0.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_l0.json"))
print("C4 ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "parity", d["parity"]["rel_dE"], d["parity"]["max_dF"])
print({k: round(v["avg_ms"] * v["launches"] / d["steps"], 2) for k, v in d["kernels"].items() if "layer0" in k})
PY
