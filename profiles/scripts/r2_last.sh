#!/bin/bash
# last check of the round: full GPU suite, plan-build profile, bench line
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -2) > gpurun_out/r2_pytest_last.log
tail -1 gpurun_out/r2_pytest_last.log
timeout 200 python profiles/graph_build_profile2.py 2>&1 | tail -2 | head -1
timeout 600 python bench.py > gpurun_out/bench_r2_last_c4_n1.json 2> gpurun_out/bench_r2_last_c4_n1.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_r2_last_c4_n1.json"))
print("C4 ms/step", d["ms_per_step"], "eager", d["ms_per_step_eager"], "e2e", d["e2e"]["ms_per_step"], "parity", d["parity"]["rel_dE"], d["parity"]["max_dF"])
PY
