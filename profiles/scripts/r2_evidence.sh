#!/bin/bash
# 1-GPU evidence of a round: GPU tests, bench line, ncu launch list of the same bench command, ncu --set full of the TC edge kernels
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3) > gpurun_out/r2_pytest.log
timeout 600 python bench.py > gpurun_out/bench_r2_c4_n1.json 2> gpurun_out/bench_r2_c4_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 2 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-parity > gpurun_out/launches_r2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:edge_tc -s 9 -c 8 -f -o gpurun_out/prof_r2_edge_tc \
    python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-parity > gpurun_out/prof_r2_edge_tc.log 2>&1
tail -2 gpurun_out/r2_pytest.log
python - <<PY
import json
d = json.load(open("gpurun_out/bench_r2_c4_n1.json"))
print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "parity", d.get("parity", {}).get("rel_dE"), d.get("parity", {}).get("max_dF"))
print({k: (round(v["avg_ms"], 2), v["launches"], round(v.get("frac_of_hbm_peak", 0), 3)) for k, v in d["kernels"].items()})
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
PY
