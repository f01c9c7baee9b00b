#!/bin/bash
# final 1-GPU evidence of the round: full GPU test suite, bench line (+ reference arm), full-step ncu launch list, C3 MD line
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3) > gpurun_out/r2_pytest_final.log
tail -2 gpurun_out/r2_pytest_final.log
timeout 600 python bench.py > gpurun_out/bench_r2_final_c4_n1.json 2> gpurun_out/bench_r2_final_c4_n1.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_r2_reference_arm.json 2> gpurun_out/bench_r2_reference_arm.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 2 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-parity --cuda-graph 0 > gpurun_out/launches_r2.log 2>&1
timeout 600 python bench.py --workload C3 --steps 10 --warmup 3 --e2e-steps 5 > gpurun_out/bench_r2_c3_n1.json 2> gpurun_out/bench_r2_c3_n1.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_r2_final_c4_n1.json"))
print("C4 ms/step", d["ms_per_step"], "eager", d["ms_per_step_eager"], "e2e", d["e2e"]["ms_per_step"], "parity", d["parity"]["rel_dE"], d["parity"]["max_dF"])
print("roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3), {k: round(v.get("frac_of_hbm_peak", 0), 3) for k, v in d["kernels"].items() if "edge" in k and "geom" not in k})
print({k: round(v["avg_ms"] * v["launches"] / d["steps"], 2) for k, v in d["kernels"].items()})
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["cpu_baseline"].get("c1_192_atoms"))
r = json.load(open("gpurun_out/bench_r2_reference_arm.json"))
print("reference arm", r["value"], r["unit"], r.get("cpu_baseline", {}).get("cores"))
c = json.load(open("gpurun_out/bench_r2_c3_n1.json"))
print("C3 ms/step", c["ms_per_step"], c.get("neighbour_list"), "e2e", c["e2e"]["ms_per_step"])
PY
