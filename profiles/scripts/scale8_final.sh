#!/bin/bash
# final multi-GPU lines of the round (one gpurun --gpus 8 call): C4 domain decomposition at 8 and 4 GPUs, default bench flags
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
run() {
  name=$1; n=$2; shift 2
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/$name.json"))
    print("$name", "ms/step", round(d["ms_per_step"], 2), "eager", round(d.get("ms_per_step_eager", 0), 2), "e2e ms", round(d["e2e"]["ms_per_step"], 2), "value", round(d["value"]), "parity", d.get("parity", {}).get("rel_dE"))
except Exception as e:
    print("$name FAILED", e)
PY
}
run c4_n8_final 8 --steps 10 --warmup 3
run c4_n4_final 4 --steps 10 --warmup 3
