#!/bin/bash
# 2-GPU check after node-side / first-layer changes: DD == single GPU and DDP checks on hardware, C4 bench line at N=2
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_at_size.py -m gpu -q -k "domain_decomposition or ddp" 2>&1 | tail -2
grep dd_gpu_check gpurun_out/dd_gpu_check.log
grep ddp_gpu_check gpurun_out/ddp_gpu_check.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c4_n2_final.json 2> gpurun_out/c4_n2_final.err
python - <<PY
import json
d = json.load(open("gpurun_out/c4_n2_final.json"))
print("c4_n2", "ms/step", round(d["ms_per_step"], 2), "eager", d.get("ms_per_step_eager"), "e2e ms", round(d["e2e"]["ms_per_step"], 2), "value", round(d["value"]), d["config"].get("launch_mode"))
PY
