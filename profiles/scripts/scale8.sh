#!/bin/bash
# 8-GPU evidence run (one gpurun --gpus 8 call): C4 domain decomposition at 8 (peer / nccl halo) and 4 GPUs, C2 DDP training at
# 8 and 4 GPUs, and the hardware DD / DDP checks.  Everything lands in gpurun_out/.
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
run() {  # name gpus extra-env... -- bench args
  name=$1; n=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/$name.json"))
    print("$name", "ms/step", round(d["ms_per_step"], 2), "e2e ms", round(d["e2e"]["ms_per_step"], 2), "value", round(d["value"]))
except Exception as e:
    print("$name FAILED", e)
PY
}
HERMNET_B200_HALO=peer run c4_n8_peer 8 --steps 10 --warmup 3
HERMNET_B200_HALO=nccl run c4_n8_nccl 8 --steps 10 --warmup 3
HERMNET_B200_HALO=peer run c4_n4_peer 4 --steps 10 --warmup 3
run c2_n8 8 --workload C2 --steps 5 --warmup 3 --e2e-steps 2
run c2_n4 4 --workload C2 --steps 5 --warmup 3 --e2e-steps 2
timeout 600 python -m pytest tests/test_parity_at_size.py -m gpu -q -k "domain_decomposition or ddp" 2>&1 | tail -2
cat gpurun_out/dd_gpu_check.log | grep dd_gpu_check
cat gpurun_out/ddp_gpu_check.log | grep ddp_gpu_check
