"""Run under torchrun on N GPUs (NCCL): DDP force-matching gradients (example/dist_train.py:84-99, micro-batched) ==
single-process gradients on the whole batch (rank 0 checks).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/ddp_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hermnet_b200 as H  # noqa: E402
from hermnet_b200 import parallel, synthetic  # noqa: E402


def graphs(seeds, dev):
    out = []
    for s in seeds:
        pos, Z, cell = synthetic.cubic_lattice(6, 2.3, ("Li", "Si", "O"), (1 / 3, 1 / 6, 1 / 2), 0.1, s)
        rng = np.random.default_rng(s)
        out.append(H.Data(pos=torch.from_numpy(pos).to(dev), atomic_number=torch.from_numpy(Z).to(dev),
                          cell=torch.from_numpy(cell)[None].to(dev), y=torch.tensor([float(rng.normal())], device=dev),
                          forces=torch.from_numpy(rng.normal(size=pos.shape).astype(np.float32)).to(dev)))
    return out


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    ok = True
    for kind in ("HPNet", "HVNet"):
        torch.manual_seed(3)
        model = getattr(H, kind)(elems=["Li", "Si", "O"], rc=5.0, num_layers=2, hidden_channels=64, num_rbf=32).to(dev).train()
        ref = getattr(H, kind)(elems=["Li", "Si", "O"], rc=5.0, num_layers=2, hidden_channels=64, num_rbf=32).to(dev).train()
        ref.load_state_dict(model.state_dict())
        ddp = parallel.data_parallel(model, device_ids=[local], output_device=local)
        opt = torch.optim.SGD(ddp.parameters(), lr=0.0)
        seeds = list(range(100, 100 + 4 * world))
        parallel.force_matching_step_microbatched(ddp, graphs(seeds[rank::world], dev), opt, micro=2)
        if rank == 0:
            opt1 = torch.optim.SGD(ref.parameters(), lr=0.0)
            parallel.force_matching_step_microbatched(ref, graphs(seeds, dev), opt1, micro=4)
            worst, n = 0.0, 0
            for (k, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
                if q.grad is None:
                    continue
                assert p.grad is not None, k
                worst = max(worst, float((p.grad - q.grad).abs().max() / (q.grad.abs().max() + 1e-6)))
                n += 1
            good = worst < 2e-4 and n > 10
            ok &= good
            print(f"[ddp_gpu_check] {kind} world={world}: {n} parameter tensors, worst relative gradient difference {worst:.2e} "
                  f"{'OK' if good else 'FAIL'}", flush=True)
        dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
