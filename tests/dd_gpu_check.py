"""Run under torchrun on N GPUs: domain-decomposed energy/forces == single-GPU result (rank 0 checks).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/dd_gpu_check.py [scale]
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hermnet_b200 as H  # noqa: E402
from hermnet_b200 import parallel, synthetic  # noqa: E402


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.3
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    ok = True
    for kind, F in (("HVNet", 128), ("HPNet", 64), ("HTNet", 64)):
        (pos, Z, cell), cfg = synthetic.config("C4", scale)
        cfg = dict(cfg, hidden_channels=F)
        cfg.pop("kind")
        torch.manual_seed(7)
        model = getattr(H, kind)(**cfg).to(dev).eval()
        for p in model.parameters():
            p.requires_grad_(False)
        for p in model.parameters():           # identical weights on every rank
            dist.broadcast(p.data, 0)
        p_, z_, c_ = torch.from_numpy(pos).to(dev), torch.from_numpy(Z).to(dev), torch.from_numpy(cell)[None].to(dev)
        dd = parallel.DomainDecomposition(model, dev).build(p_, z_, c_)
        e, g = dd.energy_forces(p_)
        if rank == 0:
            d = H.Data(pos=p_.clone().requires_grad_(True), atomic_number=z_, cell=c_)
            e1 = model(d)
            (g1,) = torch.autograd.grad(e1.sum(), d.pos)
            de = float((e - e1.detach()).abs() / e1.detach().abs())
            dg = float((g - g1).abs().max())
            fs = float(g1.abs().max())
            good = de < 1e-5 and dg < 1e-4 * max(1.0, fs)
            ok &= good
            print(f"[dd_gpu_check] {kind} N={len(Z)} world={world} grid={dd.grid} owned<= {dd.n_owned_max} ghosts<= {dd.n_ghost_max} "
                  f"edges={dd.global_edges} halo={dd.halo.transport}{'' if dd.peer_error is None else ' (' + dd.peer_error[:80] + ')'}: rel dE={de:.2e} max|dG|={dg:.2e} (|G|max {fs:.2f}) {'OK' if good else 'FAIL'}", flush=True)
        dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
