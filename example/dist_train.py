"""Distributed force-matching training on the B200-native hot path -- counterpart of the reference's
``example/dist_train.py`` (one process per GPU, NCCL gradient all-reduce by DDP, loss = 0.2 MSE(E) + 0.8 MSE(F) with the
forces from ``autograd.grad(..., create_graph=True)``), on a SYNTHETIC dataset (the reference downloads MD17; there is no
network here): periodic Li/Si/O cells (BASELINE configs[1]) labelled by a fixed random "teacher" network.

    python example/dist_train.py -w 2                      # spawns 2 ranks on this node (like the reference's -w)
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 example/dist_train.py

Differences from the reference script, all forced by the environment: no PyG / DGL (``hermnet_b200.DataLoader`` /
``torch.multiprocessing``), neighbour lists built on the GPU per batch (``DataLoader(rc=...)``) instead of per sample on
the CPU, micro-batched gradient accumulation for large graphs, HPNet by default (``--model HVNet`` is the reference's).
"""
import argparse
import os
import sys
from math import inf

import numpy as np
import torch
from torch import distributed as dist
from torch import nn
from torch.utils.data import Subset, distributed

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hermnet_b200 as H  # noqa: E402
from hermnet_b200 import parallel, synthetic  # noqa: E402
from hermnet_b200.utils import DistributedEvalSampler  # noqa: E402

ELEMS = ["Li", "Si", "O"]


def make_dataset(n_graphs, n_side, device, hidden=64, seed0=100):
    """``n_graphs`` cells of ``n_side^3`` sites (a = 2.3 A, jitter 0.1 A, Li:Si:O = 2:1:3) with energies / forces of a fixed
    random HVNet teacher (evaluated on the fused inference path)."""
    torch.manual_seed(4321)
    teacher = H.HVNet(elems=ELEMS, rc=5.0, num_layers=2, hidden_channels=hidden, num_rbf=32).to(device).eval()
    for p in teacher.parameters():
        p.requires_grad_(False)
    out = []
    for g in range(n_graphs):
        pos, Z, cell = synthetic.cubic_lattice(n_side, 2.3, ELEMS, (1 / 3, 1 / 6, 1 / 2), 0.10, seed0 + g)
        d = H.Data(pos=torch.from_numpy(pos).to(device).requires_grad_(True), atomic_number=torch.from_numpy(Z).to(device),
                   cell=torch.from_numpy(cell)[None].to(device))
        e = teacher(d)
        (gr,) = torch.autograd.grad(e.sum(), d.pos)
        out.append(H.Data(pos=torch.from_numpy(pos), atomic_number=torch.from_numpy(Z), cell=torch.from_numpy(cell)[None],
                          y=e.detach().cpu().reshape(1), forces=(-gr).detach().cpu()))
    return out


def main(world_size, rank, args):
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "1226")
    assert world_size <= torch.cuda.device_count() and 0 <= rank < world_size
    torch.cuda.set_device(rank)
    device = torch.device("cuda", rank)
    dist.init_process_group(backend="nccl", world_size=world_size, rank=rank, device_id=device)

    dataset = make_dataset(args.graphs, args.n_side, device)
    num_data = len(dataset)
    indices = np.random.RandomState(seed=123).permutation(num_data)
    n_val = max(world_size, int(0.1 * num_data))
    trainset, valset = Subset(dataset, indices[n_val:]), Subset(dataset, indices[:n_val])
    trn_mean = float(torch.stack([dataset[i].y for i in trainset.indices]).mean())
    n_atoms = args.n_side ** 3

    trn_sampler = distributed.DistributedSampler(trainset)
    trainloader = H.DataLoader(trainset, batch_size=args.batch_size, sampler=trn_sampler, collate_fn=list)
    val_sampler = DistributedEvalSampler(valset)
    valloader = H.DataLoader(valset, batch_size=args.micro, sampler=val_sampler, rc=5.0, device=device)

    torch.manual_seed(0)
    model = getattr(H, args.model)(elems=ELEMS, rc=5.0, num_layers=args.layers, hidden_channels=args.hidden, num_rbf=args.num_rbf)
    model = parallel.data_parallel(model.to(device), device_ids=[rank], output_device=rank)
    optimizer = torch.optim.Adam(model.parameters(), lr=3e-4)
    evaluation = nn.L1Loss(reduction="sum")
    best_log = inf
    for epoch in range(args.epochs):
        loss_sum = torch.zeros((), device=device)
        model.train()
        trainloader.sampler.set_epoch(epoch)
        for graphs in trainloader:                    # a list of Data: micro-batched inside the step
            loss_sum += parallel.force_matching_step_microbatched(model, graphs, optimizer, args.micro, 0.8, trn_mean, device)
        dist.all_reduce(loss_sum)
        val_e = torch.zeros((), device=device)
        val_f = torch.zeros((), device=device)
        model.eval()                                  # fused inference kernels, first-order backward
        for val_data in valloader:
            val_data.pos.requires_grad_(True)
            pred_e = model(val_data)
            val_e += evaluation(pred_e, val_data.y.reshape(-1) - trn_mean).detach()
            pred_f = -torch.autograd.grad(pred_e.sum(), val_data.pos)[0]
            val_f += evaluation(pred_f, val_data.forces)
        dist.all_reduce(val_e)
        dist.all_reduce(val_f)
        if rank == 0:
            mae_f = val_f.item() / len(valset) / n_atoms / 3
            print("Epoch #{:01d} | train loss {:.5f} | Val MAE_E: {:.4f} | Val MAE_F: {:.5f}.".format(
                epoch + 1, loss_sum.item() / max(1, len(trainloader)) / world_size, val_e.item() / len(valset), mae_f), flush=True)
            if best_log >= mae_f:
                best_log = mae_f
                torch.save(model.module.state_dict(), args.out)
    dist.destroy_process_group()


if __name__ == "__main__":
    parser = argparse.ArgumentParser(description="Distributed training (synthetic data)")
    parser.add_argument("-w", "--world_size", help="# of GPUs (omit under torchrun)", type=int, default=None)
    parser.add_argument("--model", default="HPNet", choices=["HVNet", "HPNet", "HTNet"])
    parser.add_argument("--graphs", type=int, default=24)
    parser.add_argument("--n-side", type=int, default=8)
    parser.add_argument("--batch-size", type=int, default=4, help="graphs per rank and step")
    parser.add_argument("--micro", type=int, default=2, help="graphs per micro-batch")
    parser.add_argument("--epochs", type=int, default=2)
    parser.add_argument("--layers", type=int, default=2)
    parser.add_argument("--hidden", type=int, default=64)
    parser.add_argument("--num-rbf", type=int, default=32)
    parser.add_argument("--out", default="best-model.pt")
    args = parser.parse_args()
    if "RANK" in os.environ:                          # launched by torchrun
        main(int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), args)
    else:
        import torch.multiprocessing as mp
        w = args.world_size or 1
        ctx = mp.get_context("spawn")
        procs = [ctx.Process(target=main, args=(w, r, args)) for r in range(w)]
        for p in procs:
            p.start()
        for p in procs:
            p.join()
