"""Multi-process paths on CPU: gloo backend, world_size 2 (and 4), kernel layer emulated (tests/emulator.py).

* domain decomposition: brick partition, local row CSR with ghost sources, per-layer halo exchange and its transposed
  backward (reverse force accumulation), energy / gradient all-reduce  == single-process result;
* data parallelism: DDP gradients of the force-matching step == single-process gradients on the concatenated batch.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import hermnet_b200 as H
from hermnet_b200 import parallel, synthetic
from tests import util


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _init(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests import emulator
    emulator.install_plain()
    torch.set_num_threads(2)


def _dd_worker(rank, world, port, kind, out, intensive=False):
    _init(rank, world, port)
    pos, Z, cell = synthetic.cubic_lattice(8, 2.3, ("Li", "Si", "O"), (1 / 3, 1 / 6, 1 / 2), 0.1, 5)
    cfg = dict(elems=["Li", "Si", "O"], rc=4.0, num_layers=3, hidden_channels=32, num_rbf=32)
    model, _ = util.make_model(kind, cfg, 9, intensive=intensive)
    p, z, c = torch.from_numpy(pos), torch.from_numpy(Z), torch.from_numpy(cell)[None]
    dd = parallel.DomainDecomposition(model, torch.device("cpu")).build(p, z, c)
    e, g = dd.energy_forces(p)
    assert dd.n_ghost_max > 0 and dd.graph.n_ghost < len(Z) - dd.n_owned     # a real halo, not the whole box
    if rank == 0:
        torch.save({"e": e, "g": g, "edges": dd.global_edges, "ghost": dd.n_ghost_max}, out)
    dist.destroy_process_group()


@pytest.mark.parametrize("world,kind,intensive", [(2, "HVNet", False), (4, "HVNet", False), (2, "HPNet", False),
                                                  (2, "HTNet", False), (2, "HVNet", True)])
def test_domain_decomposition_equals_single_process(emu, tmp_path, world, kind, intensive):
    out = str(tmp_path / "dd.pt")
    mp.spawn(_dd_worker, args=(world, _free_port(), kind, out, intensive), nprocs=world, join=True)
    got = torch.load(out)
    pos, Z, cell = synthetic.cubic_lattice(8, 2.3, ("Li", "Si", "O"), (1 / 3, 1 / 6, 1 / 2), 0.1, 5)
    cfg = dict(elems=["Li", "Si", "O"], rc=4.0, num_layers=3, hidden_channels=32, num_rbf=32)
    model, _ = util.make_model(kind, cfg, 9, intensive=intensive)
    d = H.Data(pos=torch.from_numpy(pos).requires_grad_(True), atomic_number=torch.from_numpy(Z),
               cell=torch.from_numpy(cell)[None])
    e = model(d)
    (g,) = torch.autograd.grad(e.sum(), d.pos)
    assert got["edges"] == d.graph.n_edges
    assert util.rel_err(got["e"], e.detach()) < 1e-5
    assert float((got["g"] - g).abs().max()) < 1e-4 * max(1.0, float(g.abs().max()))


def _make_batch(seeds):
    parts = []
    for s in seeds:
        pos, Z, cell = synthetic.cubic_lattice(3, 2.3, ("Li", "Si", "O"), (1 / 3, 1 / 6, 1 / 2), 0.1, s)
        rng = np.random.default_rng(s)
        d = H.Data(pos=torch.from_numpy(pos), atomic_number=torch.from_numpy(Z), cell=torch.from_numpy(cell)[None],
                   y=torch.tensor([float(rng.normal())]), forces=torch.from_numpy(rng.normal(size=pos.shape).astype(np.float32)))
        parts.append(H.transform(d, 4.0))
    return H.Batch.from_data_list(parts)


def _ddp_worker(rank, world, port, out):
    _init(rank, world, port)
    cfg = dict(elems=["Li", "Si", "O"], rc=4.0, num_layers=2, hidden_channels=32, num_rbf=16)
    model, _ = util.make_model("HPNet", cfg, 4)
    model.train()
    ddp = parallel.data_parallel(model)
    opt = torch.optim.SGD(ddp.parameters(), lr=0.0)
    seeds = [100, 101, 102, 103]
    per = len(seeds) // world
    batch = _make_batch(seeds[rank * per:(rank + 1) * per])
    parallel.force_matching_step(ddp, batch, opt)
    if rank == 0:
        torch.save({k: (p.grad.clone() if p.grad is not None else None) for k, p in model.named_parameters()}, out)
    dist.destroy_process_group()


def test_ddp_force_matching_gradients_equal_single_process(emu, tmp_path):
    out = str(tmp_path / "ddp.pt")
    mp.spawn(_ddp_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    cfg = dict(elems=["Li", "Si", "O"], rc=4.0, num_layers=2, hidden_channels=32, num_rbf=16)
    model, _ = util.make_model("HPNet", cfg, 4)
    model.train()
    opt = torch.optim.SGD(model.parameters(), lr=0.0)
    # DDP averages the per-rank mean losses; with equal graph and atom counts per rank that is the global mean
    parallel.force_matching_step(model, _make_batch([100, 101, 102, 103]), opt)
    n = 0
    for k, p in model.named_parameters():
        if p.grad is None:
            continue
        assert got[k] is not None, k
        assert float((got[k] - p.grad).abs().max()) < 1e-4 * (float(p.grad.abs().max()) + 1e-6) + 1e-7, k
        n += 1
    assert n > 10


def test_grid_factorisation():
    assert parallel._grid(1) == (1, 1, 1) and parallel._grid(2) == (2, 1, 1)
    assert parallel._grid(4) == (2, 2, 1) and parallel._grid(8) == (2, 2, 2)
