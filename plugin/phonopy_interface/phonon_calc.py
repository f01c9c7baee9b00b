"""Finite-displacement force sets (reference: ``/root/reference/plugin/phonopy_interface/phonopy_calc.py:36-44`` loops
``atoms.get_forces()`` over displaced supercells on the host).  SURVEY 8(f) rank 3: all displaced supercells are
evaluated in ONE batched forward on the GPU (graphs of a batch never share edges)."""
import numpy as np
import torch

from hermnet_b200.data import Batch, Data


def batched_force_sets(model, numbers, cell, supercells, device='cuda', chunk=64):
    """``supercells``: iterable of ``[N,3]`` position arrays (the displaced supercells).  Returns ``[S,N,3]`` forces."""
    dev = torch.device(device)
    Z = torch.as_tensor(np.asarray(numbers)).long()
    c = torch.as_tensor(np.asarray(cell), dtype=torch.float32).reshape(1, 3, 3)
    out = []
    sc = list(supercells)
    for i in range(0, len(sc), chunk):
        parts = [Data(pos=torch.as_tensor(np.asarray(p), dtype=torch.float32), atomic_number=Z, cell=c) for p in sc[i:i + chunk]]
        b = Batch.from_data_list(parts).to(dev)
        b.pos.requires_grad_(True)
        e = model(b)
        (g,) = torch.autograd.grad(e.sum(), b.pos)
        out.append((-g).detach().cpu().numpy().reshape(len(parts), -1, 3))
    return np.concatenate(out, 0)
