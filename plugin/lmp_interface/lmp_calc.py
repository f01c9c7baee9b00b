"""LAMMPS ``fix client/md`` server entry points (reference: ``/root/reference/plugin/lmp_interface/lmp_calc.py``).

The CSlib message loop (lmp_calc.py:136-240) is a wire protocol over a C++ library that is not in this image and is
out of scope; the two functions a server loop calls are kept signature-compatible and functional:
``build_graph(cell, elements, pos, rc)`` (lmp_calc.py:16-33) and ``calculator(model, data, trn_mean, device, pbc,
ensemble)`` (lmp_calc.py:36-85) returning ``(energy, forces[N,3], virial[6])``."""
import numpy as np
import torch

from hermnet_b200.plugin.calculator import build_graph  # noqa: F401
from hermnet_b200.utils import virial_calc


def calculator(model, data, trn_mean, device='cuda', pbc=True, ensemble='NVT'):
    dev = torch.device(device)
    data = data.to(dev)
    data.pos.requires_grad_(True)
    npt = ensemble.lower() == 'npt'
    if npt and pbc:
        data.cell.requires_grad_(True)
    model.eval()
    energy = model(data) + trn_mean
    forces = -torch.autograd.grad(energy.sum(), data.pos, retain_graph=npt and pbc)[0]
    if npt:
        v = virial_calc(cell=data.cell if pbc else None, pos=data.pos.detach(), forces=forces, energy=energy,
                        units='metal', pbc=pbc).detach().cpu().numpy()
        virial = np.array([v[0, 0], v[1, 1], v[2, 2], v[0, 1], v[0, 2], v[1, 2]])   # LAMMPS order xx yy zz xy xz yz
    else:
        virial = np.zeros(6)
    return energy.detach().cpu().item(), forces.detach().cpu().numpy().reshape(-1, 3), virial
