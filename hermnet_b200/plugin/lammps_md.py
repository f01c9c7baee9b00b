"""LAMMPS ``fix client/md`` server loop -- the message protocol of ``/root/reference/plugin/lmp_interface/lmp_calc.py:136-240``
(message / field ids of LAMMPS' ``FixClientMD``), written against the CSlib call surface the reference uses (``recv``,
``unpack_string``, ``unpack_int``, ``unpack``, ``send``, ``pack``, ``pack_double``).  ``cslib`` (the LAMMPS messaging
library) is not installed in this image: ``serve_md`` takes any object with that surface -- the real ``CSlib`` when it is
importable, ``LoopbackCS`` (an in-process stand-in used by the tests and for embedding) otherwise.  No arithmetic lives
here: energies / forces / virial come from the B200 hot path through ``calculator``."""
from __future__ import annotations

from collections import deque
from typing import List, Optional, Sequence

import numpy as np
import torch

from ..symbols import atomic_numbers
from ..utils import virial_calc
from .calculator import build_graph

# enums matching FixClientMD (lmp_calc.py:136-138)
SETUP, STEP = 1, 2
DIM, PERIODICITY, ORIGIN, BOX, NATOMS, NTYPES, TYPES, COORDS, UNITS, CHARGE = range(1, 11)
FORCES, ENERGY, VIRIAL, ERROR = range(1, 5)


def calculator(data, model, trn_mean, device='cuda', pbc=True, units='metal', ensemble='NVT', verlet=None):
    """lmp_calc.py:36-85: ``(energy, forces flattened [3N], virial [6])``; the virial vector follows LAMMPS' order
    xx yy zz xy xz yz (the reference fills slot 0 with ``virial[0,1]``, a typo)."""
    dev = torch.device(device)
    data = data.to(dev)
    if verlet is not None and pbc and verlet.skin > 0.0:
        data.graph = verlet.get(data.pos, data.atomic_number, data.get("cell"))
    data.pos.requires_grad_(True)
    npt = ensemble.lower() == 'npt'
    if npt and pbc:
        data.cell.requires_grad_(True)
    model.eval()
    energy = model(data) + trn_mean
    forces = -torch.autograd.grad(energy.sum(), data.pos, retain_graph=npt and pbc)[0]
    if npt:
        v = virial_calc(cell=data.cell if pbc else None, pos=data.pos.detach(), forces=forces, energy=energy, units=units,
                        pbc=pbc).detach().cpu().numpy()
        virial = np.array([v[0, 0], v[1, 1], v[2, 2], v[0, 1], v[0, 2], v[1, 2]])
    else:
        virial = np.zeros(6)
    return energy.detach().cpu().item(), forces.detach().cpu().view(-1).numpy(), virial


def serve_md(cs, model, elems: Sequence[str], rc: float, trn_mean: float = 0.0, device: str = 'cuda', periodic: bool = True,
             units: str = 'metal', ensemble: str = 'NVT', skin: float = 0.0) -> int:
    """The server loop of lmp_calc.py:140-240.  Returns the number of SETUP / STEP messages answered; protocol errors raise
    ``RuntimeError`` (the reference prints and ``sys.exit(1)``; the CLI wrapper maps the exception to that)."""
    from .md import VerletGraph
    msgID, nfield, fieldID, fieldtype, fieldlen = cs.recv()
    if msgID != 0:
        raise RuntimeError('Error: Bad initial client/server handshake')
    if cs.unpack_string(1) != b'md':
        raise RuntimeError('Error: Mismatch in client/server protocol')
    cs.send(0, 0)
    verlet = VerletGraph(model, skin)
    dim = natoms = ntypes = None
    periodicity = origin = box = types = coords = None
    answered = 0
    while True:
        msgID, nfield, fieldID, fieldtype, fieldlen = cs.recv()
        if msgID < 0:
            break
        if msgID == SETUP:      # required: DIM, PERIODICITY, ORIGIN, BOX, NATOMS, NTYPES, TYPES, COORDS; optional UNITS, CHARGE
            for field in fieldID:
                if field == DIM:
                    dim = cs.unpack_int(DIM)
                elif field == PERIODICITY:
                    periodicity = cs.unpack(PERIODICITY, 1)
                elif field == ORIGIN:
                    origin = cs.unpack(ORIGIN, 1)
                elif field == BOX:
                    box = cs.unpack(BOX, 1)
                elif field == NATOMS:
                    natoms = cs.unpack_int(NATOMS)
                elif field == NTYPES:
                    ntypes = cs.unpack_int(NTYPES)
                elif field == TYPES:
                    types = cs.unpack(TYPES, 1)
                elif field == COORDS:
                    coords = cs.unpack(COORDS, 1)
                elif field == UNITS:
                    sent = cs.unpack_string(UNITS)
                    if sent is not None and sent.decode().strip('\\0') not in ('', units):
                        raise RuntimeError(f'Error: LAMMPS runs units {sent!r}, the server was started with {units!r}')
            if None in (dim, natoms, ntypes) or box is None or types is None or coords is None:
                raise RuntimeError('Error: SETUP message is missing required fields')
            if ntypes > len(elems):
                raise RuntimeError('Error: more LAMMPS atom types than elements given to the server')
        elif msgID == STEP:     # required: COORDS; optional: ORIGIN, BOX
            for field in fieldID:
                if field == COORDS:
                    coords = cs.unpack(COORDS, 1)
                elif field == ORIGIN:
                    origin = cs.unpack(ORIGIN, 1)
                elif field == BOX:
                    box = cs.unpack(BOX, 1)
            if natoms is None:
                raise RuntimeError('Error: STEP received before SETUP')
        else:
            raise RuntimeError('Error: HermNet wrapper received unrecognized message')
        cell = np.array(box, dtype=np.float64).reshape(3, 3)
        elements = np.array(types).astype(np.int64)
        z = np.zeros_like(elements)
        for i, el in enumerate(elems):       # LAMMPS type i+1 -> element i (order of the data file)
            z[elements == (i + 1)] = atomic_numbers[el]
        pos = np.array(coords, dtype=np.float64).reshape(natoms, 3)
        data = build_graph(cell=cell if periodic else None, elements=z, pos=pos, rc=rc, device=device)
        energy, forces, virial = calculator(data, model, trn_mean, device, periodic, units, ensemble, verlet)
        cs.send(msgID, 3)
        cs.pack(FORCES, 4, 3 * natoms, forces.tolist())
        cs.pack_double(ENERGY, energy)
        cs.pack(VIRIAL, 4, 6, virial.tolist())
        answered += 1
    cs.send(0, 0)
    return answered


class LoopbackCS:
    """In-process stand-in for a CSlib endpoint pair: the test / embedding side queues messages with ``client_send`` and
    reads the replies from ``replies``; the server side is driven through the CSlib call surface."""

    def __init__(self):
        self.inbox, self.replies = deque(), []
        self._fields, self._out = {}, None

    # ---- client side
    def client_send(self, msgID: int, fields: Optional[List[tuple]] = None):
        """``fields``: list of ``(fieldID, value)``; value = int, bytes or a sequence of numbers."""
        self.inbox.append((msgID, list(fields or [])))

    # ---- CSlib surface used by the server
    def recv(self):
        if not self.inbox:
            return -1, 0, [], [], []
        msgID, fields = self.inbox.popleft()
        self._fields = {fid: v for fid, v in fields}
        ids = [fid for fid, _ in fields]
        return msgID, len(ids), ids, [0] * len(ids), [len(v) if hasattr(v, "__len__") else 1 for _, v in fields]

    def unpack_string(self, fid):
        return self._fields.get(fid)

    def unpack_int(self, fid):
        return int(self._fields[fid])

    def unpack(self, fid, _tflag):
        return list(self._fields[fid])

    def send(self, msgID, nfield):
        self._out = {"msgID": msgID, "nfield": nfield, "fields": {}}
        self.replies.append(self._out)

    def pack(self, fid, _ftype, n, values):
        assert len(values) == n
        self._out["fields"][fid] = list(values)

    def pack_double(self, fid, value):
        self._out["fields"][fid] = float(value)
