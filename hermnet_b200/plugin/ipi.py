"""i-PI socket protocol, driver (client) side -- what the reference delegates to ``ase.calculators.socketio.SocketClient``
(plugin/i-pi_interface/ipi_calc.py:5-18); ASE is not installed here, so the protocol itself is implemented (pure Python,
no arithmetic on the hot path -- the forces come from ``NNCalculator``).

Wire format (i-PI "driver" protocol): 12-byte ASCII headers padded with blanks.  The server sends ``STATUS`` (reply
``NEEDINIT`` once, then ``READY`` / ``HAVEDATA``), ``INIT`` (int32 bead index, int32 n, n bytes), ``POSDATA`` (cell 9 f64
and inverse cell 9 f64, both TRANSPOSED, int32 natoms, 3N f64 positions -- atomic units: Bohr), ``GETFORCE`` (reply
``FORCEREADY``, f64 energy [Hartree], int32 natoms, 3N f64 forces [Hartree/Bohr], 9 f64 virial [Hartree], int32 n, n bytes
of extra text) and ``EXIT``.
"""
from __future__ import annotations

import socket
from typing import Optional

import numpy as np

BOHR = 0.52917721067          # Angstrom
HARTREE = 27.211386245988     # eV
HEADER = 12


def _pad(msg: str) -> bytes:
    return msg.encode("ascii").ljust(HEADER)


def _recv_exact(sock, n: int) -> bytes:
    buf = bytearray()
    while len(buf) < n:
        chunk = sock.recv(n - len(buf))
        if not chunk:
            raise ConnectionError("i-PI socket closed")
        buf += chunk
    return bytes(buf)


def _recv_array(sock, dtype, count):
    return np.frombuffer(_recv_exact(sock, np.dtype(dtype).itemsize * count), dtype=dtype).copy()


def read_poscar(path: str):
    """Minimal VASP-5 POSCAR reader (the reference reads the initial structure with ``ase.io.vasp.read_vasp``,
    ipi_calc.py:8): returns ``(symbols list, positions [N,3] Angstrom, cell [3,3])``."""
    with open(path) as fh:
        lines = [ln.strip() for ln in fh if ln.strip()]
    scale = float(lines[1])
    cell = np.array([[float(v) for v in lines[i].split()[:3]] for i in (2, 3, 4)]) * scale
    names = lines[5].split()
    counts = [int(v) for v in lines[6].split()]
    i = 7
    if lines[i][0].lower() == "s":       # selective dynamics
        i += 1
    direct = lines[i][0].lower() == "d"
    n = sum(counts)
    xyz = np.array([[float(v) for v in lines[i + 1 + k].split()[:3]] for k in range(n)])
    pos = xyz @ cell if direct else xyz * scale
    symbols = [s for s, c in zip(names, counts) for _ in range(c)]
    return symbols, pos, cell


class IPIClient:
    """Driver loop: receives positions / cell from an i-PI server, answers with energy, forces and virial computed by
    ``calc`` (any object with ``calculate(atoms)`` and ``results``, e.g. ``NNCalculator``) on ``atoms`` (``positions``,
    ``cell``, ``pbc``, ``get_chemical_symbols()`` -- ``md.SimpleAtoms`` or an ASE ``Atoms``)."""

    def __init__(self, host: str = "localhost", port: Optional[int] = None, unixsocket: Optional[str] = None, timeout: float = 600.0):
        if unixsocket is not None:
            self.sock = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            self.sock.connect(unixsocket if unixsocket.startswith("/") else "/tmp/ipi_" + unixsocket)
        else:
            self.sock = socket.create_connection((host, port), timeout=timeout)
        self.sock.settimeout(timeout)

    def run(self, atoms, calc=None, use_stress: bool = True, max_steps: Optional[int] = None) -> int:
        calc = calc if calc is not None else atoms.calc
        state, result, steps = "NEEDINIT", None, 0
        try:
            while True:
                msg = _recv_exact(self.sock, HEADER).decode("ascii").strip()
                if msg == "STATUS":
                    self.sock.sendall(_pad(state))
                elif msg == "INIT":
                    _recv_array(self.sock, np.int32, 1)
                    n = int(_recv_array(self.sock, np.int32, 1)[0])
                    _recv_exact(self.sock, n)
                    state = "READY"
                elif msg == "POSDATA":
                    cell = _recv_array(self.sock, np.float64, 9).reshape(3, 3).T * BOHR
                    _recv_array(self.sock, np.float64, 9)                       # inverse cell: recomputed where needed
                    n = int(_recv_array(self.sock, np.int32, 1)[0])
                    pos = _recv_array(self.sock, np.float64, 3 * n).reshape(n, 3) * BOHR
                    atoms.positions = pos
                    atoms.cell = cell
                    calc.calculate(atoms, ["energy", "forces", "stress"] if use_stress else ["energy", "forces"])
                    res = calc.results
                    vol = abs(np.linalg.det(cell))
                    s = np.asarray(res.get("stress", np.zeros(6)), dtype=np.float64)
                    stress = np.array([[s[0], s[5], s[4]], [s[5], s[1], s[3]], [s[4], s[3], s[2]]])
                    virial = -stress * vol if use_stress else np.zeros((3, 3))
                    result = (float(res["energy"]) / HARTREE, np.asarray(res["forces"], dtype=np.float64) / (HARTREE / BOHR),
                              virial / HARTREE)
                    state = "HAVEDATA"
                elif msg == "GETFORCE":
                    e, f, v = result
                    self.sock.sendall(_pad("FORCEREADY"))
                    self.sock.sendall(np.float64(e).tobytes())
                    self.sock.sendall(np.int32(f.shape[0]).tobytes())
                    self.sock.sendall(np.ascontiguousarray(f, dtype=np.float64).tobytes())
                    self.sock.sendall(np.ascontiguousarray(v.T, dtype=np.float64).tobytes())
                    self.sock.sendall(np.int32(0).tobytes())
                    state = "READY"
                    steps += 1
                    if max_steps is not None and steps >= max_steps:
                        return steps
                elif msg == "EXIT":
                    return steps
                else:
                    raise RuntimeError(f"i-PI protocol error: unexpected message {msg!r}")
        finally:
            self.sock.close()
