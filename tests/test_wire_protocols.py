"""Wire protocols of the plugins over the CPU kernel emulator: the i-PI socket driver against a
minimal in-test i-PI server, the LAMMPS ``fix client/md`` message loop over an in-process CSlib stand-in, and the hydra
trainer's checkpoint format.  No arithmetic here: the numbers are compared with a direct calculator call."""
import os
import socket
import threading

import numpy as np
import pytest
import torch

import hermnet_b200 as H
from hermnet_b200.plugin import ipi, lammps_md, md
from hermnet_b200.plugin.calculator import NNCalculator
from hermnet_b200.symbols import atomic_numbers
from tests import util


def _model_and_atoms():
    case = util.load_case("c1_hvnet")
    cfg = dict(case["cfg"], num_layers=1, hidden_channels=32, num_rbf=16)
    model, _ = util.make_model("HVNet", cfg, 3)
    n = 24
    pos = case["pos"][:n].double().numpy()
    Z = case["Z"][:n].numpy()
    cell = case["cell"][0].double().numpy()
    return model, cfg, Z, pos, cell


def _ipi_server(srv, cell, frames, out):
    """The server half of the i-PI protocol for a fixed list of position frames (Angstrom in, atomic units on the wire)."""
    conn, _ = srv.accept()
    pad = lambda m: m.encode().ljust(12)

    def status():
        conn.sendall(pad("STATUS"))
        return ipi._recv_exact(conn, 12).decode().strip()
    assert status() == "NEEDINIT"
    conn.sendall(pad("INIT") + np.int32(0).tobytes() + np.int32(4).tobytes() + b"init")
    for pos in frames:
        assert status() == "READY"
        h = cell / ipi.BOHR
        conn.sendall(pad("POSDATA") + np.ascontiguousarray(h.T).tobytes() + np.ascontiguousarray(np.linalg.inv(h).T).tobytes()
                     + np.int32(len(pos)).tobytes() + np.ascontiguousarray(pos / ipi.BOHR).tobytes())
        assert status() == "HAVEDATA"
        conn.sendall(pad("GETFORCE"))
        assert ipi._recv_exact(conn, 12).decode().strip() == "FORCEREADY"
        e = ipi._recv_array(conn, np.float64, 1)[0]
        n = int(ipi._recv_array(conn, np.int32, 1)[0])
        f = ipi._recv_array(conn, np.float64, 3 * n).reshape(n, 3)
        v = ipi._recv_array(conn, np.float64, 9).reshape(3, 3)
        nx = int(ipi._recv_array(conn, np.int32, 1)[0])
        ipi._recv_exact(conn, nx)
        out.append((e * ipi.HARTREE, f * ipi.HARTREE / ipi.BOHR, v * ipi.HARTREE))
    conn.sendall(pad("EXIT"))
    conn.close()


def test_ipi_socket_driver_round_trip(emu, tmp_path):
    model, cfg, Z, pos, cell = _model_and_atoms()
    calc = NNCalculator(model, None, trn_mean=1.5, device_="cpu", ensemble="NPT")
    rng = np.random.default_rng(0)
    frames = [pos, pos + rng.normal(0, 0.02, pos.shape)]
    srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
    srv.bind(("127.0.0.1", 0))
    srv.listen(1)
    port = srv.getsockname()[1]
    got = []
    th = threading.Thread(target=_ipi_server, args=(srv, cell, frames, got), daemon=True)
    th.start()
    # POSCAR on disk -> atoms, like the reference's read_vasp (ipi_calc.py:8)
    symbols = [md.chemical_symbols[z] for z in Z]
    order = sorted(set(symbols), key=symbols.index)
    perm = [i for s in order for i, t in enumerate(symbols) if t == s]
    poscar = tmp_path / "POSCAR"
    with open(poscar, "w") as fh:
        fh.write("water\n1.0\n" + "\n".join(" ".join(f"{v:.10f}" for v in row) for row in cell) + "\n")
        fh.write(" ".join(order) + "\n" + " ".join(str(symbols.count(s)) for s in order) + "\nCartesian\n")
        fh.write("\n".join(" ".join(f"{v:.10f}" for v in pos[i]) for i in perm) + "\n")
    sym2, pos2, cell2 = ipi.read_poscar(str(poscar))
    assert sym2 == [symbols[i] for i in perm] and np.allclose(pos2, pos[perm]) and np.allclose(cell2, cell)
    atoms = md.SimpleAtoms(Z, pos, cell)
    steps = ipi.IPIClient(host="127.0.0.1", port=port).run(atoms, calc)
    th.join(timeout=30)
    assert steps == 2 and len(got) == 2
    for frame, (e, f, v) in zip(frames, got):
        calc.calculate(md.SimpleAtoms(Z, frame, cell), ["energy", "forces", "stress"])
        assert abs(e - calc.results["energy"]) < 1e-6 * max(1.0, abs(e))
        assert np.allclose(f, calc.results["forces"], atol=1e-6)
        s = calc.results["stress"]
        want = -np.array([[s[0], s[5], s[4]], [s[5], s[1], s[3]], [s[4], s[3], s[2]]]) * abs(np.linalg.det(cell))
        assert np.allclose(v, want.T, atol=1e-6 * max(1.0, np.abs(want).max()))


def test_lammps_md_message_loop(emu):
    model, cfg, Z, pos, cell = _model_and_atoms()
    elems = cfg["elems"]                                   # LAMMPS type i+1 <-> elems[i]
    types = [elems.index(md.chemical_symbols[z]) + 1 for z in Z]
    cs = lammps_md.LoopbackCS()
    L = lammps_md
    cs.client_send(0, [(1, b"md")])
    cs.client_send(L.SETUP, [(L.DIM, 3), (L.PERIODICITY, [1, 1, 1]), (L.ORIGIN, [0.0, 0.0, 0.0]), (L.BOX, cell.reshape(-1).tolist()),
                             (L.NATOMS, len(Z)), (L.NTYPES, 2), (L.TYPES, types), (L.COORDS, pos.reshape(-1).tolist()),
                             (L.UNITS, b"metal")])
    pos2 = pos + 0.01
    cs.client_send(L.STEP, [(L.COORDS, pos2.reshape(-1).tolist())])
    n = lammps_md.serve_md(cs, model, elems, cfg["rc"], trn_mean=0.25, device="cpu", periodic=True, ensemble="NPT")
    assert n == 2
    assert [r["msgID"] for r in cs.replies] == [0, L.SETUP, L.STEP, 0]
    calc = NNCalculator(model, None, trn_mean=0.25, device_="cpu", ensemble="NPT")
    for reply, p in zip(cs.replies[1:3], (pos, pos2)):
        calc.calculate(md.SimpleAtoms(Z, p, cell), ["energy", "forces", "stress"])
        assert reply["nfield"] == 3 and set(reply["fields"]) == {L.FORCES, L.ENERGY, L.VIRIAL}
        assert abs(reply["fields"][L.ENERGY] - calc.results["energy"]) < 1e-6
        assert np.allclose(np.array(reply["fields"][L.FORCES]).reshape(-1, 3), calc.results["forces"], atol=1e-6)
        assert len(reply["fields"][L.VIRIAL]) == 6
    # protocol errors
    bad = lammps_md.LoopbackCS()
    bad.client_send(0, [(1, b"mc")])
    with pytest.raises(RuntimeError, match="Mismatch"):
        lammps_md.serve_md(bad, model, elems, cfg["rc"], device="cpu")
    bad = lammps_md.LoopbackCS()
    bad.client_send(0, [(1, b"md")])
    bad.client_send(7, [])
    with pytest.raises(RuntimeError, match="unrecognized"):
        lammps_md.serve_md(bad, model, elems, cfg["rc"], device="cpu")


def test_trainer_checkpoint_format(tmp_path):
    """example/hydra-train/train.py:172-180: OrderedDict(model, trn_e_loss, trn_f_loss, val_e_loss, val_f_loss, trn_mean)."""
    model = H.HVNet(elems=["H", "O"], rc=4.0, num_layers=1, hidden_channels=32, num_rbf=8)
    path = str(tmp_path / "ckpt.pt")
    infos = H.save_checkpoint(path, model, trn_mean=-3.5, trn_e_loss=0.1, trn_f_loss=0.2, val_e_loss=0.3, val_f_loss=0.4)
    assert list(infos.keys()) == ["model", "trn_e_loss", "trn_f_loss", "val_e_loss", "val_f_loss", "trn_mean"]
    sd, meta = H.load_checkpoint(path)
    assert meta["trn_mean"] == -3.5 and set(sd) == set(model.state_dict())
    H.HVNet(elems=["H", "O"], rc=4.0, num_layers=1, hidden_channels=32, num_rbf=8).load_state_dict(sd, strict=True)
    torch.save(model.state_dict(), path)                      # the bare flavour of example/dist_train.py:141
    sd2, meta2 = H.load_checkpoint(path)
    assert meta2 == {} and set(sd2) == set(sd)
