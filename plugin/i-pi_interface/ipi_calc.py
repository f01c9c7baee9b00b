"""i-PI driver entry point with the reference's signature (``/root/reference/plugin/i-pi_interface/ipi_calc.py:5-18``):
``ipi_communicate(poscar, calc, host, port, mode)``.  With ASE installed the reference's own ``SocketClient`` is used;
otherwise the pure-Python client of ``hermnet_b200.plugin.ipi`` speaks the same wire protocol."""
from hermnet_b200.plugin import ipi, md
from hermnet_b200.plugin.calculator import NNCalculator  # noqa: F401  (re-export, as the reference imports it here)


def ipi_communicate(poscar: str, calc, host: str = 'localhost', port: int = 8888, mode: str = 'unix'):
    assert mode in ['inet', 'unix']
    try:
        from ase.calculators.socketio import SocketClient
        from ase.io.vasp import read_vasp
        atoms = read_vasp(poscar)
        atoms.calc = calc
        client = SocketClient(host=host, port=port) if mode == 'inet' else SocketClient(unixsocket=host)
        return client.run(atoms)
    except ImportError:
        symbols, pos, cell = ipi.read_poscar(poscar)
        from hermnet_b200.symbols import atomic_numbers
        atoms = md.SimpleAtoms([atomic_numbers[s] for s in symbols], pos, cell)
        atoms.calc = calc
        client = ipi.IPIClient(host=host, port=port) if mode == 'inet' else ipi.IPIClient(unixsocket=host)
        return client.run(atoms, calc)
