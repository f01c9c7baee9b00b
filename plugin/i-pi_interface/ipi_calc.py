"""i-PI driver entry point (reference: ``/root/reference/plugin/i-pi_interface/ipi_calc.py:5-18``).  The socket client
is ASE's ``SocketClient`` (absent here, out of scope); the calculator handed to it is the B200-native one."""
from hermnet_b200.plugin.calculator import NNCalculator


def ipi_communicate(atoms, model, model_path, trn_mean, device='cuda', ensemble='NVT', port=31415, host='localhost'):
    assert ensemble.lower() in ('nvt', 'npt', 'nve')
    atoms.calc = NNCalculator(model, model_path, trn_mean, device_=device, ensemble=ensemble)
    try:
        from ase.calculators.socketio import SocketClient
    except Exception as exc:  # noqa: BLE001
        raise RuntimeError("i-PI needs ase.calculators.socketio.SocketClient, which is not installed") from exc
    client = SocketClient(host=host, port=port)
    client.run(atoms, use_stress=ensemble.lower() == 'npt')
