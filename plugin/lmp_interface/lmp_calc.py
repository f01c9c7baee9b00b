"""HermNet as a server for LAMMPS (``fix client/md``) -- counterpart of ``/root/reference/plugin/lmp_interface/lmp_calc.py``:
same command line (``-m -p -d -f -s -r -c -u -t -e``), same ``build_graph`` / ``calculator`` entry points, same message
loop (``hermnet_b200.plugin.lammps_md.serve_md``).  The transport is LAMMPS' CSlib (``from cslib import CSlib``), which is
not part of this image: the CLI fails with a clear message when it is missing, the loop itself is covered by the tests
over ``lammps_md.LoopbackCS``."""
import argparse
import sys

import torch

from hermnet_b200 import HVNet
from hermnet_b200.plugin.calculator import build_graph  # noqa: F401
from hermnet_b200.plugin.lammps_md import calculator, serve_md  # noqa: F401
from hermnet_b200.utils import load_checkpoint


def main(argv=None):
    parser = argparse.ArgumentParser(description="HermNet works as a server for LAMMPS.")
    parser.add_argument('-m', '--mode', help='The mode for exchange messages', type=str, choices=['file', 'zmq'], default='zmq')
    parser.add_argument('-p', '--ptr', help='Filename or socket ID or MPI communicator', type=str, default='tmp.couple')
    parser.add_argument('-d', '--device', help='Device to allocate HermNet', type=str, choices=['cpu', 'cuda'], default='cuda')
    parser.add_argument('-f', '--model', help='The path that saves trained model', type=str, required=True)
    parser.add_argument('-s', '--stats', help='The mean value of trainset that shifts the output of model', type=float, default=None)
    parser.add_argument('-r', '--radius', help='Cutof radius', type=float, required=True)
    parser.add_argument('-c', '--periodic', help='If the system is PBC or not', type=str, required=True)
    parser.add_argument('-u', '--units', help='Units', type=str, default='metal')
    parser.add_argument('-t', '--elems', help='Elements. The order should be the same with data file', type=str, nargs='*', required=True)
    parser.add_argument('-e', '--ensemble', help='Ensemble', type=str, default='NVT')
    parser.add_argument('--skin', help='Verlet skin (A): re-use the neighbour list between steps', type=float, default=0.0)
    args = parser.parse_args(argv)
    if args.device != 'cuda':
        print('Error: the B200-native hot path has no CPU implementation (use -d cuda)')
        sys.exit(1)
    try:
        from cslib import CSlib
    except ImportError:
        print('Error: LAMMPS CSlib python bindings (cslib) are not installed')
        sys.exit(1)
    device = torch.device(args.device)
    model = HVNet(elems=args.elems, rc=args.radius, intensive=False).to(device)
    sd, meta = load_checkpoint(args.model, map_location=device)     # plain state_dict or the hydra trainer's infos dict
    model.load_state_dict(sd)
    trn_mean = args.stats if args.stats is not None else float(meta.get('trn_mean', 0.0))
    cs = CSlib(1, args.mode.encode('ascii'), args.ptr.encode('ascii'), None)
    try:
        serve_md(cs, model, args.elems, args.radius, trn_mean, args.device, args.periodic.lower() in ('true', '1', 'yes'),
                 args.units, args.ensemble, args.skin)
    except RuntimeError as exc:
        print(exc)
        sys.exit(1)
    del cs


if __name__ == '__main__':
    main()
