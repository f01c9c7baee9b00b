"""C4-size timing of the first-layer basis kernels alone (hn_layer0_basis_fwd / _bwd); run under ncu for a --set full capture:

    ncu --set full --clock-control none -k regex:layer0_basis -c 2 -o gpurun_out/prof_r2_layer0 python profiles/layer0_time.py 1
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hermnet_b200 as H  # noqa: E402
from hermnet_b200 import ops, synthetic  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    dev = torch.device("cuda", 0)
    (pos, Z, cell), cfg = synthetic.config("C4", 1.0)
    cfg = dict(cfg)
    cfg.pop("kind")
    torch.manual_seed(7)
    model = H.HVNet(**cfg).to(dev).eval()
    for q in model.parameters():
        q.requires_grad_(False)
    p_, z_, c_ = torch.from_numpy(pos).to(dev), torch.from_numpy(Z).to(dev), torch.from_numpy(cell)[None].to(dev)
    g = model.build_graph(p_, z_, c_)
    uniq, g0 = model._layer0_tables(g, z_[g.perm])
    geom = ops.edge_geom_fwd(p_[g.perm].contiguous(), c_, g)
    F, K = model.hidden_channels, model.num_rbf
    p = ops.edge_params(g, g.n_modules, F, K, int(model.radial_basis.envelope.p), model.rc, model.radial_basis.rbf.coeff)
    off = model.radial_basis.rbf.offset
    nz = int(uniq.numel())
    kp = ops.layer0_row_len(nz, K)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    Sa, Sc = ops.layer0_basis_fwd(p, g0, geom, None, off, nz, kp)
    gg = ops.layer0_basis_bwd(p, g0, geom, None, off, nz, kp, Sa, Sc)
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(reps):
        Sa, Sc = ops.layer0_basis_fwd(p, g0, geom, None, off, nz, kp)
    ev[1].record()
    for _ in range(reps):
        gg = ops.layer0_basis_bwd(p, g0, geom, None, off, nz, kp, Sa, Sc)
    ev[2].record()
    torch.cuda.synchronize()
    E = int(g.n_edges)
    print(f"N={len(Z)} E={E} n_elem={nz} KP={kp}: fwd {ev[0].elapsed_time(ev[1]) / reps:.3f} ms  bwd {ev[1].elapsed_time(ev[2]) / reps:.3f} ms  "
          f"(row bytes {16 * kp}, S traffic {16e-9 * kp * len(Z):.1f} GB per pass)", flush=True)


if __name__ == "__main__":
    main()
