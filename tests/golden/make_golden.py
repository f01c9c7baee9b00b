"""Generate the committed golden fixtures (run in the AUTHORING container only; needs /root/reference).

    python tests/golden/make_golden.py

For every HVNet case the fixture outputs come from the reference's OWN source
(/root/reference/HermNet/{hermnet,rmnet,utils,data}.py, imported unmodified over oracle/ref_shims.py):
  * ``energy``      -- unpatched reference forward;
  * ``forces``, ``cell_grad`` -- reference forward with ONLY ``HVNet.with_edge`` replaced by the out-of-place
    variant (the shipped in-place write makes autograd raise on torch 2.11), then
    ``-autograd.grad(E.sum(), pos)`` exactly as plugin/ase_interface/calculator.py:77-83 does;
  * ``edges``       -- ``HermNet.data.neighbor_search`` (reference code; ASE / torch_cluster calls shimmed by the
    numpy oracle) canonicalised to sorted rows (dst, src, Sx, Sy, Sz).
Weights are NOT stored: both sides regenerate them with ``oracle.hermnet_oracle.make_state_dict(kind, cfg, seed)``
and the reference model loads them through ``load_state_dict`` (pins the checkpoint key layout).
HPNet / HTNet do not exist in the reference; their fixtures come from the builder-owned oracle spec and are
regression vectors only ("parity unpinned").
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import hermnet_oracle as O  # noqa: E402
from oracle import neighbor_oracle as NO  # noqa: E402
from oracle import ref_shims  # noqa: E402
from hermnet_b200 import synthetic  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(8)


def patched_with_edge(self, data):
    """hermnet.py:133-152 with the F4 fix only."""
    j, i = data.edge_index
    dv = data.pos[j] - data.pos[i]
    if data.get('cell') is not None and data.get('edge_shift') is not None:
        dv = dv + torch.einsum('ni, nij -> nj', data.edge_shift, data.cell[data.batch[j]])
    d = dv.norm(dim=-1)
    d = torch.where(torch.isclose(d, torch.tensor(0.0), atol=1e-6), torch.full_like(d, 1.0e-6), d)
    data.edge_dist = d
    data.edge_vec = dv / d[:, None]
    return data


def cases():
    rng = np.random.default_rng(7)
    out = {}
    (pos, Z, cell), cfg = synthetic.config("C1")
    out["c1_hvnet"] = dict(pos=pos, Z=Z, cell=cell[None], batch=np.zeros(len(Z), np.int64), cfg=cfg, seed=1234)
    # tiny triclinic cell, rc > L/2 -> several images of the same pair, self-image edges, unwrapped positions
    cell = np.array([[4.1, 0.0, 0.0], [1.3, 3.9, 0.0], [0.7, -0.9, 4.4]], np.float32)
    pos = (rng.uniform(-0.6, 1.7, (5, 3)) @ cell).astype(np.float32)
    out["triclinic_multi_image"] = dict(pos=pos, Z=np.array([1, 8, 1, 8, 6], np.int64), cell=cell[None],
                                        batch=np.zeros(5, np.int64), seed=11,
                                        cfg=dict(kind="HVNet", elems=["H", "O", "C"], rc=5.0, num_layers=2,
                                                 hidden_channels=32, num_rbf=32))
    # non-periodic molecule-like cluster (radius_graph branch, 32-neighbour cap inactive and active)
    pos = rng.normal(0, 2.2, (21, 3)).astype(np.float32)
    out["cluster_nonpbc"] = dict(pos=pos, Z=rng.choice([1, 6, 8], 21).astype(np.int64), cell=None,
                                 batch=np.zeros(21, np.int64), seed=12,
                                 cfg=dict(kind="HVNet", elems=["C", "H", "O"], rc=5.0, num_layers=2,
                                          hidden_channels=64, num_rbf=20))
    pos = rng.normal(0, 1.6, (60, 3)).astype(np.float32)
    out["cluster_nonpbc_capped"] = dict(pos=pos, Z=rng.choice([1, 8], 60).astype(np.int64), cell=None,
                                        batch=np.zeros(60, np.int64), seed=13,
                                        cfg=dict(kind="HVNet", elems=["H", "O"], rc=5.0, num_layers=1,
                                                 hidden_channels=32, num_rbf=16))
    # batch of three periodic graphs: different cells, an element absent from elems (Ar) and an element (C)
    # whose only atom has no incoming edge in graph 2 (isolated in a huge cell) -> zero rows (hermnet.py:56-57)
    ps, zs, cs, bs = [], [], [], []
    for g, (n, L) in enumerate([(14, 7.3), (9, 6.1), (1, 40.0)]):
        c = (np.eye(3) * L + rng.normal(0, 0.3, (3, 3))).astype(np.float32)
        ps.append((rng.uniform(0, 1, (n, 3)) @ c).astype(np.float32))
        zs.append(np.array([6], np.int64) if n == 1 else rng.choice([1, 8, 18], n).astype(np.int64))
        cs.append(c); bs.append(np.full(n, g, np.int64))
    out["batch3_mixed"] = dict(pos=np.concatenate(ps), Z=np.concatenate(zs), cell=np.stack(cs),
                               batch=np.concatenate(bs), seed=14,
                               cfg=dict(kind="HVNet", elems=["H", "O", "C"], rc=5.0, num_layers=2,
                                        hidden_channels=32, num_rbf=24))
    # alternative radial bases / envelopes (rmnet.py:196-275) on a 24-atom water box, bar-sized K so that the
    # Bernstein binomials stay in fp32 range
    (pos, Z, cell), _ = synthetic.water_box(2, seed=31), None
    for nm, rbf, env in (("a17_bessel_poly", {"name": "spherical_bessel"}, {"name": "polynomial", "exponent": 5}),
                         ("a17_bernstein_exp", {"name": "bernstein"}, {"name": "exponential"}),
                         ("a17_gauss_exp", {"name": "gaussian"}, {"name": "exponential"})):
        out[nm] = dict(pos=pos, Z=Z, cell=cell[None], batch=np.zeros(len(Z), np.int64), seed=17,
                       cfg=dict(kind="HVNet", elems=["H", "O"], rc=4.0, num_layers=2, hidden_channels=32, num_rbf=16,
                                rbf=rbf, envelope=env))
    # builder-owned models (oracle only)
    (pos, Z, cell), _ = synthetic.water_box(2, seed=21), None
    out["water24_hpnet"] = dict(pos=pos, Z=Z, cell=cell[None], batch=np.zeros(len(Z), np.int64), seed=15,
                                cfg=dict(kind="HPNet", elems=["H", "O"], rc=4.0, num_layers=2,
                                         hidden_channels=32, num_rbf=32))
    out["water24_htnet"] = dict(pos=pos, Z=Z, cell=cell[None], batch=np.zeros(len(Z), np.int64), seed=16,
                                cfg=dict(kind="HTNet", elems=["H", "O"], rc=4.0, num_layers=2,
                                         hidden_channels=32, num_rbf=32))
    return out


def build_edges(H, case):
    """Reference neighbour search per graph, collated the PyG way (node offset, concatenation)."""
    pos, batch, cell = case["pos"], case["batch"], case["cell"]
    eis, ess, off = [], [], 0
    for g in range(int(batch.max()) + 1):
        p = torch.from_numpy(pos[batch == g])
        if cell is None:
            ei = H.neighbor_search(p, case["cfg"]["rc"])
            es = None
        else:
            ei, es = H.neighbor_search(p, case["cfg"]["rc"], torch.from_numpy(cell[g:g + 1]))
            ess.append(es)
        eis.append(ei + off)
        off += p.shape[0]
    return torch.cat(eis, 1), (torch.cat(ess, 0) if ess else None)


def main():
    H = ref_shims.import_reference()
    only = set(sys.argv[1:])           # python make_golden.py [case ...]: (re)generate only these, keep the other index entries
    index = json.load(open(os.path.join(OUT, "index.json"))) if only and os.path.exists(os.path.join(OUT, "index.json")) else {}
    for name, case in cases().items():
        if only and name not in only:
            continue
        cfg = case["cfg"]
        kind = cfg["kind"]
        mcfg = {k: v for k, v in cfg.items() if k != "kind"}
        sd = O.make_state_dict(kind, mcfg, case["seed"])
        ei, es = build_edges(H, case)
        pos = torch.from_numpy(case["pos"]); Z = torch.from_numpy(case["Z"]); batch = torch.from_numpy(case["batch"])
        cell = None if case["cell"] is None else torch.from_numpy(case["cell"])
        rec = dict(pos=case["pos"], Z=case["Z"], batch=case["batch"], edges=NO.canonical_edges(ei.numpy(), None if es is None else es.numpy()))
        if cell is not None:
            rec["cell"] = case["cell"]
        if kind == "HVNet":
            extra = {k: mcfg[k] for k in ("rbf", "envelope") if k in mcfg}
            model = H.HVNet(elems=mcfg["elems"], rc=mcfg["rc"], num_layers=mcfg["num_layers"],
                            hidden_channels=mcfg["hidden_channels"], num_rbf=mcfg["num_rbf"], **extra)
            model.load_state_dict(sd, strict=True)
            mk = lambda p, c: ref_shims.Data(pos=p, atomic_number=Z, edge_index=ei, batch=batch,
                                             **({} if c is None else dict(cell=c, edge_shift=es)))
            with torch.no_grad():
                e_ref = model(mk(pos, cell))                      # unpatched reference forward
            H.HVNet.with_edge = patched_with_edge                 # F4 fix for the autograd path only
            p = pos.clone().requires_grad_(True)
            c = None if cell is None else cell.clone().requires_grad_(True)
            e2 = model(mk(p, c))
            grads = torch.autograd.grad(e2.sum(), [p] + ([] if c is None else [c]))
            assert torch.equal(e2.detach(), e_ref), "F4 patch must be forward-identical"
            rec.update(energy=e_ref.numpy(), forces=(-grads[0]).numpy(), source="reference-over-shims")
            if c is not None:
                rec["cell_grad"] = grads[1].numpy()
            # the oracle must reproduce the reference bit for bit on CPU
            eo, fo = O.energy_and_forces(kind, sd, mcfg, pos, Z, ei, cell, es, batch)
            assert torch.equal(eo, e_ref) and torch.allclose(fo, -grads[0], atol=1e-6), name
        else:
            out = O.energy_and_forces(kind, sd, mcfg, pos, Z, ei, cell, es, batch, want_cell_grad=True)
            rec.update(energy=out[0].numpy(), forces=out[1].numpy(), cell_grad=out[2].numpy(), source="oracle-spec")
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **{k: v for k, v in rec.items() if k != "source"})
        index[name] = dict(cfg=cfg, seed=case["seed"], source=rec["source"], n_atoms=int(len(Z)),
                           n_edges=int(ei.shape[1]), energy=[float(v) for v in rec["energy"]])
        print(name, index[name]["n_atoms"], index[name]["n_edges"], rec["energy"], rec["source"])
    with open(os.path.join(OUT, "index.json"), "w") as f:
        json.dump(index, f, indent=1)


if __name__ == "__main__":
    main()
