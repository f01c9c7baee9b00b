"""Shared helpers for the test-suite: golden fixtures, model construction, data objects."""
import json
import os

import numpy as np
import torch

import hermnet_b200 as H
from oracle import hermnet_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
INDEX = json.load(open(os.path.join(GOLD, "index.json")))
HV_CASES = [k for k, v in INDEX.items() if v["cfg"]["kind"] == "HVNet"]
ALL_CASES = list(INDEX.keys())


def load_case(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    meta = INDEX[name]
    cfg = dict(meta["cfg"])
    kind = cfg.pop("kind")
    rec = {k: z[k] for k in z.files}
    edges = rec["edges"]   # rows (dst, src, Sx, Sy, Sz), canonically sorted
    out = dict(kind=kind, cfg=cfg, seed=meta["seed"],
               pos=torch.from_numpy(rec["pos"]), Z=torch.from_numpy(rec["Z"]), batch=torch.from_numpy(rec["batch"]),
               cell=torch.from_numpy(rec["cell"]) if "cell" in rec else None,
               edge_index=torch.from_numpy(np.stack([edges[:, 1], edges[:, 0]])),
               edge_shift=torch.from_numpy(edges[:, 2:5].astype(np.float32)) if "cell" in rec else None,
               edges=edges, energy=torch.from_numpy(rec["energy"]), forces=torch.from_numpy(rec["forces"]),
               cell_grad=torch.from_numpy(rec["cell_grad"]) if "cell_grad" in rec else None)
    return out


def make_model(kind, cfg, seed, device="cpu", **kw):
    model = getattr(H, kind)(elems=cfg["elems"], rc=cfg["rc"], num_layers=cfg["num_layers"],
                             hidden_channels=cfg["hidden_channels"], num_rbf=cfg["num_rbf"], **kw)
    sd = O.make_state_dict(kind, cfg, seed)
    model.load_state_dict(sd, strict=True)
    return model.to(device).eval(), sd


def make_data(case, device="cpu", with_edges=True, requires_grad=True):
    d = H.Data(pos=case["pos"].clone().to(device), atomic_number=case["Z"].to(device), batch=case["batch"].to(device))
    if case["cell"] is not None:
        d.cell = case["cell"].clone().to(device)
    if with_edges:
        d.edge_index = case["edge_index"].to(device)
        if case["edge_shift"] is not None:
            d.edge_shift = case["edge_shift"].to(device)
    if requires_grad:
        d.pos.requires_grad_(True)
        if case["cell"] is not None:
            d.cell.requires_grad_(True)
    return d


def energy_forces(model, data):
    e = model(data)
    wrt = [data.pos] + ([data.cell] if data.get("cell") is not None and data.cell.requires_grad else [])
    grads = torch.autograd.grad(e.sum(), wrt)
    return e.detach(), -grads[0], (grads[1] if len(grads) > 1 else None)


def rel_err(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))
