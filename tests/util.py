"""Shared helpers for the test-suite: golden fixtures, model construction, data objects."""
import json
import os

import numpy as np
import torch

import hermnet_b200 as H
from hermnet_b200 import ops
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
    extra = {k: cfg[k] for k in ("rbf", "envelope") if k in cfg}
    model = getattr(H, kind)(elems=cfg["elems"], rc=cfg["rc"], num_layers=cfg["num_layers"],
                             hidden_channels=cfg["hidden_channels"], num_rbf=cfg["num_rbf"], **extra, **kw)
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


# ---- synthetic lattice systems / frozen models / random edge-kernel inputs shared by the kernel-level tests ----------
def lattice_system(n_side, elems_z, seed, a=2.3, jitter=0.1):
    rng = np.random.default_rng(seed)
    g = np.stack(np.meshgrid(*[np.arange(n_side)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float64)
    pos = (g * a + rng.normal(0, jitter, g.shape)).astype(np.float32)
    Z = rng.choice(np.array(elems_z), size=len(pos))
    cell = (np.eye(3) * n_side * a).astype(np.float32)[None]
    return torch.from_numpy(pos), torch.from_numpy(Z).long(), torch.from_numpy(cell)


def frozen_model(kind, elems, F, K, dev, layers=2, seed=7):
    torch.manual_seed(seed)
    m = getattr(H, kind)(elems=elems, rc=5.0, num_layers=layers, hidden_channels=F, num_rbf=K).to(dev).eval()
    for p in m.parameters():
        p.requires_grad_(False)
    return m


def edge_inputs(model, g, pos, cell, seed=3):
    F, K = model.hidden_channels, model.num_rbf
    dev = pos.device
    gen = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=gen).to(dev)
    geom = ops.edge_geom_fwd(pos[g.perm].contiguous(), cell, g)
    xh = rn(g.xh_base[-1], 3 * F)
    vec = rn(g.n_atoms, 3, F)
    Wt = rn(g.n_modules, K, 3 * F) / np.sqrt(K)
    bias = rn(g.n_modules, 3 * F)
    g_dx, g_dvec = rn(g.n_rows, F), rn(g.n_rows, 3, F)
    p = ops.edge_params(g, g.n_modules, F, K, int(model.radial_basis.envelope.p), model.rc, model.radial_basis.rbf.coeff)
    return p, geom, xh, vec, Wt, bias, model.radial_basis.rbf.offset, g_dx, g_dvec
