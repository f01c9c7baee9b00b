"""Energy / force parity at benchmark size against the oracle on a CUT-OUT of the big system (test infrastructure).

The oracle (CPU restatement of the reference, oracle/hermnet_oracle.py) cannot run 10^5..10^6 atoms, but message passing
is local: an atom's energy ``e_a`` depends on the atoms within ``L * rc`` (L layers).  For a region ``A`` (atoms within
``r_A`` of a centre atom) the partial energy ``E_A = sum_{a in A} e_a`` and its gradient w.r.t. EVERY position are exact
functions of the atoms within ``r_A + L rc`` -- the cut-out.  The product path evaluates ``E_A`` on the FULL system
(full-size graph, full-size kernel launches, ``atom_weight`` selects the region in the readout); the oracle evaluates
the same sum on the cut-out alone, with the same fp32 positions, the full cell and the edge shifts of the full system
(so the geometry arithmetic is the reference's, hermnet.py:136-148).  For ``r_A >= r_in + L rc`` the gradient rows of the
atoms within ``r_in`` are the system's true forces, which is checked against an ordinary full evaluation as well.
"""
from __future__ import annotations

import time

import numpy as np
import torch

from hermnet_b200.graph import GraphBuilder
from oracle import hermnet_oracle as O


def min_image_dist(pos: torch.Tensor, cell: torch.Tensor, centre: int) -> torch.Tensor:
    """Minimum-image distance of every atom to atom ``centre`` (float64, general cell)."""
    c = cell.reshape(3, 3).double()
    d = pos.double() - pos[centre].double()
    frac = torch.linalg.solve(c.T, d.T).T
    frac = frac - torch.round(frac)
    return (frac @ c).norm(dim=1)


def cutout_parity(model, sd, cfg, pos, Z, cell, r_in: float = 4.0, centre: int | None = None, graph=None,
                  check_full_forces: bool = True, r_region: float | None = None, fp64_reference: bool = False):
    """Returns a dict with ``rel_dE`` (partial energy), ``max_dF`` (gradient of the partial energy over the whole
    cut-out, eV/A), ``max_dF_interior`` (true forces of the interior atoms vs the oracle), sizes and timings.
    ``sd`` / ``cfg``: oracle-format state dict and config of ``model`` (tests/util.make_model)."""
    dev = pos.device
    kind = model.KIND
    L, rc = model.num_layers, model.rc
    n = pos.size(0)
    if centre is None:      # the atom closest to the middle of the cell
        mid = cell.reshape(3, 3).double().sum(0) * 0.5
        centre = int((pos.double() - mid).norm(dim=1).argmin())
    d = min_image_dist(pos, cell, centre)
    r_A = r_in + L * rc if r_region is None else float(r_region)      # (a smaller region: no true-force claim)
    check_full_forces = check_full_forces and r_A >= r_in + L * rc
    in_A = d < r_A
    in_cut = d < r_A + L * rc
    interior = d < r_in
    g = graph if graph is not None else model.build_graph(pos, Z, cell)
    # ---- product path: E_A and its gradient on the full system
    t0 = time.perf_counter()
    p = pos.detach().clone().requires_grad_(True)
    e_A, _, _ = model.forward_graph(p, Z, cell.reshape(-1, 3, 3), g, atom_weight=in_A)
    (grad_A,) = torch.autograd.grad(e_A.sum(), p)
    torch.cuda.synchronize() if dev.type == "cuda" else None
    t_gpu = time.perf_counter() - t0
    assert float(grad_A[~in_cut].abs().max() if bool((~in_cut).any()) else 0.0) == 0.0, "E_A must not depend on atoms outside the cut-out"
    # ---- physical edge list of the cut-out (both end points inside), original atom ids
    phys = g if g.rows_per_atom == 1 else GraphBuilder("HVNet", model.elems, rc, model.pbc_shift).from_positions(pos, Z, cell, None)
    src = phys.perm[phys.col.long()]
    dst = phys.perm[phys.edge_row.long()]
    keep = in_cut[src] & in_cut[dst]
    ids = torch.nonzero(in_cut).squeeze(1)
    g2l = torch.full((n,), -1, dtype=torch.long, device=dev)
    g2l[ids] = torch.arange(ids.numel(), device=dev)
    ei = torch.stack([g2l[src[keep]], g2l[dst[keep]]]).cpu()
    es = (phys.shift[keep][:, :3].to(torch.float32) * phys.sign).cpu()
    # ---- oracle on the cut-out: graph 0 = region A, graph 1 = the rest of the cut-out (same cell for both)
    pc, zc = pos[ids].detach().cpu(), Z[ids].cpu()
    batch = (~in_A[ids]).long().cpu()
    cell2 = cell.reshape(1, 3, 3).cpu().repeat(2, 1, 1)
    t0 = time.perf_counter()
    po = pc.clone().requires_grad_(True)
    eo = O.FORWARDS[kind](sd, cfg, po, zc, ei, cell2, es, batch, 2)
    (go,) = torch.autograd.grad(eo[0], po)
    t_cpu = time.perf_counter() - t0
    out = {"kind": kind, "n_atoms": n, "n_cutout": int(ids.numel()), "n_region": int(in_A.sum()), "n_interior": int(interior.sum()),
           "edges_cutout": int(ei.size(1)), "r_in": r_in, "r_region": r_A, "r_cutout": r_A + L * rc,
           "rel_dE": float((e_A.detach().cpu().sum() - eo[0].detach()).abs() / eo[0].detach().abs().clamp(min=1e-30)),
           "E_region": float(eo[0].detach()),
           "max_dF": float((grad_A[ids].cpu() - go).abs().max()), "max_F": float(go.abs().max()),
           "seconds_product": t_gpu, "seconds_oracle": t_cpu}
    if fp64_reference:       # context for the fp32 bar: both fp32 implementations against the same sum evaluated in float64
        sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
        e64 = O.FORWARDS[kind](sd64, cfg, pc.double(), zc, ei, cell2.double(), es.double(), batch, 2)[0].detach()
        out["rel_dE_vs_fp64"] = float((e_A.detach().cpu().double().sum() - e64).abs() / e64.abs().clamp(min=1e-30))
        out["oracle_fp32_vs_fp64"] = float((eo[0].detach().double() - e64).abs() / e64.abs().clamp(min=1e-30))
    if check_full_forces:
        p2 = pos.detach().clone().requires_grad_(True)
        e_full, _, _ = model.forward_graph(p2, Z, cell.reshape(-1, 3, 3), g)
        (grad_full,) = torch.autograd.grad(e_full.sum(), p2)
        loc = g2l[torch.nonzero(interior).squeeze(1)].cpu()
        out["max_dF_interior"] = float((grad_full[interior].cpu() - go[loc]).abs().max())
        out["E_total"] = float(e_full.detach().sum())
    return out
