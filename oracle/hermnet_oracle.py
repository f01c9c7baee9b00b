"""Pure-PyTorch CPU restatement of the reference forward pass (TEST INFRASTRUCTURE).

Written functionally against the reference ``state_dict`` layout instead
of ``nn.Module`` classes, so that it also pins checkpoint-key compatibility.  Every
function cites the reference lines it follows.  Works in fp32 (parity) and fp64
(finite-difference / gradcheck) -- dtype follows ``pos``/the weights.

Permitted deviations from the literal reference:
  1. ``with_edge``: the in-place ``edge_dist[mask] = 1e-6`` (hermnet.py:146-147) is done
     with ``torch.where`` -- forward-identical, and autograd-safe on torch >= 2.x.
  2. a bare graph gets ``batch = 0`` and ``cell`` of shape ``[1,3,3]``.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as Fnn

from .symbols import atomic_numbers

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# small building blocks
# --------------------------------------------------------------------------------------
def scaled_silu(z: Tensor) -> Tensor:
    """rmnet.py:110-117 -- ``SiLU(x) * (1/0.6)``."""
    return Fnn.silu(z) * (1.0 / 0.6)


def scatter_rows(src: Tensor, index: Tensor, dim_size: int, reduce: str = "sum") -> Tensor:
    """torch_scatter.scatter(src, index, dim=0, dim_size=..., reduce=...) [upstream].

    ``sum``: rows added in index order (``index_add_`` on CPU is sequential).
    ``mean``: sum divided by the per-row count, empty rows stay 0.
    """
    out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    out.index_add_(0, index, src)
    if reduce == "mean":
        cnt = torch.zeros(dim_size, dtype=src.dtype, device=src.device)
        cnt.index_add_(0, index, torch.ones_like(index, dtype=src.dtype))
        cnt = cnt.clamp(min=1)
        out = out / cnt.view((-1,) + (1,) * (src.dim() - 1))
    elif reduce != "sum":
        raise ValueError(reduce)
    return out


def polynomial_envelope(u: Tensor, p: int) -> Tensor:
    """rmnet.py:183-193."""
    a = -(p + 1) * (p + 2) / 2
    b = p * (p + 2)
    c = -p * (p + 1) / 2
    val = 1 + a * u ** p + b * u ** (p + 1) + c * u ** (p + 2)
    return torch.where(u < 1, val, torch.zeros_like(u))


def exponential_envelope(u: Tensor) -> Tensor:
    """rmnet.py:196-208."""
    val = torch.exp(-(u ** 2) / ((1 - u) * (1 + u)))
    return torch.where(u < 1, val, torch.zeros_like(u))


def gaussian_smearing(u: Tensor, offset: Tensor) -> Tensor:
    """PyG ``GaussianSmearing(start=0, stop=1, num_gaussians=K)`` [upstream]:
    ``coeff = -0.5/(offset[1]-offset[0])**2`` (a Python float), ``exp(coeff*(u-offset)^2)``."""
    coeff = -0.5 / (offset[1] - offset[0]).item() ** 2
    diff = u.view(-1, 1) - offset.view(1, -1)
    return torch.exp(coeff * diff.pow(2))


def radial_basis(d: Tensor, sd: Dict[str, Tensor], cfg: dict, prefix: str = "radial_basis.") -> Tensor:
    """rmnet.py:168-172: ``env(d/rc)[:,None] * rbf(d/rc)``."""
    u = d * (1.0 / cfg["rc"])
    env_cfg = cfg.get("envelope", {"name": "polynomial", "exponent": 5})
    rbf_cfg = cfg.get("rbf", {"name": "gaussian"})
    if env_cfg["name"].lower() == "polynomial":
        env = polynomial_envelope(u, env_cfg.get("exponent", 5))
    elif env_cfg["name"].lower() == "exponential":
        env = exponential_envelope(u)
    else:
        raise ValueError(f"Unknown envelope function '{env_cfg['name']}'.")
    name = rbf_cfg["name"].lower()
    if name == "gaussian":
        basis = gaussian_smearing(u, sd[prefix + "rbf.offset"])
    elif name == "spherical_bessel":  # rmnet.py:211-233
        norm = math.sqrt(2 / (cfg["rc"] ** 3))
        freq = sd[prefix + "rbf.frequencies"]
        basis = norm / u[:, None] * torch.sin(freq * u[:, None])
    elif name == "bernstein":  # rmnet.py:236-275
        from scipy.special import binom
        import numpy as np
        K = cfg["num_rbf"]
        pref = torch.tensor(binom(K - 1, np.arange(K)), dtype=torch.float).to(u.dtype)
        gamma = Fnn.softplus(sd[prefix + "rbf.pregamma"])
        e = torch.exp(-gamma * u)[:, None]
        k = torch.arange(K)[None, :]
        basis = pref * (e ** k) * ((1 - e) ** (K - 1 - k))
    else:
        raise ValueError(f"Unknown radial basis function '{name}'.")
    return env[:, None] * basis


# --------------------------------------------------------------------------------------
# geometry
# --------------------------------------------------------------------------------------
def with_edge(pos: Tensor, edge_index: Tensor, cell: Optional[Tensor], edge_shift: Optional[Tensor],
              batch: Tensor, pbc_shift: str = "reference"):
    """hermnet.py:133-152.  ``pbc_shift='reference'`` reproduces the +S quirk,
    ``'physical'`` uses -S."""
    src, dst = edge_index[0], edge_index[1]
    dvec = pos[src] - pos[dst]
    if cell is not None and edge_shift is not None:
        sgn = 1.0 if pbc_shift == "reference" else -1.0
        dvec = dvec + sgn * torch.einsum("ni,nij->nj", edge_shift.to(pos.dtype), cell[batch[src]])
    dist = dvec.norm(dim=-1)
    near0 = torch.isclose(dist, torch.zeros((), dtype=dist.dtype), atol=1e-6)
    dist = torch.where(near0, torch.full_like(dist, 1.0e-6), dist)  # deviation (1)
    unit = dvec / dist[:, None]
    return dist, unit


# --------------------------------------------------------------------------------------
# heterogeneous sub-graph (utils.py:11-24)
# --------------------------------------------------------------------------------------
def in_subgraph_edges_literal(edge_dst: Tensor, nids: Tensor) -> Tensor:
    """The literal O(|nids|*E) loop of utils.py:14."""
    if nids.numel() == 0:
        return torch.zeros(0, dtype=torch.long)
    return torch.cat([torch.where(edge_dst == nid)[0] for nid in nids])


def in_subgraph_edges(edge_dst: Tensor, nids: Tensor, num_nodes: int) -> Tensor:
    """Vectorised equivalent: edges whose destination is in ``nids``, regrouped by
    destination ascending, original order inside a destination (== stable sort)."""
    member = torch.zeros(num_nodes, dtype=torch.bool)
    member[nids] = True
    sel = torch.where(member[edge_dst])[0]
    order = torch.sort(edge_dst[sel], stable=True).indices
    return sel[order]


# --------------------------------------------------------------------------------------
# modified PaiNN block (rmnet.py:11-107)
# --------------------------------------------------------------------------------------
def _lin(x, sd, key, bias=True):
    return Fnn.linear(x, sd[key + ".weight"], sd[key + ".bias"] if bias else None)


def painn_message_terms(x, vec, src, emb, unit, sd, pre, F):
    """rmnet.py:51-67: per-edge scalar message ``[E,F]`` and vector message ``[E,3,F]``."""
    h = Fnn.layer_norm(x, (F,), sd[pre + "x_layernorm.weight"], sd[pre + "x_layernorm.bias"], 1e-5)
    xh = _lin(scaled_silu(_lin(h, sd, pre + "x_proj.0")), sd, pre + "x_proj.2")
    rbfh = _lin(emb, sd, pre + "rbf_proj")
    prod = xh[src] * rbfh
    m_x, g2, g3 = torch.split(prod, F, dim=-1)
    g2 = g2 * (1 / math.sqrt(3.0))
    m_vec = vec[src] * g2.unsqueeze(1) + g3.unsqueeze(1) * unit.unsqueeze(2)
    m_vec = m_vec * (1 / math.sqrt(F))
    return m_x, m_vec


def painn_update(x, vec, sd, pre, F, vdot_override=None):
    """rmnet.py:94-107 (``vdot_override`` is the HTNet triadic replacement)."""
    vp = Fnn.linear(vec, sd[pre + "vec_proj.weight"])
    v1, v2 = torch.split(vp, F, dim=-1)
    vdot = (v1 * v2).sum(dim=1) * (1 / math.sqrt(F)) if vdot_override is None else vdot_override
    vn = torch.sqrt(torch.sum(v2 ** 2, dim=-2) + 1e-8)
    h = _lin(scaled_silu(_lin(torch.cat([x, vn], dim=-1), sd, pre + "xvec_proj.0")), sd, pre + "xvec_proj.2")
    a1, a2, a3 = torch.split(h, F, dim=-1)
    dx = (a1 + a2 * vdot) * (1 / math.sqrt(2.0))
    dvec = a3.unsqueeze(1) * v1
    return dx, dvec


def painn_module(x, vec, src, dst, emb, unit, sd, pre, F):
    """rmnet.py:21-32 -- returns ``(vec, x)`` in the reference's order."""
    N = x.size(0)
    m_x, m_vec = painn_message_terms(x, vec, src, emb, unit, sd, pre + "message_layer.", F)
    dx = scatter_rows(m_x, dst, N)
    dvec = scatter_rows(m_vec, dst, N)
    x = (x + dx) * (1 / math.sqrt(2.0))
    vec = vec + dvec
    dx2, dvec2 = painn_update(x, vec, sd, pre + "update_layer.", F)
    return vec + dvec2, x + dx2


# --------------------------------------------------------------------------------------
# HVNet (hermnet.py:37-65, 118-131)
# --------------------------------------------------------------------------------------
def _prologue(sd, cfg, pos, Z, edge_index, cell, edge_shift, batch, pbc_shift):
    if batch is None:
        batch = torch.zeros(pos.size(0), dtype=torch.long)  # deviation (2)
    if cell is not None and cell.dim() == 2:
        cell = cell.unsqueeze(0)
    dist, unit = with_edge(pos, edge_index, cell, edge_shift, batch, pbc_shift)
    emb = radial_basis(dist, sd, cfg)
    x = sd["embed.weight"][Z.long()]
    vec = torch.zeros((x.size(0), 3, cfg["hidden_channels"]), dtype=x.dtype)
    return batch, dist, unit, emb, x, vec


def _readout(sd, cfg, x, batch, num_graphs):
    """hermnet.py:129-130."""
    e_atom = _lin(scaled_silu(_lin(x, sd, "out_energy.0")), sd, "out_energy.2").squeeze(1)
    if num_graphs is None:
        num_graphs = int(batch.max().item()) + 1 if batch.numel() else 0
    return scatter_rows(e_atom, batch, num_graphs, "mean" if cfg.get("intensive", False) else "sum")


def hvnet_forward(sd: Dict[str, Tensor], cfg: dict, pos: Tensor, Z: Tensor, edge_index: Tensor,
                  cell: Optional[Tensor] = None, edge_shift: Optional[Tensor] = None,
                  batch: Optional[Tensor] = None, num_graphs: Optional[int] = None,
                  pbc_shift: str = "reference", literal_subgraph: bool = False,
                  return_features: bool = False):
    """HVNet.forward (hermnet.py:118-131) with HeteroVertexConv (hermnet.py:37-65)."""
    F = cfg["hidden_channels"]
    batch, dist, unit, emb, x, vec = _prologue(sd, cfg, pos, Z, edge_index, cell, edge_shift, batch, pbc_shift)
    N = x.size(0)
    dst_all = edge_index[1]
    for layer in range(cfg["num_layers"]):
        x_new, vec_new = torch.zeros_like(x), torch.zeros_like(vec)
        for el in cfg["elems"]:
            nid = torch.where(Z == atomic_numbers[el])[0]
            sel = (in_subgraph_edges_literal(dst_all, nid) if literal_subgraph
                   else in_subgraph_edges(dst_all, nid, N))
            if sel.numel() == 0:      # hermnet.py:56-57 -- rows stay zero
                continue
            pre = f"hermconvs.{layer}.mods.{el}."
            vec_t, x_t = painn_module(x, vec, edge_index[0][sel], edge_index[1][sel], emb[sel], unit[sel],
                                      sd, pre, F)
            mask = torch.zeros(N, dtype=x.dtype)
            mask[nid] = 1
            vec_new = vec_new + vec_t * mask[:, None, None]   # vrsts[nid] += vrst[nid]
            x_new = x_new + x_t * mask[:, None]
        x, vec = x_new, vec_new
    energy = _readout(sd, cfg, x, batch, num_graphs)
    if return_features:
        return energy, x, vec
    return energy


# --------------------------------------------------------------------------------------
# HPNet / HTNet -- builder-owned spec; NOT in the reference: parity unpinned
# --------------------------------------------------------------------------------------
def pair_key(src_el: str, dst_el: str) -> str:
    return f"{src_el}-{dst_el}"


def triad_key(a_el: str, centre_el: str, c_el: str) -> str:
    return f"{a_el}-{centre_el}-{c_el}"


def hpnet_forward(sd, cfg, pos, Z, edge_index, cell=None, edge_shift=None, batch=None, num_graphs=None,
                  pbc_shift="reference", return_features=False):
    """One PaiNN sub-network per ordered element pair ``src->dst`` (figs/arch.svg (c)); the outputs of
    all ``*->t`` sub-networks are summed into the rows of element ``t``."""
    F = cfg["hidden_channels"]
    batch, dist, unit, emb, x, vec = _prologue(sd, cfg, pos, Z, edge_index, cell, edge_shift, batch, pbc_shift)
    N = x.size(0)
    zs, zd = Z[edge_index[0]], Z[edge_index[1]]
    for layer in range(cfg["num_layers"]):
        x_new, vec_new = torch.zeros_like(x), torch.zeros_like(vec)
        for dst_el in cfg["elems"]:
            rows = (Z == atomic_numbers[dst_el]).to(x.dtype)
            for src_el in cfg["elems"]:
                sel = torch.where((zs == atomic_numbers[src_el]) & (zd == atomic_numbers[dst_el]))[0]
                sel = sel[torch.sort(edge_index[1][sel], stable=True).indices]
                if sel.numel() == 0:
                    continue
                pre = f"hermconvs.{layer}.mods.{pair_key(src_el, dst_el)}."
                vec_t, x_t = painn_module(x, vec, edge_index[0][sel], edge_index[1][sel], emb[sel], unit[sel],
                                          sd, pre, F)
                vec_new = vec_new + vec_t * rows[:, None, None]
                x_new = x_new + x_t * rows[:, None]
        x, vec = x_new, vec_new
    energy = _readout(sd, cfg, x, batch, num_graphs)
    if return_features:
        return energy, x, vec
    return energy


def htnet_triads(elems: Sequence[str]) -> List[tuple]:
    """All (centre t, A, C) with A<=C in constructor order of ``elems``."""
    out = []
    for t in elems:
        for ia, a in enumerate(elems):
            for c in elems[ia:]:
                out.append((t, a, c))
    return out


def htnet_forward(sd, cfg, pos, Z, edge_index, cell=None, edge_shift=None, batch=None, num_graphs=None,
                  pbc_shift="reference", return_features=False, explicit_triplets=False):
    """Triadic network (figs/arch.svg (d), subnetwork.svg (d)-(f)) per SURVEY.md A.3:
    sub-network per (centre t, {A,C}); radial part over edges with source element in {A,C};
    ``vdot`` of the update replaced by <P^A/|P^A|, P^C/|P^C|> where P^X sums the vector messages
    coming from neighbours of element X.  ``explicit_triplets=True`` evaluates the inner product
    as an explicit double loop over neighbour pairs (j,k) -- the identity the triplet kernel is
    tested against."""
    F = cfg["hidden_channels"]
    batch, dist, unit, emb, x, vec = _prologue(sd, cfg, pos, Z, edge_index, cell, edge_shift, batch, pbc_shift)
    N = x.size(0)
    zs, zd = Z[edge_index[0]], Z[edge_index[1]]
    for layer in range(cfg["num_layers"]):
        x_new, vec_new = torch.zeros_like(x), torch.zeros_like(vec)
        for (t, a, c) in htnet_triads(cfg["elems"]):
            zt, za, zc = atomic_numbers[t], atomic_numbers[a], atomic_numbers[c]
            rows = (Z == zt).to(x.dtype)
            sel = torch.where((zd == zt) & ((zs == za) | (zs == zc)))[0]
            sel = sel[torch.sort(edge_index[1][sel], stable=True).indices]
            if sel.numel() == 0:
                continue
            pre = f"hermconvs.{layer}.mods.{triad_key(a, t, c)}."
            src, dst = edge_index[0][sel], edge_index[1][sel]
            m_x, m_vec = painn_message_terms(x, vec, src, emb[sel], unit[sel], sd, pre + "message_layer.", F)
            is_a = (zs[sel] == za)
            is_c = (zs[sel] == zc)
            p_a = scatter_rows(m_vec * is_a[:, None, None].to(x.dtype), dst, N)
            p_c = scatter_rows(m_vec * is_c[:, None, None].to(x.dtype), dst, N)
            if explicit_triplets:
                dots = torch.zeros(N, F, dtype=x.dtype)
                ea = torch.where(is_a)[0]
                ec = torch.where(is_c)[0]
                for e1 in ea.tolist():
                    same = ec[dst[ec] == dst[e1]]
                    if same.numel():
                        dots[dst[e1]] = dots[dst[e1]] + (m_vec[e1].unsqueeze(0) * m_vec[same]).sum(dim=(0, 1))
            else:
                dots = (p_a * p_c).sum(dim=1)
            n_a = torch.sqrt((p_a ** 2).sum(dim=1) + 1e-8)
            n_c = torch.sqrt((p_c ** 2).sum(dim=1) + 1e-8)
            vdot = dots / (n_a * n_c)
            dx = scatter_rows(m_x, dst, N)
            dvec = p_a if a == c else p_a + p_c
            xm = (x + dx) * (1 / math.sqrt(2.0))
            vm = vec + dvec
            dx2, dvec2 = painn_update(xm, vm, sd, pre + "update_layer.", F, vdot_override=vdot)
            vec_new = vec_new + (vm + dvec2) * rows[:, None, None]
            x_new = x_new + (xm + dx2) * rows[:, None]
        x, vec = x_new, vec_new
    energy = _readout(sd, cfg, x, batch, num_graphs)
    if return_features:
        return energy, x, vec
    return energy


FORWARDS = {"HVNet": hvnet_forward, "HPNet": hpnet_forward, "HTNet": htnet_forward}


# --------------------------------------------------------------------------------------
# state_dict construction (layout of SURVEY.md A.4) and consumers
# --------------------------------------------------------------------------------------
def module_names(kind: str, elems: Sequence[str]) -> List[str]:
    if kind == "HVNet":
        return list(elems)
    if kind == "HPNet":
        return [pair_key(s, d) for d in elems for s in elems]
    if kind == "HTNet":
        return [triad_key(a, t, c) for (t, a, c) in htnet_triads(list(elems))]
    raise ValueError(kind)


def state_dict_shapes(kind: str, cfg: dict) -> Dict[str, tuple]:
    """Key -> shape, in the reference's registration order (hermnet.py:95-116, rmnet.py:16-17,40-49,84-89)."""
    F, K = cfg["hidden_channels"], cfg["num_rbf"]
    shapes: Dict[str, tuple] = {"embed.weight": (len(atomic_numbers), F)}
    name = cfg.get("rbf", {"name": "gaussian"})["name"].lower()
    if name == "gaussian":
        shapes["radial_basis.rbf.offset"] = (K,)
    elif name == "spherical_bessel":
        shapes["radial_basis.rbf.frequencies"] = (K,)
    elif name == "bernstein":
        shapes["radial_basis.rbf.pregamma"] = ()
    for layer in range(cfg["num_layers"]):
        for m in module_names(kind, cfg["elems"]):
            p = f"hermconvs.{layer}.mods.{m}."
            shapes[p + "message_layer.x_proj.0.weight"] = (F, F)
            shapes[p + "message_layer.x_proj.0.bias"] = (F,)
            shapes[p + "message_layer.x_proj.2.weight"] = (3 * F, F)
            shapes[p + "message_layer.x_proj.2.bias"] = (3 * F,)
            shapes[p + "message_layer.rbf_proj.weight"] = (3 * F, K)
            shapes[p + "message_layer.rbf_proj.bias"] = (3 * F,)
            shapes[p + "message_layer.x_layernorm.weight"] = (F,)
            shapes[p + "message_layer.x_layernorm.bias"] = (F,)
            shapes[p + "update_layer.vec_proj.weight"] = (2 * F, F)
            shapes[p + "update_layer.xvec_proj.0.weight"] = (F, 2 * F)
            shapes[p + "update_layer.xvec_proj.0.bias"] = (F,)
            shapes[p + "update_layer.xvec_proj.2.weight"] = (3 * F, F)
            shapes[p + "update_layer.xvec_proj.2.bias"] = (3 * F,)
    shapes["out_energy.0.weight"] = (F // 2, F)
    shapes["out_energy.0.bias"] = (F // 2,)
    shapes["out_energy.2.weight"] = (1, F // 2)
    shapes["out_energy.2.bias"] = (1,)
    return shapes


def make_state_dict(kind: str, cfg: dict, seed: int = 1234, dtype=torch.float32) -> Dict[str, Tensor]:
    """Deterministic synthetic weights (numpy PCG64 stream, key order of ``state_dict_shapes``):
    the goldens store only inputs/outputs, both sides regenerate the weights from the seed."""
    import numpy as np
    rng = np.random.default_rng(seed)
    K = cfg["num_rbf"]
    sd: Dict[str, Tensor] = {}
    for key, shape in state_dict_shapes(kind, cfg).items():
        if key.endswith("rbf.offset"):
            val = torch.linspace(0.0, 1.0, K)                  # GaussianSmearing buffer
        elif key.endswith("rbf.frequencies"):
            val = math.pi * torch.arange(1, K + 1).float()
        elif key.endswith("rbf.pregamma"):
            val = torch.tensor(0.45264)
        elif key == "embed.weight":
            val = torch.from_numpy(rng.standard_normal(shape)).float()
        elif key.endswith("x_layernorm.weight"):
            val = torch.from_numpy(1.0 + 0.1 * rng.standard_normal(shape)).float()
        elif key.endswith(".bias"):
            val = torch.from_numpy(0.1 * rng.standard_normal(shape)).float()
        else:  # linear weights: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) like nn.Linear's default bound
            bound = 1.0 / math.sqrt(shape[-1])
            val = torch.from_numpy(rng.uniform(-bound, bound, size=shape)).float()
        sd[key] = val.to(dtype) if val.is_floating_point() else val
    return sd


def energy_and_forces(kind, sd, cfg, pos, Z, edge_index, cell=None, edge_shift=None, batch=None,
                      num_graphs=None, pbc_shift="reference", want_cell_grad=False, **kw):
    """E and F = -dE_total/dpos as the consumers compute them (calculator.py:75-83, dist_train.py:92-94)."""
    pos = pos.detach().clone().requires_grad_(True)
    inputs = [pos]
    if want_cell_grad:
        cell = cell.detach().clone().requires_grad_(True)
        inputs.append(cell)
    e = FORWARDS[kind](sd, cfg, pos, Z, edge_index, cell, edge_shift, batch, num_graphs, pbc_shift, **kw)
    grads = torch.autograd.grad(e.sum(), inputs)
    if want_cell_grad:
        return e.detach(), -grads[0], grads[1]
    return e.detach(), -grads[0]


def virial_calc(cell, pos, forces, energy, units="metal", pbc=False):
    """utils.py:138-160."""
    table = {"metal": 1.6021765e6, "real": 68568.415, "electron": 2.94210108e13}
    if units in table:
        nktv2p = table[units]
    elif units in ("lj", "si", "cgs", "micro", "nano"):
        nktv2p = 1.0
    else:
        raise ValueError("Illegal units command")
    if pbc:
        assert cell.requires_grad
        vir = torch.einsum("ij,ik->jk", pos, forces) - cell.T @ torch.autograd.grad(energy, cell)[0]
        return (vir + vir.T) / 2 * nktv2p
    vir = torch.einsum("ij,ik->jk", pos, forces) * nktv2p
    return (vir + vir.T) / 2
