"""Stand-ins for the reference's third-party imports (TEST INFRASTRUCTURE, authoring container only).

``install()`` registers minimal ``torch_geometric``, ``torch_scatter``, ``torch_cluster`` and ``ase``
modules in ``sys.modules`` so that the reference's OWN, UNMODIFIED source files
(``/root/reference/HermNet/{hermnet,rmnet,utils,data}.py``) can be imported and executed here, which is
how the oracle is pinned and how ``tests/golden`` fixtures are produced (``tests/golden/make_golden.py``).
Nothing here is shipped to, or needed on, the GPU box.

Each shim restates the published semantics of the call the reference makes -- all of them
"[upstream, unverified here]" (SURVEY.md 8c): PyG ``Data`` / ``MessagePassing.propagate`` /
``GaussianSmearing``, ``torch_scatter.scatter``, torch_cluster ``radius_graph``, ASE
``primitive_neighbor_list`` and ``ase.data.atomic_numbers``.
"""
from __future__ import annotations

import copy
import inspect
import sys
import types

import numpy as np
import torch

from . import neighbor_oracle
from .hermnet_oracle import scatter_rows
from .symbols import atomic_numbers, chemical_symbols

_NODE_KEYS = {"x", "feat", "pos", "batch", "node_type", "n_id", "tensor", "atomic_number", "vec"}


class Data:
    """Duck of ``torch_geometric.data.Data`` covering what hermnet.py / utils.py touch."""

    def __init__(self, **kwargs):
        object.__setattr__(self, "_store", {})
        for k, v in kwargs.items():
            self._store[k] = v

    # attribute / item access -----------------------------------------------------------------
    def __getattr__(self, key):
        store = object.__getattribute__(self, "_store")
        if key in store:
            return store[key]
        raise AttributeError(key)

    def __setattr__(self, key, value):
        self._store[key] = value

    def __getitem__(self, key):
        return self._store[key]

    def __setitem__(self, key, value):
        self._store[key] = value

    def get(self, key, default=None):
        return self._store.get(key, default)

    def __iter__(self):
        return iter(list(self._store.items()))

    def keys(self):
        return list(self._store.keys())

    def __copy__(self):
        out = Data()
        out._store.update(self._store)
        return out

    # sizes -------------------------------------------------------------------------------------
    @property
    def num_edges(self):
        ei = self._store.get("edge_index")
        return 0 if ei is None else int(ei.size(1))

    @property
    def num_nodes(self):
        for k in ("x", "pos", "atomic_number", "batch"):
            v = self._store.get(k)
            if isinstance(v, torch.Tensor):
                return int(v.size(0))
        return None

    def is_edge_attr(self, key):
        value = self._store[key]
        if not isinstance(value, torch.Tensor) or value.dim() == 0:
            return False
        cat_dim = -1 if "index" in key else 0
        if value.size(cat_dim) != self.num_edges:
            return False
        if self.num_nodes != self.num_edges:
            return True
        return "edge" in key

    def to(self, device):
        out = Data()
        for k, v in self._store.items():
            out._store[k] = v.to(device) if isinstance(v, torch.Tensor) else v
        return out


class MessagePassing(torch.nn.Module):
    """``propagate``: ``*_j`` args gathered with ``edge_index[0]``, ``*_i`` with ``edge_index[1]``;
    aggregation index ``edge_index[1]``; ``dim_size`` = number of nodes."""

    def __init__(self, aggr="add", node_dim=0, **kw):
        super().__init__()
        self.aggr, self.node_dim = aggr, node_dim

    def propagate(self, edge_index, size=None, **kwargs):
        n = None
        for v in kwargs.values():
            if isinstance(v, torch.Tensor) and v.size(0) != edge_index.size(1):
                n = v.size(0) if n is None else n
        msg_args = {}
        for name in inspect.signature(self.message).parameters:
            if name.endswith("_j"):
                msg_args[name] = kwargs[name[:-2]].index_select(self.node_dim, edge_index[0])
            elif name.endswith("_i"):
                msg_args[name] = kwargs[name[:-2]].index_select(self.node_dim, edge_index[1])
            else:
                msg_args[name] = kwargs[name]
        out = self.message(**msg_args)
        out = self.aggregate(out, index=edge_index[1], ptr=None, dim_size=n)
        return self.update(out)


class GaussianSmearing(torch.nn.Module):
    def __init__(self, start=0.0, stop=5.0, num_gaussians=50):
        super().__init__()
        offset = torch.linspace(start, stop, num_gaussians)
        self.coeff = -0.5 / (offset[1] - offset[0]).item() ** 2
        self.register_buffer("offset", offset)

    def forward(self, dist):
        dist = dist.view(-1, 1) - self.offset.view(1, -1)
        return torch.exp(self.coeff * torch.pow(dist, 2))


def scatter(src, index, dim=0, dim_size=None, reduce="sum"):
    assert dim == 0
    if dim_size is None:
        dim_size = int(index.max().item()) + 1 if index.numel() else 0
    return scatter_rows(src, index, dim_size, "sum" if reduce in ("sum", "add") else reduce)


def radius_graph(x, r, batch=None, loop=False, max_num_neighbors=32, flow="source_to_target"):
    ei = neighbor_oracle.radius_graph_nonpbc(x.detach().numpy(), r, max_num_neighbors,
                                             None if batch is None else batch.numpy())
    return torch.from_numpy(ei)


def primitive_neighbor_list(quantities, pbc, cell, positions, cutoff, **kw):
    assert quantities == "ijS" and all(pbc)
    i, j, S = neighbor_oracle.neighbor_list_pbc(np.asarray(positions), np.asarray(cell), float(cutoff))
    return i, j, S


def install():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class InMemoryDataset:  # only needs to exist for ``class BaseDataModule(InMemoryDataset)``
        def __init__(self, *a, **k):
            pass

    class Calculator:
        def __init__(self, *a, **k):
            self.results = {}

        def calculate(self, atoms=None, properties=None, system_changes=None):
            self.atoms = atoms

    mod("torch_scatter", scatter=scatter)
    mod("torch_cluster", radius_graph=radius_graph)
    tg = mod("torch_geometric")
    tg.data = mod("torch_geometric.data", Data=Data, InMemoryDataset=InMemoryDataset,
                  download_url=lambda *a, **k: None)
    tg.nn = mod("torch_geometric.nn", MessagePassing=MessagePassing, radius_graph=radius_graph)
    tg.nn.models = mod("torch_geometric.nn.models")
    tg.nn.models.schnet = mod("torch_geometric.nn.models.schnet", GaussianSmearing=GaussianSmearing)
    tg.loader = mod("torch_geometric.loader", DataLoader=None)
    ase = mod("ase")
    ase.data = mod("ase.data", atomic_numbers=atomic_numbers, chemical_symbols=chemical_symbols)
    ase.units = mod("ase.units", kcal=2.611447418269555e22 / 6.02214076e23 * 1.0, mol=6.02214076e23)
    ase.neighborlist = mod("ase.neighborlist", primitive_neighbor_list=primitive_neighbor_list)
    ase.calculators = mod("ase.calculators")
    ase.calculators.calculator = mod("ase.calculators.calculator", Calculator=Calculator,
                                     all_changes=["positions", "numbers", "cell", "pbc"])


def import_reference(root="/root/reference"):
    """Import the reference package (unmodified) over the shims; returns the ``HermNet`` module."""
    install()
    if root not in sys.path:
        sys.path.insert(0, root)
    import importlib
    return importlib.import_module("HermNet")
