"""Graph containers and neighbour search with the reference's call contract.

``neighbor_search`` / ``transform`` keep the signatures of ``/root/reference/HermNet/data.py:14-35``; the search
itself runs on the GPU (cell list, ``hermnet_b200/csrc/hn_graph.cu``) instead of ASE / torch_cluster on the CPU.
``Data`` / ``Batch`` / ``DataLoader`` are PyG-free duck types of the torch_geometric classes the reference's
callers use: attribute and item access, ``get``, iteration over ``(key, value)``,
``is_edge_attr``, ``num_edges``, ``to``; collation concatenates node / edge tensors, offsets ``edge_index`` by the
cumulative node count, builds ``batch`` and stacks ``cell`` on dim 0.  A real PyG object also works.
The file-backed datasets of data.py:38-247 are out of scope (file IO / downloads).
"""
from __future__ import annotations

import copy
from typing import Callable, List, Optional, Sequence

import torch
from torch import Tensor

from . import ops

__all__ = ["neighbor_search", "transform", "transform_batch", "Data", "Batch", "DataLoader"]


def neighbor_search(pos: Tensor, rc: float, cell: Optional[Tensor] = None):
    """data.py:14-24.  Non-periodic: ``edge_index [2,E] int64`` (row 0 = neighbour, row 1 = centre, at most 32
    neighbours per centre).  Periodic (``cell [1,3,3]`` or ``[3,3]``): ``(edge_index, edge_shift [E,3] float32)``
    with ``edge_index = [i; j]``, ``|| pos_j - pos_i + S.cell || < rc`` -- sorted by ``edge_index[0]``.
    Inputs may live on the CPU (as in the reference) or on the GPU; outputs follow ``pos.device``."""
    dev = ops.compute_device(pos)
    p = pos.detach().to(dev, torch.float32).contiguous()
    n = p.size(0)
    gptr = torch.tensor([0, n], dtype=torch.int32, device=dev)
    if cell is None:
        rowptr, col, _ = ops.radius_graph(p, None, gptr, rc, None, 1, 32)
        centre = ops.expand_rowptr(rowptr, col.numel())
        return torch.stack([col.long(), centre.long()]).to(pos.device)
    c = cell.detach().to(dev, torch.float32).reshape(-1, 3, 3)[:1].contiguous()
    rowptr, col, shift = ops.radius_graph(p, c, gptr, rc, None, 1, 0)
    centre = ops.expand_rowptr(rowptr, col.numel())
    edge_index = torch.stack([centre.long(), col.long()]).to(pos.device)
    return edge_index, shift[:, :3].to(torch.float32).to(pos.device)


def transform(data, rc: float):
    """data.py:27-35."""
    assert data.pos is not None
    if data.get('cell') is None:
        data.edge_index = neighbor_search(data.pos, rc)
    else:
        data.edge_index, data.edge_shift = neighbor_search(data.pos, rc, data.cell)
    return data


def transform_batch(batch, rc: float):
    """Loader-side graph build for a COLLATED batch: one batched device neighbour search over all
    graphs of the batch (every graph in its own cell) instead of ``transform`` per sample in ``__getitem__``
    (data.py:58-67).  Attaches the same ``edge_index`` (and ``edge_shift``) the per-sample path followed by PyG collation
    would give, as a set; edges are grouped by graph and sorted by ``edge_index[0]``."""
    assert batch.pos is not None
    dev = ops.compute_device(batch.pos)
    p = batch.pos.detach().to(dev, torch.float32).contiguous()
    n = p.size(0)
    b = batch.get("batch")
    b = torch.zeros(n, dtype=torch.long, device=dev) if b is None else b.to(dev).long()
    n_graphs = int(b.max().item()) + 1 if n else 1
    if n and not bool((b[1:] >= b[:-1]).all()):
        raise ValueError("transform_batch: atoms must be grouped by graph (PyG collation order)")
    gptr = torch.zeros(n_graphs + 1, dtype=torch.int32, device=dev)
    gptr[1:] = torch.cumsum(torch.bincount(b, minlength=n_graphs), 0)
    cell = batch.get("cell")
    if cell is None:
        rowptr, col, _ = ops.radius_graph(p, None, gptr, rc, None, 1, 32)
        centre = ops.expand_rowptr(rowptr, col.numel())
        batch.edge_index = torch.stack([col.long(), centre.long()]).to(batch.pos.device)
        return batch
    c = cell.detach().to(dev, torch.float32).reshape(-1, 3, 3).contiguous()
    rowptr, col, shift = ops.radius_graph(p, c, gptr, rc, None, 1, 0)
    centre = ops.expand_rowptr(rowptr, col.numel())
    batch.edge_index = torch.stack([centre.long(), col.long()]).to(batch.pos.device)
    batch.edge_shift = shift[:, :3].to(torch.float32).to(batch.pos.device)
    return batch


class Data:
    """Minimal torch_geometric.data.Data duck type."""

    def __init__(self, **kwargs):
        object.__setattr__(self, "_store", dict(kwargs))

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

    def __contains__(self, key):
        return key in self._store

    def get(self, key, default=None):
        return self._store.get(key, default)

    def keys(self):
        return list(self._store.keys())

    def __iter__(self):
        return iter(list(self._store.items()))

    def __copy__(self):
        return self.__class__(**self._store)

    @property
    def num_edges(self) -> int:
        ei = self._store.get("edge_index")
        return 0 if ei is None else int(ei.size(1))

    @property
    def num_nodes(self) -> Optional[int]:
        for k in ("pos", "x", "atomic_number", "batch"):
            v = self._store.get(k)
            if isinstance(v, Tensor):
                return int(v.size(0))
        return None

    def is_node_attr(self, key) -> bool:
        v = self._store.get(key)
        return isinstance(v, Tensor) and v.dim() > 0 and key not in ("cell", "y") and "edge" not in key \
            and v.size(0) == self.num_nodes

    def is_edge_attr(self, key) -> bool:
        v = self._store.get(key)
        if not isinstance(v, Tensor) or v.dim() == 0:
            return False
        if v.size(-1 if "index" in key else 0) != self.num_edges:
            return False
        return True if self.num_nodes != self.num_edges else "edge" in key

    def to(self, device, non_blocking: bool = False):
        out = self.__class__()
        for k, v in self._store.items():
            if isinstance(v, Tensor):
                out._store[k] = v.to(device, non_blocking=non_blocking)
            elif k == "graph":
                continue          # device-resident graph handles are rebuilt, never moved
            else:
                out._store[k] = v
        return out

    def __repr__(self):
        items = ", ".join(f"{k}={list(v.shape) if isinstance(v, Tensor) else v!r}" for k, v in self._store.items())
        return f"{self.__class__.__name__}({items})"


class Batch(Data):
    """PyG ``Batch.from_data_list`` semantics for the fields the hot path uses."""

    @staticmethod
    def from_data_list(data_list: Sequence[Data]) -> "Batch":
        keys = data_list[0].keys()
        out, offset, batch = {k: [] for k in keys if k != "graph"}, 0, []
        for g, d in enumerate(data_list):
            n = d.num_nodes
            for k in out:
                v = d[k]
                if k == "edge_index":
                    v = v + offset
                elif k == "cell" and isinstance(v, Tensor):
                    v = v.reshape(-1, 3, 3)
                elif isinstance(v, Tensor) and v.dim() == 0:
                    v = v.reshape(1)
                elif not isinstance(v, Tensor):
                    v = torch.as_tensor(v).reshape(-1)
                out[k].append(v)
            batch.append(torch.full((n,), g, dtype=torch.long))
            offset += n
        b = Batch(**{k: torch.cat(v, dim=1 if k == "edge_index" else 0) for k, v in out.items()})
        b.batch = torch.cat(batch).to(b.pos.device)
        b._store["num_graphs"] = len(data_list)
        return b


class DataLoader:
    """Tiny stand-in for ``torch_geometric.loader.DataLoader`` (``batch_size``, ``shuffle``, ``sampler``)."""

    def __init__(self, dataset, batch_size: int = 1, shuffle: bool = False, sampler=None,
                 collate_fn: Optional[Callable[[List[Data]], Data]] = None, rc: Optional[float] = None, device=None):
        """``rc``: build the neighbour list of every collated batch with ONE batched device search (``transform_batch``)
        -- for datasets that do not run ``transform`` per sample; ``device``: move the batch there first."""
        self.dataset, self.batch_size, self.shuffle, self.sampler = dataset, batch_size, shuffle, sampler
        base = collate_fn or Batch.from_data_list
        if rc is None and device is None:
            self.collate_fn = base
        else:
            def collate(items):
                b = base(items)
                if device is not None:
                    b = b.to(device)
                return b if rc is None else transform_batch(b, rc)
            self.collate_fn = collate

    def _indices(self):
        if self.sampler is not None:
            return list(iter(self.sampler))
        if self.shuffle:
            return torch.randperm(len(self.dataset)).tolist()
        return list(range(len(self.dataset)))

    def __iter__(self):
        idx = self._indices()
        for i in range(0, len(idx), self.batch_size):
            yield self.collate_fn([self.dataset[j] for j in idx[i:i + self.batch_size]])

    def __len__(self):
        n = len(self.sampler) if self.sampler is not None else len(self.dataset)
        return (n + self.batch_size - 1) // self.batch_size
