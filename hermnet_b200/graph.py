"""Device-resident graph structures for the hot path.

``RowGraph`` replaces, in one object built once per configuration, what the reference rebuilds per layer and
per element with ``in_subgraph`` (HermNet/utils.py:11-24, called from hermnet.py:53-54): the grouping of edges by
destination (and sub-network), in HBM as a row CSR plus its transpose.

Internal atom order: atoms are stably sorted by element-type index (position in the model's ``elems``; unknown
elements last), so that the rows of one sub-network are a contiguous slice and the node-side GEMMs run on
slices instead of gather/scatter.  ``perm[k]`` is the original index of internal atom ``k``.

Row layouts (``rows_per_atom`` rows per destination atom, ``row_mod`` = weight-set id or -1):
  * HVNet  (vertex):  1 row per atom; module = element of the destination            (hermnet.py:51-61)
  * HPNet  (pair):    one row per source element; module = (src element -> dst element)   
  * HTNet  (triad):   two rows per unordered source-element pair {A,C}: edges from A, edges from C;
                      module = (dst element, {A,C})                                        
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from . import ops
from .symbols import atomic_numbers

Tensor = torch.Tensor


@dataclass
class Segments:
    """A partition of ``n_items`` items into ``n_rows`` rows: ``index[item] = row``; items of row r are
    ``perm[rowptr[r]:rowptr[r+1]]`` (``perm=None``: items are already grouped)."""
    index: Tensor
    rowptr: Tensor
    perm: Optional[Tensor]
    n_rows: int

    @staticmethod
    def from_index(index: Tensor, n_rows: int) -> "Segments":
        rowptr, order = ops.sort_by_key(index.contiguous(), n_rows)
        return Segments(index.contiguous(), rowptr, order, n_rows)


def pair_list(n_types: int):
    return [(a, c) for a in range(n_types) for c in range(a, n_types)]


class RowGraph:
    def __init__(self):
        self.kind = "HVNet"
        self.n_atoms = 0
        self.n_rows = 0
        self.rows_per_atom = 1
        self.n_edges = 0
        self.n_modules = 0
        self.n_graphs = 1
        self.sign = 1.0
        self.rowptr = self.col = self.shift = self.row_mod = None
        self.edge_row = self.t_rowptr = self.t_eid = None
        self.atom_graph = None        # int32 [N], graph id of each internal atom
        self.perm = self.inv_perm = None
        self.types = None             # int32 [N] element-type index of each internal atom (sorted ascending)
        self.type_ptr: List[int] = []  # host offsets of the type slices, len T+2 (last slice = unknown elements)
        self.row_xoff = None          # int64 [R]: row offset of the row's sub-network block in the flat xh buffer
        self.xh_sources, self.xh_base = [], [0]
        self.mod_active_host: List[bool] = []
        self.mod_active = None        # float [M]: 1 if the sub-network has at least one edge (hermnet.py:56-57)
        self.own_count: List[int] = []  # per type: number of OWNED atoms (they come first inside the type slice);
        #                                 the rest of the slice are ghost atoms of a domain-decomposed system
        self.energy_index = None      # int32 [N]: graph id of owned atoms, n_graphs for ghosts
        self.masked = False           # True: a Verlet-skin SUPERSET list (searched with rc + skin); the kernels drop d >= rc
        self.list_rc = None           # search radius of the list
        self._lazy = {}

    # ---- lazily built segment views (only the differentiable / training formulation needs them) ----------
    @property
    def seg_dst(self) -> Segments:      # row-edges grouped by row
        return Segments(self.edge_row, self.rowptr, None, self.n_rows)

    @property
    def seg_src(self) -> Segments:      # row-edges grouped by source atom
        return Segments(self.col, self.t_rowptr, self.t_eid, self.n_atoms)

    @property
    def seg_dst_atom(self) -> Segments:  # row-edges grouped by destination atom
        if "dst_atom" not in self._lazy:
            idx = torch.div(self.edge_row, self.rows_per_atom, rounding_mode="floor").to(torch.int32)
            self._lazy["dst_atom"] = Segments(idx, self.rowptr[:: self.rows_per_atom].contiguous(), None, self.n_atoms)
        return self._lazy["dst_atom"]

    @property
    def seg_batch(self) -> Segments:    # internal atoms grouped by graph (ghost atoms fall into the extra last row)
        if "batch" not in self._lazy:
            self._lazy["batch"] = Segments.from_index(self.energy_index, self.n_graphs + 1)
        return self._lazy["batch"]

    @property
    def edge_mod(self) -> Tensor:       # module of each row-edge
        if "edge_mod" not in self._lazy:
            self._lazy["edge_mod"] = self.row_mod[self.edge_row.long()]
        return self._lazy["edge_mod"]

    @property
    def seg_xh(self) -> Segments:       # row-edges grouped by (module, source atom): adjoint of the xh gather
        if "xh" not in self._lazy:
            idx = (self.row_xoff[self.edge_row.long()] + self.col.long()).clamp(min=0).to(torch.int32)
            self._lazy["xh"] = Segments.from_index(idx, self.xh_base[-1])
        return self._lazy["xh"]

    @property
    def seg_edge_graph(self) -> Segments:  # row-edges grouped by the graph of their source atom (cell gather)
        if "edge_graph" not in self._lazy:
            self._lazy["edge_graph"] = Segments.from_index(self.atom_graph[self.col.long()].contiguous(), self.n_graphs)
        return self._lazy["edge_graph"]

    def module_edges(self, m: int) -> Tensor:
        key = ("mod_edges", m)
        if key not in self._lazy:
            self._lazy[key] = torch.nonzero(self.edge_mod == m).squeeze(1)
        return self._lazy[key]

    def type_slice(self, t: int) -> slice:
        """All local atoms (owned, then ghost) of element type ``t``."""
        return slice(self.type_ptr[t], self.type_ptr[t + 1])

    def dst_slice(self, t: int) -> slice:
        """The owned atoms of type ``t`` -- the rows a sub-network with destination ``t`` updates."""
        return slice(self.type_ptr[t], self.type_ptr[t] + self.own_count[t])

    @property
    def n_ghost(self) -> int:
        return self.n_atoms - sum(self.own_count)


class GraphBuilder:
    """Builds ``RowGraph`` objects for one model (kind + ``elems``)."""

    def __init__(self, kind: str, elems: Sequence[str], rc: float, pbc_shift: str = "reference",
                 num_rbf: Optional[int] = None, hidden: Optional[int] = None):
        self.num_rbf, self.hidden = num_rbf, hidden
        # (environment switch: A/B measurements only)
        self.spatial_sort = os.environ.get("HERMNET_B200_SPATIAL", "1") != "0"   # Morton order inside every type slice
        if kind not in ("HVNet", "HPNet", "HTNet"):
            raise ValueError(kind)
        if pbc_shift not in ("reference", "physical"):
            raise ValueError("pbc_shift must be 'reference' or 'physical'")
        self.kind, self.elems, self.rc = kind, list(elems), float(rc)
        self.sign = 1.0 if pbc_shift == "reference" else -1.0
        self.T = len(self.elems)
        z2t = torch.full((len(atomic_numbers),), self.T, dtype=torch.int32)
        for t, el in enumerate(self.elems):
            z2t[atomic_numbers[el]] = t
        self._z2t_cpu = z2t
        self._z2t = {}
        self.pairs = pair_list(self.T)
        if kind == "HVNet":
            self.n_modules, self.n_groups = self.T, 1
        elif kind == "HPNet":
            self.n_modules, self.n_groups = self.T * self.T, self.T + 1
        else:
            self.n_modules, self.n_groups = self.T * len(self.pairs), self.T + 1

    def z2t(self, dev) -> Tensor:
        if dev not in self._z2t:
            self._z2t[dev] = self._z2t_cpu.to(dev)
        return self._z2t[dev]

    # ---------------------------------------------------------------------------------------------------
    @staticmethod
    def _morton(pos: Tensor, cell: Optional[Tensor], batch: Optional[Tensor]) -> Tensor:
        """30-bit Morton (Z-order) code of every atom: 10 bits per axis of the fractional (periodic) or bounding-box
        (open) coordinate.  Only used as a sort key, so that atoms that are close in space are close in memory and the
        feature rows gathered by neighbouring destination rows share L2 lines."""
        p = pos.detach().to(torch.float32)
        if cell is not None:
            c = cell.detach().to(torch.float32).reshape(-1, 3, 3)
            inv = torch.linalg.inv(c)
            f = torch.einsum("ni,nij->nj", p, inv[batch.long()] if (batch is not None and c.size(0) > 1) else
                             inv[:1].expand(p.size(0), 3, 3))
            f = f - torch.floor(f)
        else:
            lo = p.min(0).values if p.numel() else p.new_zeros(3)
            span = ((p.max(0).values - lo) if p.numel() else p.new_ones(3)).clamp(min=1e-6)
            f = (p - lo) / (span * 1.0001)
        q = (f * 1024.0).long().clamp_(0, 1023)

        def spread(v):      # 10 bits -> every third bit
            v = (v | (v << 16)) & 0x030000FF
            v = (v | (v << 8)) & 0x0300F00F
            v = (v | (v << 4)) & 0x030C30C3
            v = (v | (v << 2)) & 0x09249249
            return v

        return spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)

    def _order(self, Z: Tensor, owned: Optional[Tensor] = None, pos: Optional[Tensor] = None,
               cell: Optional[Tensor] = None, batch: Optional[Tensor] = None):
        """Internal order: by element type, owned atoms before ghost atoms inside a type, Morton order inside that."""
        types = self.z2t(Z.device)[Z.long()]
        major = types.long() * 2 + (0 if owned is None else (~owned).long())
        if pos is not None and self.spatial_sort and Z.numel() > 0:
            key_sorted, perm = torch.sort((major << 30) | self._morton(pos, cell, batch), stable=True)
            key_sorted = key_sorted >> 30
        else:
            key_sorted, perm = torch.sort(major, stable=True)
        types_sorted = torch.div(key_sorted, 2, rounding_mode="floor")
        counts = torch.bincount(key_sorted, minlength=2 * (self.T + 1)).view(-1, 2)
        c = counts.tolist()                                          # one host sync per graph build
        type_ptr = [0]
        for t in range(self.T + 1):
            type_ptr.append(type_ptr[-1] + c[t][0] + c[t][1])
        self._own_count = [c[t][0] for t in range(self.T + 1)]
        inv = torch.empty_like(perm)
        inv[perm] = torch.arange(perm.numel(), device=perm.device)
        return types_sorted.to(torch.int32).contiguous(), perm, inv, type_ptr

    def from_positions(self, pos: Tensor, Z: Tensor, cell: Optional[Tensor], batch: Optional[Tensor],
                       max_neighbors: int = 0, skin: float = 0.0) -> RowGraph:
        """Native path: cell-list radius graph on the device (replaces data.py:14-24 + utils.py:11-24).
        ``skin > 0``: Verlet list -- the search radius is ``rc + skin`` and the graph is marked ``masked``: it stays valid
        while no atom has moved by more than ``skin / 2``, the edge kernels ignore entries with ``d >= rc``."""
        g = self._from_positions(pos, Z, cell, batch, max_neighbors, float(skin))
        g.masked, g.list_rc = skin > 0.0, self.rc + float(skin)
        return g

    def _from_positions(self, pos, Z, cell, batch, max_neighbors, skin) -> RowGraph:
        rc_list = self.rc + skin
        if skin > 0.0 and (cell is None or max_neighbors):
            raise ValueError("Verlet-skin lists need a periodic cell (the non-periodic branch caps at 32 neighbours)")
        dev = pos.device
        n = pos.size(0)
        n_graphs = 1 if batch is None else (int(batch.max().item()) + 1 if n else 1)
        types, perm, inv, type_ptr = self._order(Z, None, pos, cell, batch)
        pos32 = pos.detach().to(torch.float32)
        cell32 = None if cell is None else cell.detach().to(torch.float32).reshape(-1, 3, 3).contiguous()
        if cell32 is None and max_neighbors == 0:
            max_neighbors = 32   # torch_cluster default of the reference's non-periodic branch (data.py:16)
        if n_graphs == 1 and max_neighbors == 0:
            gptr = torch.tensor([0, n], dtype=torch.int32, device=dev)
            rowptr, col, shift = ops.radius_graph(pos32[perm].contiguous(), cell32, gptr, rc_list,
                                                  types if self.n_groups > 1 else None, self.n_groups, 0)
            shift = -shift    # centre-row entry (j, S') is the reference edge (src=j, dst=centre, S=-S')
            return self._finish(n, n_graphs, rowptr, col, shift, types, perm, inv, type_ptr,
                                torch.zeros(n, dtype=torch.int32, device=dev), pos_i=pos32[perm].contiguous(), cell=cell32)
        # batches / capped lists: search in the original (graph-contiguous) order, then regroup
        b = torch.zeros(n, dtype=torch.long, device=dev) if batch is None else batch.long()
        gptr = torch.zeros(n_graphs + 1, dtype=torch.int32, device=dev)
        gptr[1:] = torch.cumsum(torch.bincount(b, minlength=n_graphs), 0)
        rowptr, col, shift = ops.radius_graph(pos32.contiguous(), cell32, gptr, rc_list, None, 1, max_neighbors)
        centre = ops.expand_rowptr(rowptr, col.numel())
        return self.from_coo(n, src=col.long(), dst=centre.long(), shift=-shift, Z=Z, batch=batch,
                             order=(types, perm, inv, type_ptr), pos=pos32, cell=cell32)

    def from_local_positions(self, pos: Tensor, Z: Tensor, cell: Tensor, owned: Tensor) -> RowGraph:
        """Domain-decomposed system: ``pos/Z`` are the rank's local atoms (owned + halo candidates) inside the FULL
        periodic ``cell``; rows are kept for owned destinations only, every local atom may be a source."""
        dev = pos.device
        n = pos.size(0)
        types, perm, inv, type_ptr = self._order(Z, owned, pos, cell, None)
        own_count = list(self._own_count)
        pos32 = pos.detach().to(torch.float32)
        cell32 = cell.detach().to(torch.float32).reshape(-1, 3, 3).contiguous()
        gptr = torch.tensor([0, n], dtype=torch.int32, device=dev)
        G = self.n_groups
        rowptr, col, shift = ops.radius_graph(pos32[perm].contiguous(), cell32, gptr, self.rc,
                                              types if G > 1 else None, G, 0)
        owned_i = owned[perm]
        lens = (rowptr[1:] - rowptr[:-1]).view(n, G) * owned_i.view(n, 1).to(torch.int32)
        centre = torch.div(ops.expand_rowptr(rowptr, col.numel()).long(), G, rounding_mode="floor")
        keep = owned_i[centre]
        new_rowptr = torch.zeros(n * G + 1, dtype=torch.int32, device=dev)
        new_rowptr[1:] = torch.cumsum(lens.reshape(-1), 0)
        return self._finish(n, 1, new_rowptr, col[keep].contiguous(), (-shift[keep]).contiguous(), types, perm, inv,
                            type_ptr, torch.zeros(n, dtype=torch.int32, device=dev), owned_i, own_count,
                            pos_i=pos32[perm].contiguous(), cell=cell32)

    def from_edge_index(self, Z: Tensor, edge_index: Tensor, edge_shift: Optional[Tensor], batch: Optional[Tensor],
                        pos: Optional[Tensor] = None, cell: Optional[Tensor] = None) -> RowGraph:
        """General path for a user-supplied reference-format ``edge_index`` (row 0 = source, row 1 = destination)
        and ``edge_shift``: one stable device sort instead of the reference's per-atom scans."""
        shift = None
        if edge_shift is not None:
            s = torch.round(edge_shift).to(torch.int8)
            shift = torch.cat([s, torch.zeros((s.size(0), 1), dtype=torch.int8, device=s.device)], 1)
        pos32 = None if pos is None else pos.detach().to(torch.float32)
        cell32 = None if (cell is None or edge_shift is None) else cell.detach().to(torch.float32).reshape(-1, 3, 3).contiguous()
        return self.from_coo(Z.numel(), edge_index[0].long(), edge_index[1].long(), shift, Z, batch, pos=pos32, cell=cell32)

    def from_coo(self, n: int, src: Tensor, dst: Tensor, shift: Optional[Tensor], Z: Tensor, batch: Optional[Tensor],
                 order=None, pos: Optional[Tensor] = None, cell: Optional[Tensor] = None) -> RowGraph:
        dev = Z.device
        types, perm, inv, type_ptr = order if order is not None else self._order(Z, None, pos, cell, batch)
        n_graphs = 1 if batch is None else (int(batch.max().item()) + 1 if n else 1)
        src_i, dst_i = inv[src], inv[dst]
        G = self.n_groups
        key = dst_i * G + (types[src_i].long() if G > 1 else 0)
        rowptr, order_e = ops.sort_by_key(key.to(torch.int32).contiguous(), n * G)
        col = src_i[order_e.long()].to(torch.int32).contiguous()
        if shift is None:
            shift = torch.zeros((col.numel(), 4), dtype=torch.int8, device=dev)
        else:
            shift = shift[order_e.long()].contiguous()
        atom_graph = (torch.zeros(n, dtype=torch.int32, device=dev) if batch is None
                      else batch[perm].to(torch.int32).contiguous())
        return self._finish(n, n_graphs, rowptr, col, shift, types, perm, inv, type_ptr, atom_graph,
                            pos_i=None if pos is None else pos[perm].contiguous(), cell=cell)

    # ---------------------------------------------------------------------------------------------------
    def _distance_bins(self, pos_i, cell, atom_graph, rowptr, col, shift, rows_per_atom) -> Tensor:
        """8-bit distance bin of every entry of a CSR (positions in internal order) -- the sort key that makes the
        Gaussian bands of consecutive row entries overlap (csrc/hn_edge.cu)."""
        from types import SimpleNamespace
        e = int(col.numel())
        tmp = SimpleNamespace(atom_graph=atom_graph, edge_row=ops.expand_rowptr(rowptr, e), rows_per_atom=rows_per_atom,
                              col=col, shift=shift, sign=self.sign, n_edges=e)
        d = ops.edge_geom_fwd(pos_i, cell, tmp)[:, 3]
        return torch.clamp(d * (255.0 / self.rc), max=255.0).to(torch.int32).contiguous(), tmp.edge_row

    def _finish(self, n, n_graphs, rowptr, col, shift, types, perm, inv, type_ptr, atom_graph, owned=None,
                own_count=None, pos_i=None, cell=None) -> RowGraph:
        """``rowptr/col/shift`` is the base CSR with ``n_groups`` rows per atom (source-element groups)."""
        dev = col.device
        dbin = None
        # rows sorted by distance: only the FMA-pipe edge kernels (F != 128) sweep consecutive row entries with overlapping
        # Gaussian bands; the tensor-core kernels regroup the edges in their tile plans, so the three sort passes are skipped
        want_dbin = not (self.hidden is not None and self.num_rbf is not None and ops.edge_use_tc(self.hidden, self.num_rbf))
        if want_dbin and pos_i is not None and col.numel() > 0:
            # sort every base row by distance (stable two-pass: distance bin, then row)
            dbin, base_row = self._distance_bins(pos_i, cell, atom_graph, rowptr, col, shift, self.n_groups)
            o1 = ops.sort_by_key(dbin, 256)[1].long()
            o2 = ops.sort_by_key(base_row[o1].contiguous(), n * self.n_groups)[1].long()
            order = o1[o2]
            col, shift, dbin = col[order].contiguous(), shift[order].contiguous(), dbin[order].contiguous()
        g = RowGraph()
        g.own_count = list(own_count) if own_count is not None else list(self._own_count)
        g.energy_index = atom_graph if owned is None else torch.where(
            owned, atom_graph, torch.full_like(atom_graph, n_graphs)).contiguous()
        g.kind, g.n_atoms, g.n_graphs, g.sign = self.kind, n, n_graphs, self.sign
        g.perm, g.inv_perm, g.types, g.type_ptr, g.atom_graph = perm, inv, types, type_ptr, atom_graph
        g.n_modules = self.n_modules
        T, G = self.T, self.n_groups
        t_long = types.long()
        known = t_long < T
        if owned is not None:
            known = known & owned           # ghost atoms are sources only: their rows are inactive
        if self.kind == "HVNet":
            g.rows_per_atom = 1
            g.row_mod = torch.where(known, t_long, torch.full_like(t_long, -1)).to(torch.int32)
        elif self.kind == "HPNet":
            g.rows_per_atom = G
            s = torch.arange(G, device=dev).view(1, G).expand(n, G)
            mod = t_long.view(n, 1) * T + s                      # module (src s -> dst t) = t*T + s
            ok = known.view(n, 1) & (s < T)
            g.row_mod = torch.where(ok, mod, torch.full_like(mod, -1)).to(torch.int32).reshape(-1)
        else:  # HTNet: replicate base rows (i, A) / (i, C) for every pair {A, C}
            P = len(self.pairs)
            g.rows_per_atom = 2 * P
            sel = torch.tensor([[a, c] for (a, c) in self.pairs], device=dev).reshape(-1)         # [2P] source type
            live = torch.tensor([[1, 0 if a == c else 1] for (a, c) in self.pairs], device=dev).reshape(-1)
            base_len = (rowptr[1:] - rowptr[:-1]).view(n, G).long()
            new_len = base_len[:, sel] * live.view(1, -1)                                          # [n, 2P]
            new_rowptr = torch.zeros(n * 2 * P + 1, dtype=torch.int32, device=dev)
            new_rowptr[1:] = torch.cumsum(new_len.reshape(-1), 0)
            e_new = int(new_rowptr[-1].item())
            new_edge_row = ops.expand_rowptr(new_rowptr, e_new).long()
            atom_of = torch.div(new_edge_row, 2 * P, rounding_mode="floor")
            slot = new_edge_row - atom_of * 2 * P
            base_row = atom_of * G + sel[slot]
            base_e = rowptr.long()[base_row] + (torch.arange(e_new, device=dev) - new_rowptr.long()[new_edge_row])
            col, shift, rowptr = col[base_e].contiguous(), shift[base_e].contiguous(), new_rowptr
            if dbin is not None:
                dbin = dbin[base_e].contiguous()
            pidx = torch.arange(P, device=dev).repeat_interleave(2).view(1, -1).expand(n, 2 * P)
            mod = t_long.view(n, 1) * P + pidx
            ok = known.view(n, 1) & (live.view(1, -1) > 0)
            g.row_mod = torch.where(ok, mod, torch.full_like(mod, -1)).to(torch.int32).reshape(-1)
        g.rowptr, g.col, g.shift = rowptr.contiguous(), col, shift
        # compact layout of the projected source features: module m owns rows [xh_base[m], xh_base[m+1]) of one
        # flat [rows, 3F] buffer, holding x_proj_m(LN_m(x)) for the source-type slices listed in xh_sources[m]
        tp = type_ptr
        g.xh_sources, g.xh_base = [], [0]
        table = torch.zeros((T + 1, g.rows_per_atom), dtype=torch.long)
        for m in range(self.n_modules):
            if self.kind == "HVNet":
                srcs = [(0, n)]
                table[m, 0] = g.xh_base[-1]
            elif self.kind == "HPNet":
                t, s = divmod(m, T)
                srcs = [(tp[s], tp[s + 1])]
                table[t, s] = g.xh_base[-1] - tp[s]
            else:
                t, pi = divmod(m, len(self.pairs))
                a, c = self.pairs[pi]
                srcs = [(tp[a], tp[a + 1])] + ([] if a == c else [(tp[c], tp[c + 1])])
                table[t, 2 * pi] = g.xh_base[-1] - tp[a]
                table[t, 2 * pi + 1] = g.xh_base[-1] + (tp[a + 1] - tp[a]) - tp[c]
            g.xh_sources.append(srcs)
            g.xh_base.append(g.xh_base[-1] + sum(hi - lo for lo, hi in srcs))
        g.row_xoff = table.to(dev)[t_long].reshape(-1).contiguous()
        g.n_rows = n * g.rows_per_atom
        g.n_edges = int(col.numel())
        g.row_mod = g.row_mod.contiguous()
        g.edge_row = ops.expand_rowptr(g.rowptr, g.n_edges)
        # transposed view, every source row sorted by (module, distance) so that tiles of consecutive entries share
        # the module's filter rows (edge_bwd_src_kernel)
        M1 = self.n_modules + 1
        major = (g.col.long() * M1 + (g.row_mod[g.edge_row.long()].long() + 1)).to(torch.int32).contiguous()
        if dbin is not None and g.n_edges > 0:
            o1 = ops.sort_by_key(dbin, 256)[1].long()
            rp, o2 = ops.sort_by_key(major[o1].contiguous(), n * M1)
            g.t_eid = o1[o2.long()].to(torch.int32).contiguous()
        else:
            rp, g.t_eid = ops.sort_by_key(major, n * M1)
        g.t_rowptr = rp[::M1].contiguous()
        lens = (g.rowptr[1:] - g.rowptr[:-1]).long()
        cnt = torch.zeros(self.n_modules + 1, dtype=torch.long, device=dev)
        cnt.index_add_(0, torch.where(g.row_mod >= 0, g.row_mod.long(), torch.full_like(lens, self.n_modules)), lens)
        g.mod_active = (cnt[: self.n_modules] > 0).to(torch.float32)
        g.mod_active_host = (cnt[: self.n_modules] > 0).tolist()   # (graph build already synchronises)
        return g
