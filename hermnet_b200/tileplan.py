"""Tile plans of the tensor-core edge kernels (``csrc/hn_edge_tc.cu``).

A plan regroups the row-edges of a ``RowGraph`` once per configuration -- what the reference redoes per layer and per
element with ``in_subgraph`` (HermNet/utils.py:11-24) -- into *blocks* (32 rows of one sub-network, or 32 consecutive
source atoms) and *tiles* (<= 64 edges of a block whose Gaussian bands fit one 32-wide window of the basis index), so
that the filter projection of a tile (rmnet.py:45,55) is ONE tensor-core tile  W[3F x 32] . basis[32 x 64].

``dst`` plans drive the forward and the destination-major backward, ``src`` plans the source-major backward.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from ._lib import TcPlan

Tensor = torch.Tensor


class TilePlan:
    def __init__(self, kind: str, blk_info, blk_tile, tile_info, erec, n_blocks: int, n_tiles: int, has_inactive: bool,
                 blk_xoff=None):
        self.kind = kind
        self.blk_info, self.blk_tile, self.tile_info, self.erec = blk_info, blk_tile, tile_info, erec
        # dst-major: the xh row of source s for the rows of block b is blk_xoff[b] + s
        self.blk_xoff = blk_xoff if blk_xoff is not None else torch.zeros(max(n_blocks, 1), dtype=torch.int32, device=erec.device)
        self.n_blocks, self.n_tiles, self.has_inactive = n_blocks, n_tiles, has_inactive
        self.tile_win = torch.zeros((max(n_tiles, 1), 2), dtype=torch.int32, device=erec.device)
        self.tile_geom = torch.zeros((erec.size(0), 4), dtype=torch.float32, device=erec.device)
        self.zero_row = torch.zeros(512, dtype=torch.float32, device=erec.device)     # staged for masked (d >= rc) edges
        self.window = 32           # basis-index width the tiles were cut for
        self.blk_order = None      # int32 [n_blocks] processing order (None: as stored)
        self.win_key = None        # (data_ptr, version) of the geometry the windows were computed for
        self._c = None

    def cstruct(self) -> TcPlan:
        if self._c is None:
            self._c = TcPlan(self.n_blocks, self.n_tiles, self.blk_info.data_ptr(), self.blk_tile.data_ptr(),
                             self.blk_xoff.data_ptr(), self.tile_info.data_ptr(), self.tile_win.data_ptr(),
                             self.erec.data_ptr(), self.tile_geom.data_ptr(), self.zero_row.data_ptr(),
                             None if self.blk_order is None else self.blk_order.data_ptr())
        return self._c

    def with_erec(self, erec: Tensor, blk_xoff: Tensor) -> "TilePlan":
        """Same tiling, different edge records (first-layer element table)."""
        q = TilePlan(self.kind, self.blk_info, self.blk_tile, self.tile_info, erec, self.n_blocks, self.n_tiles,
                     self.has_inactive, blk_xoff)
        q.tile_win, q.tile_geom = self.tile_win, self.tile_geom          # shared: they only depend on the geometry
        q.window = self.window
        q.blk_order = self.blk_order
        q._shared_with = self
        return q

    def update_windows(self, geom: Tensor, inv_rc: float, num_rbf: int, live: Optional[Tensor] = None) -> None:
        """Windows + record-ordered geometry for THIS geometry tensor (``live``: uint8 [E], 0 = dead entry of a Verlet-skin
        superset list).  Cached per tensor: the plan keeps a reference to the tensors it was computed for, so their addresses
        cannot be recycled by the allocator while the cache entry is alive."""
        owner = getattr(self, "_shared_with", self)
        key = (geom.data_ptr(), geom._version, tuple(geom.shape), None if live is None else (live.data_ptr(), live._version))
        if owner.win_key != key:
            ops.tc_tile_windows(owner, geom, inv_rc, num_rbf, live)
            owner.win_key = key
            # keep the STORAGE alive (no address recycling while the cache entry lives) through a detached alias: the tensor
            # itself carries its autograd node, whose context references the graph that owns this plan -- a cycle through
            # C++ objects that no garbage collector would ever break
            owner._win_pin = (geom.detach(), live)


def _window(n_live: int, n_groups: int, num_rbf: int) -> int:
    """Basis-index width of a tile: 32 (one k-chunk) when a group has enough edges to fill 64-edge tiles inside 32-wide
    windows (HVNet blocks: ~430 edges over ~13 windows), wider for sparse groups (HPNet / HTNet rows, source blocks) -- a
    half-empty tile costs the epilogue, the producers and the barriers as much as a full one, a second k-chunk only costs MMAs."""
    import os
    forced = os.environ.get("HERMNET_B200_TC_WINDOW")
    if forced:
        return int(forced)
    per_group = n_live / max(n_groups, 1)
    k32 = (num_rbf + 31) // 32 * 32
    w = 32
    while w < k32 and per_group * w / max(num_rbf, 1) < 48:      # expected edges per window below ~3/4 of a tile
        w *= 2
    return min(w, k32)


def _interleave() -> bool:
    import os
    return os.environ.get("HERMNET_B200_TC_INTERLEAVE", "1") != "0"       # (A/B switch)


def _tiles(order, kc, grp_ptr, n_groups, num_rbf, rec, grp_mod):
    """Greedy tiles of every group; returns (grp_tile [n_groups+1], tile_info, erec, n_tiles, window, n_live)."""
    dev = order.device
    n_live_host = int(grp_ptr[n_groups])
    window = _window(n_live_host, n_groups, num_rbf)
    counts = ops.tc_plan_count(order, kc, grp_ptr, n_groups, num_rbf, window)
    grp_tile = torch.zeros(n_groups + 1, dtype=torch.int32, device=dev)
    torch.cumsum(counts, 0, dtype=torch.int32, out=grp_tile[1:])
    n_tiles, n_live = torch.stack([grp_tile[-1], grp_ptr[n_groups]]).tolist()       # one host sync per plan
    tile_start = ops.tc_plan_fill(order, kc, grp_ptr, n_groups, num_rbf, grp_tile, n_tiles, window)
    tile_mod = torch.repeat_interleave(grp_mod.to(torch.int32), counts.long(), output_size=n_tiles).contiguous() if n_tiles else \
        torch.zeros(1, dtype=torch.int32, device=dev)
    erec, tile_info = ops.tc_plan_finalize(order, tile_start, n_tiles, n_live, rec, tile_mod)
    return grp_tile, tile_info, erec, n_tiles, window, n_live


def _sorted_by_group(kc: Tensor, grp: Tensor, n_groups: int, num_rbf: int):
    """Edge ids sorted by (group, basis index) with two stable counting sorts; ``grp_ptr [n_groups+2]`` (the last
    group collects the edges of inactive rows)."""
    o1 = ops.sort_by_key(kc, num_rbf)[1].long()
    grp_ptr, o2 = ops.sort_by_key(grp[o1].contiguous(), n_groups + 1)
    return o1[o2.long()].to(torch.int32).contiguous(), grp_ptr


def build_dst_plan(g, geom: Tensor, inv_rc: float, num_rbf: int) -> TilePlan:
    """Blocks = up to 32 rows of one sub-network: rows ``(atom, slot)`` of 32 consecutive atoms of one
    (element, owned/ghost) segment and one slot."""
    dev = geom.device
    R = ops.tc_block_rows(False)
    n, rpa, E = g.n_atoms, g.rows_per_atom, g.n_edges
    if n >= (1 << 27):
        raise RuntimeError("hermnet_b200: the tensor-core tile plan supports fewer than 2^27 atoms per rank")
    # segments of the internal atom order inside which row_mod only depends on the slot
    bounds = [0]
    for t in range(len(g.type_ptr) - 1):
        lo, hi = g.type_ptr[t], g.type_ptr[t + 1]
        own = min(max(g.own_count[t] if t < len(g.own_count) else hi - lo, 0), hi - lo)
        for b in (lo + own, hi):
            if b > bounds[-1]:
                bounds.append(b)
    if bounds[-1] < n:
        bounds.append(n)
    seg_lo = torch.tensor(bounds[:-1], dtype=torch.long)
    seg_hi = torch.tensor(bounds[1:], dtype=torch.long)
    n_chunks_seg = (seg_hi - seg_lo + R - 1) // R
    chunk_base = torch.zeros(len(bounds), dtype=torch.long)
    chunk_base[1:] = torch.cumsum(n_chunks_seg, 0)
    n_chunks = int(chunk_base[-1])
    # chunk tables (host, tiny): first atom and size of every chunk
    seg_of_chunk = torch.repeat_interleave(torch.arange(len(seg_lo)), n_chunks_seg)
    idx_in_seg = torch.arange(n_chunks) - chunk_base[:-1][seg_of_chunk]
    chunk_atom0 = (seg_lo[seg_of_chunk] + idx_in_seg * R)
    chunk_size = torch.minimum(seg_hi[seg_of_chunk] - chunk_atom0, torch.tensor(R))
    chunk_atom0, chunk_size = chunk_atom0.to(dev), chunk_size.to(dev)
    n_blocks = n_chunks * rpa
    slot = torch.arange(rpa, device=dev).view(1, rpa)
    row0 = (chunk_atom0.view(-1, 1) * rpa + slot)                                   # [n_chunks, rpa]
    blk_mod = g.row_mod.long()[row0.reshape(-1)] if n_blocks else torch.zeros(0, dtype=torch.long, device=dev)
    blk_info = torch.stack([row0.reshape(-1), torch.full((n_blocks,), rpa, device=dev, dtype=torch.long),
                            chunk_size.view(-1, 1).expand(-1, rpa).reshape(-1), blk_mod], 1).to(torch.int32).contiguous()
    # per atom: chunk and position inside it
    atom_chunk = torch.repeat_interleave(torch.arange(n_chunks, device=dev), chunk_size.long(), output_size=n)
    atom_local = torch.arange(n, device=dev) - chunk_atom0[atom_chunk]
    rec, kc, sub = ops.tc_plan_records(g, geom, inv_rc, num_rbf, False, atom_local.to(torch.int32).contiguous(), 1)
    if rpa == 1 and n_blocks > 0 and num_rbf <= 8192:
        # one row per atom: a block's edges are a contiguous range of the CSR -> sort every block in place (one warp per
        # block, shared memory) instead of two global counting-sort passes over all E edges
        bounds_rows = torch.cat([chunk_atom0, torch.tensor([n], device=dev)]).long()
        in_ptr = g.rowptr[bounds_rows].contiguous()
        order, grp_ptr = ops.tc_plan_sort(in_ptr, None, kc, sub, n_blocks, 1, num_rbf)
    else:
        row = g.edge_row.long()
        atom = torch.div(row, rpa, rounding_mode="floor")
        blk = atom_chunk[atom] * rpa + (row - atom * rpa)
        grp = torch.where(sub >= 0, blk, torch.full_like(blk, n_blocks)).to(torch.int32)
        order, grp_ptr = _sorted_by_group(kc, grp, n_blocks, num_rbf)
    blk_tile, tile_info, erec, n_tiles, window, n_live = _tiles(order, kc, grp_ptr, n_blocks, num_rbf, rec, blk_mod)
    blk_xoff = g.row_xoff[row0.reshape(-1)].to(torch.int32).contiguous() if n_blocks else None
    plan = TilePlan("dst", blk_info, blk_tile, tile_info, erec, n_blocks, n_tiles, n_live < E, blk_xoff)
    plan.window = window
    # processing order: every (element, owned/ghost) segment is Morton-ordered, so the chunk at fraction x of one segment is
    # spatially close to the chunks at fraction x of the others -- visit them together (stable: slots of a chunk stay adjacent)
    if _interleave() and n_blocks > 1:
        seg_len = (seg_hi - seg_lo).clamp(min=1).double()
        frac = (idx_in_seg.double() * R + 0.5 * R) / seg_len[seg_of_chunk]
        key = frac.view(-1, 1).expand(-1, rpa).reshape(-1)
        plan.blk_order = torch.sort(key.to(dev), stable=True).indices.to(torch.int32).contiguous()     # (device sort: ~0.1 ms)
    return plan


def dst_plan_for_table(plan: TilePlan, g, g0) -> TilePlan:
    """The first-layer variant of a dst plan: xh rows come from the element table of ``g0`` (hermnet._layer0_tables)."""
    eid = plan.erec[:, 3].long()
    erec0 = plan.erec.clone()
    erec0[:, 1] = g0.col[eid]                     # the "source" the kernels index xh with is the source's ELEMENT row
    erec0[:, 0] = (g0.row_xoff[g.edge_row.long()[eid]] + g0.col.long()[eid]).to(torch.int32)
    blk_xoff0 = g0.row_xoff[plan.blk_info[:, 0].long()].to(torch.int32).contiguous()
    return plan.with_erec(erec0.contiguous(), blk_xoff0)


def build_src_plan(g, geom: Tensor, inv_rc: float, num_rbf: int) -> TilePlan:
    """Blocks = 32 consecutive source atoms; groups = (block, sub-network)."""
    dev = geom.device
    R = ops.tc_block_rows(True)
    n, E, M = g.n_atoms, g.n_edges, g.n_modules
    n_blocks = (n + R - 1) // R
    b = torch.arange(n_blocks, device=dev, dtype=torch.long)
    blk_info = torch.stack([b * R, torch.ones_like(b), torch.clamp(n - b * R, max=R), torch.full_like(b, -1)], 1).to(torch.int32).contiguous()
    n_groups = n_blocks * M
    rec, kc, sub = ops.tc_plan_records(g, geom, inv_rc, num_rbf, True, None, R)
    if n_blocks > 0 and M * num_rbf <= 8192:
        # a source block's edges are a contiguous range of the transposed CSR: segment-local sort by (sub-network, basis index)
        sb = torch.arange(n_blocks + 1, device=dev, dtype=torch.long) * R
        in_ptr = g.t_rowptr[sb.clamp(max=n)].contiguous()
        order, grp_ptr = ops.tc_plan_sort(in_ptr, g.t_eid, kc, sub, n_blocks, M, num_rbf)
    else:
        sblk = torch.div(g.col.long(), R, rounding_mode="floor")
        grp = torch.where(sub >= 0, sblk * M + sub.long(), torch.full_like(sblk, n_groups)).to(torch.int32)
        order, grp_ptr = _sorted_by_group(kc, grp, n_groups, num_rbf)
    grp_mod = torch.arange(M, device=dev).repeat(n_blocks)
    grp_tile, tile_info, erec, n_tiles, window, n_live = _tiles(order, kc, grp_ptr, n_groups, num_rbf, rec, grp_mod)
    blk_tile = grp_tile[::M].contiguous()
    plan = TilePlan("src", blk_info, blk_tile, tile_info, erec, n_blocks, n_tiles, n_live < E)
    plan.window = window
    if _interleave() and n_blocks > 1 and len(g.type_ptr) > 2:       # same idea for the source blocks (consecutive internal atoms)
        tp = torch.tensor(g.type_ptr, dtype=torch.double)
        first = (torch.arange(n_blocks, dtype=torch.double) * R + 0.5 * R).clamp(max=max(n - 1, 0))
        seg = torch.clamp(torch.searchsorted(tp, first, right=True) - 1, 0, len(g.type_ptr) - 2)
        lo_, hi_ = tp[seg], tp[seg + 1]
        key = (first - lo_) / (hi_ - lo_).clamp(min=1.0)
        plan.blk_order = torch.sort(key.to(dev), stable=True).indices.to(torch.int32).contiguous()
    return plan


def plans_of(g, geom: Tensor, inv_rc: float, num_rbf: int, want_src: bool):
    """Plans cached on the graph (built from the geometry of the first call; later geometries only refresh the windows)."""
    dst = g._lazy.get("tc_dst")
    if dst is None:
        parent = getattr(g, "_tc_parent", None)
        parent = parent() if parent is not None else None
        if parent is not None:      # first-layer element-table view of ``parent`` (same rows, edges and geometry)
            base, _ = plans_of(parent, geom, inv_rc, num_rbf, False)
            dst = dst_plan_for_table(base, parent, g)
        else:
            dst = build_dst_plan(g, geom, inv_rc, num_rbf)
        g._lazy["tc_dst"] = dst
    src: Optional[TilePlan] = g._lazy.get("tc_src")
    if want_src and src is None:
        src = g._lazy["tc_src"] = build_src_plan(g, geom, inv_rc, num_rbf)
    return dst, src
