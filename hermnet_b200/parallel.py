"""Multi-GPU execution of the hot path: one process per GPU, ``torch.distributed`` (NCCL over NVLink/NVSwitch).

The reference only has data-parallel training (example/dist_train.py:17-143: DDP gradient all-reduce); ``north_star``
adds spatial domain decomposition for a single large system (BASELINE.json configs[3]).  Both live here.

Domain decomposition (``DomainDecomposition``)
  * the periodic cell is cut into ``px x py x pz`` fractional bricks, one per rank; positions are replicated (12 MB
    for 1M atoms), so every rank selects its own atoms plus the halo candidates (within ``rc`` of the brick) locally;
  * the rank builds a row CSR whose destinations are its OWNED atoms and whose sources are owned + ghost atoms
    (``GraphBuilder.from_local_positions``); edge shifts keep the full periodicity, a ghost is one row per atom, not
    per image;
  * message passing has range ``rc`` per layer, so ghost FEATURES are refreshed before every layer but the first
    (layer 0 reads embeddings, which are local): ``Halo.exchange`` = pack kernel (``hn_gather_rows``) -> all-to-all
    -> ghost rows; its backward is the transposed exchange with a deterministic segmented accumulation
    (``hn_segment_sum``) onto the owners -- the reverse force accumulation;
  * energies are all-reduced scalars; position gradients of owned and ghost atoms are summed with one all-reduce of
    the ``[N,3]`` array.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import ops
from .graph import RowGraph, Segments

Tensor = torch.Tensor


def _grid(world: int) -> Tuple[int, int, int]:
    """Most cubic factorisation px >= py >= pz of ``world``."""
    best = (world, 1, 1)
    for a in range(1, world + 1):
        if world % a:
            continue
        for b in range(1, world // a + 1):
            if (world // a) % b:
                continue
            c = world // a // b
            cand = tuple(sorted((a, b, c), reverse=True))
            if max(cand) - min(cand) < max(best) - min(best):
                best = cand
    return best


def _all_to_all(send: Tensor, send_counts: List[int], recv_counts: List[int], group=None) -> Tensor:
    """Variable-size all-to-all of rows.  NCCL: ``all_to_all_single``; other backends (gloo in the CPU tests):
    batched point-to-point."""
    recv = torch.empty((sum(recv_counts),) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
    if dist.get_backend(group) == "nccl":
        dist.all_to_all_single(recv, send.contiguous(), recv_counts, send_counts, group=group)
        return recv
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    so = [0]
    ro = [0]
    for c in send_counts:
        so.append(so[-1] + c)
    for c in recv_counts:
        ro.append(ro[-1] + c)
    recv[ro[rank]:ro[rank + 1]] = send[so[rank]:so[rank + 1]]
    reqs = []
    for peer in range(world):
        if peer == rank:
            continue
        if recv_counts[peer]:
            reqs.append(dist.P2POp(dist.irecv, recv[ro[peer]:ro[peer + 1]], peer, group))
        if send_counts[peer]:
            reqs.append(dist.P2POp(dist.isend, send[so[peer]:so[peer + 1]].contiguous(), peer, group))
    if reqs:
        for r in dist.batch_isend_irecv(reqs):
            r.wait()
    return recv


class PeerBuffers:
    """Symmetric landing buffers of the halo exchange, mapped into every rank of the group (NVLink peer memory through
    ``torch.distributed._symmetric_memory``): ``fwd [2, cap_ghost, W]`` receives ghost rows, ``bwd [2, cap_send, W]`` the
    gradients flowing back to the owners; two slots each, used alternately (see ``_HaloExchange``).  Allocated once per
    ``DomainDecomposition`` and re-used while the capacities suffice (every rank takes the same decision)."""

    def __init__(self, cap_ghost: int, cap_send: int, width: int, device, group=None):
        import torch.distributed._symmetric_memory as symm
        self.cap_ghost, self.cap_send, self.width = cap_ghost, cap_send, width
        grp = group if group is not None else dist.group.WORLD
        self.fwd = symm.empty((2, max(cap_ghost, 1), width), dtype=torch.float32, device=device)
        self.bwd = symm.empty((2, max(cap_send, 1), width), dtype=torch.float32, device=device)
        self.h_fwd = symm.rendezvous(self.fwd, grp)
        self.h_bwd = symm.rendezvous(self.bwd, grp)
        world = dist.get_world_size(group)
        row_bytes = width * 4
        # device addresses of every peer's slots: [2, world]
        self.fwd_ptrs = torch.tensor([[int(p) + s * max(cap_ghost, 1) * row_bytes for p in self.h_fwd.buffer_ptrs[:world]]
                                      for s in range(2)], dtype=torch.int64, device=device)
        self.bwd_ptrs = torch.tensor([[int(p) + s * max(cap_send, 1) * row_bytes for p in self.h_bwd.buffer_ptrs[:world]]
                                      for s in range(2)], dtype=torch.int64, device=device)
        self.n_fwd = self.n_bwd = 0       # exchanges issued so far (slot = count & 1)


class Halo:
    """Send / receive lists of one rank.  ``send_idx``: local rows to pack, grouped by destination rank;
    ``ghost_idx``: local ghost rows in the order the peers send them (grouped by owner rank).

    Transports of ``exchange`` (chosen once per halo):
      * ``peer``  -- ``hn_halo_pack`` stores every row straight into the destination rank's ghost landing buffer over NVLink
        peer memory, one device-side barrier, ``hn_halo_unpack`` moves the landed rows into the ghost rows; the backward
        pass packs the ghost-row gradients into the OWNERS' buffers and reduces them there (``hn_segment_sum``);
      * ``nccl``  -- the same kernels around one ``all_to_all_single`` (when symmetric memory cannot be set up);
      * ``generic`` -- gather / point-to-point / index_copy on any backend (CPU tests with gloo)."""

    def __init__(self, send_idx: Tensor, send_counts: List[int], ghost_idx: Tensor, recv_counts: List[int],
                 n_local: int, group=None, peer: Optional[PeerBuffers] = None):
        self.send_idx = send_idx.to(torch.int32).contiguous()
        self.ghost_idx = ghost_idx.long().contiguous()
        self.ghost_idx32 = ghost_idx.to(torch.int32).contiguous()
        self.send_counts, self.recv_counts, self.group, self.n_local = send_counts, recv_counts, group, n_local
        self.seg_send = Segments.from_index(self.send_idx, n_local)     # adjoint of the pack gather
        self.bytes_per_exchange = 0
        dev = self.send_idx.device
        self.peer = peer
        self.transport = "generic"
        if dev.type == "cuda" and dist.is_initialized() and dist.get_backend(group) == "nccl":
            self.transport = "peer" if peer is not None else "nccl"
            rank, world = dist.get_rank(group), dist.get_world_size(group)
            n_send, n_ghost = int(self.send_idx.numel()), int(self.ghost_idx.numel())
            i32 = dict(dtype=torch.int32, device=dev)
            if self.transport == "nccl":      # one local send buffer: "peer" 0, slot = position in the list
                self.fwd_peer = torch.zeros(n_send, **i32)
                self.fwd_slot = torch.arange(n_send, **i32)
                self.bwd_peer = torch.zeros(n_ghost, **i32)
                self.bwd_slot = torch.arange(n_ghost, **i32)
            else:
                # every rank's counts: S[r][p] rows r sends to p, R[r][p] rows r receives from p (= S[p][r]); device ops only
                mine = torch.tensor([send_counts, recv_counts], dtype=torch.int64, device=dev)
                allc = torch.empty((world, 2, world), dtype=torch.int64, device=dev)
                dist.all_gather_into_tensor(allc, mine, group=group)
                S, R = allc[:, 0], allc[:, 1]
                recv_off = (torch.cumsum(R, 1) - R)[:, rank]      # [p]: where MY rows start in rank p's ghost zone
                send_off = (torch.cumsum(S, 1) - S)[:, rank]      # [o]: where the rows for ME start in owner o's send list
                ranks = torch.arange(world, device=dev)

                def lists(counts, off):      # rows grouped by rank: (rank of every row, its slot in that rank's buffer)
                    c = torch.tensor(counts, dtype=torch.int64, device=dev)
                    start = torch.cumsum(c, 0) - c
                    who = torch.repeat_interleave(ranks, c)
                    slot = torch.arange(int(sum(counts)), device=dev) - start[who] + off[who]
                    return who.to(torch.int32).contiguous(), slot.to(torch.int32).contiguous()
                self.fwd_peer, self.fwd_slot = lists(send_counts, recv_off)      # my send list is grouped by destination rank
                self.bwd_peer, self.bwd_slot = lists(recv_counts, send_off)      # my ghost list: by owner, in the owner's send order

    def exchange(self, x: Tensor, vec: Tensor) -> Tuple[Tensor, Tensor]:
        F = x.size(1)
        if self.transport == "generic":
            feat = _HaloExchangeGeneric.apply(torch.cat([x, vec.reshape(-1, 3 * F)], 1), self)
            return feat[:, :F], feat[:, F:].reshape(-1, 3, F)
        return _HaloExchange.apply(x, vec, self)


class _HaloExchangeGeneric(Function):
    @staticmethod
    def forward(ctx, feat: Tensor, halo: Halo):
        ctx.halo = halo
        send = ops.gather_rows(feat.contiguous(), halo.send_idx)
        recv = _all_to_all(send, halo.send_counts, halo.recv_counts, halo.group)
        halo.bytes_per_exchange = send.numel() * 4
        return feat.index_copy(0, halo.ghost_idx, recv)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        halo = ctx.halo
        g = g.contiguous()
        back = _all_to_all(g[halo.ghost_idx], halo.recv_counts, halo.send_counts, halo.group)
        sg = halo.seg_send
        own = ops.segment_sum(back.contiguous(), sg.rowptr, sg.perm, sg.n_rows)     # reverse accumulation on owners
        return g.index_fill(0, halo.ghost_idx, 0.0) + own, None


class _HaloExchange(Function):
    """Ghost rows of ``x`` / ``vec`` refreshed IN PLACE from their owners (the ghost rows of the inputs are padding written
    by the node update).  Slot protocol of the peer transport: exchange k writes the peers' slot k & 1, then every rank
    passes one device-side barrier, then reads its own slot; a rank can only reach exchange k + 2 (same slot) after barrier
    k + 1, which every peer enters after its unpack k in stream order -- no second barrier is needed."""

    @staticmethod
    def forward(ctx, x: Tensor, vec: Tensor, halo: Halo):
        ctx.halo = halo
        ctx.mark_dirty(x, vec)
        F = x.size(1)
        n_send, n_ghost = int(halo.send_idx.numel()), int(halo.ghost_idx.numel())
        halo.bytes_per_exchange = n_send * 4 * F * 4
        v2 = vec.view(-1, 3 * F)
        if halo.transport == "peer":
            pb = halo.peer
            slot = pb.n_fwd & 1
            pb.n_fwd += 1
            ops.halo_pack(x, v2, halo.send_idx, halo.fwd_peer, halo.fwd_slot, pb.fwd_ptrs[slot])
            pb.h_fwd.barrier(channel=0)
            ops.halo_unpack(pb.fwd[slot], halo.ghost_idx32, x, v2)
        else:
            send = torch.empty((n_send, 4 * F), dtype=x.dtype, device=x.device)
            base = torch.tensor([send.data_ptr()], dtype=torch.int64, device=x.device)
            ops.halo_pack(x, v2, halo.send_idx, halo.fwd_peer, halo.fwd_slot, base)
            recv = torch.empty((n_ghost, 4 * F), dtype=x.dtype, device=x.device)
            dist.all_to_all_single(recv, send, halo.recv_counts, halo.send_counts, group=halo.group)
            ops.halo_unpack(recv, halo.ghost_idx32, x, v2)
        return x, vec

    @staticmethod
    @once_differentiable
    def backward(ctx, g_x, g_vec):
        halo = ctx.halo
        g_x, g_vec = g_x.contiguous(), g_vec.contiguous()
        F = g_x.size(1)
        gv2 = g_vec.view(-1, 3 * F)
        n_send, n_ghost = int(halo.send_idx.numel()), int(halo.ghost_idx.numel())
        sg = halo.seg_send
        if halo.transport == "peer":
            pb = halo.peer
            slot = pb.n_bwd & 1
            pb.n_bwd += 1
            ops.halo_pack(g_x, gv2, halo.ghost_idx32, halo.bwd_peer, halo.bwd_slot, pb.bwd_ptrs[slot])
            pb.h_bwd.barrier(channel=0)
            back = pb.bwd[slot][:n_send]
        else:
            send = torch.empty((n_ghost, 4 * F), dtype=g_x.dtype, device=g_x.device)
            base = torch.tensor([send.data_ptr()], dtype=torch.int64, device=g_x.device)
            ops.halo_pack(g_x, gv2, halo.ghost_idx32, halo.bwd_peer, halo.bwd_slot, base)
            back = torch.empty((n_send, 4 * F), dtype=g_x.dtype, device=g_x.device)
            dist.all_to_all_single(back, send, halo.send_counts, halo.recv_counts, group=halo.group)
        own = ops.segment_sum(back.contiguous(), sg.rowptr, sg.perm, sg.n_rows)     # reverse accumulation on the owners
        gx = g_x.index_fill(0, halo.ghost_idx, 0.0).add_(own[:, :F])
        gv = g_vec.index_fill(0, halo.ghost_idx, 0.0)
        gv.view(-1, 3 * F).add_(own[:, F:])
        return gx, gv, None


class DomainDecomposition:
    """Energy + forces of ONE periodic system spread over the ranks of ``group`` (inference path)."""

    def __init__(self, model, device, group=None):
        self.model, self.device, self.group = model, device, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.grid = _grid(self.world)
        self.graph: Optional[RowGraph] = None
        self.halo: Optional[Halo] = None
        self.peer: Optional[PeerBuffers] = None
        # HERMNET_B200_HALO=nccl forces the all-to-all transport (A/B measurements); default: peer memory when it can be set up
        import os
        self.want_peer = os.environ.get("HERMNET_B200_HALO", "peer") == "peer"
        self.peer_error: Optional[str] = None

    # ------------------------------------------------------------------------------------------------------
    def _assign(self, pos: Tensor, cell: Tensor):
        """Brick (= owner rank) of every atom and the mask of this rank's local candidates (owned + halo)."""
        c = cell.reshape(3, 3).double()
        frac = torch.linalg.solve(c.T, pos.double().T).T
        frac = frac - torch.floor(frac)
        vol = torch.det(c).abs()
        heights = torch.stack([vol / torch.linalg.norm(torch.cross(c[(a + 1) % 3], c[(a + 2) % 3], dim=0)) for a in range(3)])
        margin = (self.model.rc * (1 + 1e-6) / heights).tolist()
        owner = torch.zeros(pos.size(0), dtype=torch.long, device=pos.device)
        local = torch.ones(pos.size(0), dtype=torch.bool, device=pos.device)
        rk = self.rank
        coords = []
        for a in range(3):
            coords.append(rk % self.grid[a])
            rk //= self.grid[a]
        stride = 1
        for a in range(3):
            p = self.grid[a]
            b = torch.clamp((frac[:, a] * p).long(), max=p - 1)
            owner += b * stride
            stride *= p
            if p > 1:
                lo, hi = coords[a] / p, (coords[a] + 1) / p
                f = frac[:, a]
                m = margin[a]
                inside = torch.zeros_like(local)
                for k in (-1.0, 0.0, 1.0):            # periodic images of the atom along this axis
                    inside |= (f + k >= lo - m) & (f + k < hi + m)
                local &= inside if 2 * m + 1.0 / p < 1.0 else torch.ones_like(local)
        return owner, local

    def build(self, pos: Tensor, Z: Tensor, cell: Tensor):
        """Partition + local graph + halo lists for the replicated ``pos [N,3]``, ``Z [N]``, ``cell [1,3,3]``."""
        dev = pos.device
        N = pos.size(0)
        owner, local = self._assign(pos.detach(), cell.detach())
        ids = torch.nonzero(local).squeeze(1)                          # global ids of local atoms, ascending
        owned = owner[ids] == self.rank
        g = self.model.builder.from_local_positions(pos.detach()[ids], Z[ids], cell, owned)
        act = g.mod_active.clone()            # "sub-network has no edges" (hermnet.py:56-57) is a GLOBAL property
        dist.all_reduce(act, op=dist.ReduceOp.MAX, group=self.group)
        g.mod_active, g.mod_active_host = act, (act > 0).tolist()
        gid = ids[g.perm]                                              # global id of each internal local atom
        n_loc = gid.numel()
        n_own = sum(g.own_count)
        owned_i = owned[g.perm]
        ghost_int = torch.nonzero(~owned_i).squeeze(1)
        g_owner = owner[gid[ghost_int]]
        order = torch.argsort(g_owner * N + gid[ghost_int])            # by owner rank, then global id
        ghost_idx = ghost_int[order]
        want = gid[ghost_idx]                                          # global ids requested, grouped by owner
        recv_counts = torch.bincount(g_owner, minlength=self.world).tolist()
        # tell every owner which of its atoms we need
        cnt_out = torch.tensor(recv_counts, dtype=torch.long, device=dev)
        cnt_in = _all_to_all(cnt_out, [1] * self.world, [1] * self.world, self.group)
        send_counts = cnt_in.tolist()
        asked = _all_to_all(want, recv_counts, send_counts, self.group)
        g2l = torch.full((N,), -1, dtype=torch.long, device=dev)
        g2l[gid] = torch.arange(n_loc, device=dev)
        send_idx = g2l[asked]
        self.graph, self.ids = g, ids
        self.Z_local = Z[ids]                                          # local (pre-permutation) order, like pos[ids]
        self.cell = cell
        self.n_atoms, self.n_owned = N, n_own
        stats = torch.tensor([n_own, g.n_edges, n_loc - n_own, int(send_idx.numel())], dtype=torch.long, device=dev)
        both = torch.stack([stats, -stats])                            # one all-reduce: maxima, and minima as -max(-x)
        sm = stats.clone()
        dist.all_reduce(both, op=dist.ReduceOp.MAX, group=self.group)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM, group=self.group)
        self.n_owned_max, self.local_edges_max, self.n_ghost_max, n_send_max = (int(v) for v in both[0].tolist())
        self.global_edges = int(sm[1])
        self.halo = Halo(send_idx, send_counts, ghost_idx, recv_counts, n_loc, self.group,
                         self._peer_buffers(self.n_ghost_max, n_send_max, dev))
        assert int(sm[0]) == N, "every atom must be owned by exactly one rank"
        return self

    def _peer_buffers(self, ng: int, ns: int, dev) -> Optional[PeerBuffers]:
        """Symmetric landing buffers (allocated once, grown when a new decomposition needs more rows); ``None`` when peer
        memory is not available.  ``ng`` / ``ns`` are the all-reduced maxima of ghost / send rows, so every rank takes the same
        decision without another collective; a failed set-up is shared with one."""
        if not (self.want_peer and dev.type == "cuda" and dist.get_backend(self.group) == "nccl" and self.world > 1):
            return None
        width = 4 * self.model.hidden_channels
        if self.peer is not None and self.peer.cap_ghost >= ng and self.peer.cap_send >= ns and self.peer.width == width:
            return self.peer
        ok = torch.ones(1, dtype=torch.long, device=dev)
        peer = None
        try:
            peer = PeerBuffers(int(ng * 1.25) + 64, int(ns * 1.25) + 64, width, dev, self.group)
        except Exception as exc:  # noqa: BLE001 -- no symmetric memory on this system: NCCL transport
            self.peer_error = f"{type(exc).__name__}: {exc}"
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if int(ok) == 0:
            self.want_peer, peer = False, None
        self.peer = peer
        return peer

    def energy_forces(self, pos: Tensor):
        """Total energy ``[1]`` and ``dE/dpos [N,3]`` (forces = minus that), identical on every rank."""
        p = pos.detach()[self.ids].requires_grad_(True)
        e, _, _ = self.model.forward_graph(p, self.Z_local, self.cell, self.graph, halo=self.halo)
        if getattr(self.model, "intensive", False):       # mean over ALL atoms of the system (hermnet.py:130), not per rank
            e = e / float(max(self.n_atoms, 1))
        (gl,) = torch.autograd.grad(e.sum(), p)
        grad = torch.zeros((self.n_atoms, 3), dtype=gl.dtype, device=gl.device)
        grad.index_add_(0, self.ids, gl)                               # ids are unique per rank: no collisions
        e = e.detach().clone()
        dist.all_reduce(e, group=self.group)
        dist.all_reduce(grad, group=self.group)
        return e, grad


# ----------------------------------------------------------------------------------------------------------
# data parallelism (example/dist_train.py:57-67,99)
# ----------------------------------------------------------------------------------------------------------
def data_parallel(model, device_ids=None, **kw):
    """DDP wrapper with ``find_unused_parameters=True``: a rank whose batch lacks an element leaves that
    sub-network's parameters untouched (hermnet.py:56-57), which plain DDP would wait on forever."""
    kw.setdefault("find_unused_parameters", True)
    return torch.nn.parallel.DistributedDataParallel(model, device_ids=device_ids, **kw)


def force_matching_step(model, data, optimizer, gamma: float = 0.8, trn_mean: float = 0.0):
    """One training step of example/dist_train.py:84-104: MSE on energies and on forces obtained with
    ``create_graph=True`` (double backward through the composite formulation), gradient all-reduce by DDP."""
    optimizer.zero_grad()
    data.pos.requires_grad_(True)
    pred_e = model(data)
    e_loss = torch.nn.functional.mse_loss(pred_e, data.y - trn_mean)
    pred_f = -torch.autograd.grad(pred_e.sum(), data.pos, create_graph=True)[0]
    f_loss = torch.nn.functional.mse_loss(pred_f, data.forces)
    loss = (1 - gamma) * e_loss + gamma * f_loss
    loss.backward()
    optimizer.step()
    return loss.detach(), pred_e.detach(), pred_f.detach()


def force_matching_step_microbatched(model, graphs, optimizer, micro: int = 4, gamma: float = 0.8, trn_mean: float = 0.0,
                                     device=None):
    """The same training step (example/dist_train.py:84-104) for a local batch given as a LIST of graphs, run in
    micro-batches of ``micro`` graphs with gradient accumulation: the composite (double-backward) formulation keeps
    ``[E,3F]`` tensors alive, a 32 x 4096-atom batch does not fit otherwise.  Under DDP only the last micro-batch
    all-reduces (``no_sync`` on the others).  ``graphs``: ``Data`` objects or already collated micro-batches
    (``Batch``; re-using them across steps re-uses their cached device graphs).  Losses are weighted so that the sum equals
    the MSE over the whole local batch (energies per graph, forces per component)."""
    import contextlib
    from .data import Batch
    optimizer.zero_grad()
    batches = list(graphs) if isinstance(graphs[0], Batch) else \
        [Batch.from_data_list(graphs[i:i + micro]) for i in range(0, len(graphs), micro)]
    n_graphs = sum(int(b.get("num_graphs") or 1) for b in batches)
    n_atoms = sum(int(b.pos.size(0)) for b in batches)
    total = None
    for i, data in enumerate(batches):
        if device is not None and data.pos.device != torch.device(device):
            data = data.to(device)
        sync = i == len(batches) - 1 or not hasattr(model, "no_sync")
        with (contextlib.nullcontext() if sync else model.no_sync()):
            data.pos.requires_grad_(True)
            pred_e = model(data)
            w_e = float(pred_e.numel()) / n_graphs
            w_f = float(data.pos.size(0)) / n_atoms
            e_loss = torch.nn.functional.mse_loss(pred_e, data.y.reshape(-1) - trn_mean)
            pred_f = -torch.autograd.grad(pred_e.sum(), data.pos, create_graph=True)[0]
            f_loss = torch.nn.functional.mse_loss(pred_f, data.forces)
            loss = (1 - gamma) * w_e * e_loss + gamma * w_f * f_loss
            loss.backward()
        total = loss.detach() if total is None else total + loss.detach()
        data.pos.requires_grad_(False)
        data.pos.grad = None
    optimizer.step()
    return total
