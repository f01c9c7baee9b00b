"""Autograd layer over the CUDA kernels.

Two formulations of the same arithmetic (HermNet/hermnet.py:133-152, rmnet.py:51-73):

* **fused** (inference / first-order): ``EdgeGeometry`` and ``PaiNNEdge`` call the fused kernels and their
  hand-written backward kernels (forces, virial, feature and filter-weight gradients).  Marked
  ``once_differentiable``.
* **composite** (training with a force loss needs ``create_graph=True`` -> double backward):
  the same maths expressed with ``gather_rows`` / ``segment_sum`` -- two mutually adjoint linear kernels, so the
  formulation is differentiable to any order -- plus ordinary torch elementwise ops and GEMMs.

Neither uses torch_scatter / PyG / atomics; there is no CPU implementation.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import ops
from .graph import RowGraph, Segments

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------------
# adjoint pair: gather rows  <->  segmented sum
# ----------------------------------------------------------------------------------------------------
class _GatherRows(Function):
    @staticmethod
    def forward(ctx, X: Tensor, seg: Segments):
        ctx.seg = seg
        ctx.shape = X.shape
        return ops.gather_rows(X.reshape(X.size(0), -1).contiguous(), seg.index).view((-1,) + tuple(X.shape[1:]))

    @staticmethod
    def backward(ctx, g):
        return segment_sum(g, ctx.seg), None


class _SegmentSum(Function):
    @staticmethod
    def forward(ctx, Y: Tensor, seg: Segments):
        ctx.seg = seg
        out = ops.segment_sum(Y.reshape(Y.size(0), -1).contiguous(), seg.rowptr, seg.perm, seg.n_rows)
        return out.view((seg.n_rows,) + tuple(Y.shape[1:]))

    @staticmethod
    def backward(ctx, g):
        return gather_rows(g, ctx.seg), None


def gather_rows(X: Tensor, seg: Segments) -> Tensor:
    """``out[i] = X[seg.index[i]]`` (rows of X = rows of the segmentation)."""
    return _GatherRows.apply(X, seg)


def segment_sum(Y: Tensor, seg: Segments) -> Tensor:
    """``out[r] = sum of Y[i] over the items i of row r`` -- deterministic, no atomics."""
    return _SegmentSum.apply(Y, seg)


# ----------------------------------------------------------------------------------------------------
# fused formulation
# ----------------------------------------------------------------------------------------------------
class _EdgeGeometry(Function):
    @staticmethod
    def forward(ctx, pos: Tensor, cell: Optional[Tensor], g: RowGraph):
        geom = ops.edge_geom_fwd(pos.contiguous(), None if cell is None else cell.contiguous(), g)
        ctx.g = g
        ctx.has_cell = cell is not None
        ctx.save_for_backward(geom)
        return geom

    @staticmethod
    @once_differentiable
    def backward(ctx, g_geom):
        (geom,) = ctx.saved_tensors
        g = ctx.g
        want_cell = ctx.has_cell and ctx.needs_input_grad[1]
        grad_pos, cellw = ops.edge_geom_bwd(geom, g_geom.contiguous().view(1, -1, 4), g, want_cell)
        grad_cell = None
        if want_cell:
            sb = g.seg_batch
            grad_cell = ops.segment_sum(cellw, sb.rowptr, sb.perm, sb.n_rows).view(-1, 3, 3)[: g.n_graphs]
        return grad_pos, grad_cell, None


def edge_geometry(pos: Tensor, cell: Optional[Tensor], g: RowGraph) -> Tensor:
    """``geom[e] = (ux, uy, uz, d)`` for every row-edge (internal atom order)."""
    return _EdgeGeometry.apply(pos, cell, g)


def _plan_geometry(g: RowGraph, geom: Tensor):
    """``geom`` re-ordered into the slot order of the two TilePlans, cached on the graph for the tensor object in hand
    (all layers of one evaluation share it)."""
    hit = g._lazy.get("plan_geom")
    if hit is not None and hit[0] is geom and hit[1] == geom._version:
        return hit[2], hit[3]
    gd = geom.detach()
    geom_b = ops.gather_rows(gd, g.plan_dst.eid)
    geom_s = ops.gather_rows(gd, g.plan_src.eid) if g.plan_src is not None else None
    g._lazy["plan_geom"] = (geom, geom._version, geom_b, geom_s)
    return geom_b, geom_s


class _PaiNNEdge(Function):
    @staticmethod
    def forward(ctx, xh, vec, geom, Wt, bias, offset, g: RowGraph, p):
        xh, vec, geom, Wt, bias = xh.contiguous(), vec.contiguous(), geom.contiguous(), Wt.contiguous(), bias.contiguous()
        ctx.tiled = g.plan_dst is not None and ops.edge_tiled_supported(p.hidden, p.num_rbf)
        if ctx.tiled:
            geom_b, _ = _plan_geometry(g, geom)
            dx, dvec = ops.painn_edge_fwd_tiled(p, xh, vec, geom_b, g.plan_dst, Wt, bias, offset)
        else:
            dx, dvec = ops.painn_edge_fwd(p, xh, vec, geom, g, Wt, bias, offset)
        ctx.g, ctx.p = g, p
        ctx.save_for_backward(xh, vec, geom, Wt, bias, offset)
        return dx, dvec

    @staticmethod
    @once_differentiable
    def backward(ctx, g_dx, g_dvec):
        xh, vec, geom, Wt, bias, offset = ctx.saved_tensors
        g, p = ctx.g, ctx.p
        g_dx, g_dvec = g_dx.contiguous(), g_dvec.contiguous()
        need = ctx.needs_input_grad
        grad_xh = grad_vec = grad_geom = grad_W = grad_b = None
        geom_b = geom_s = None
        if ctx.tiled:
            geom_b, geom_s = _plan_geometry(g, geom)
        if need[2]:
            if ctx.tiled:
                parts = ops.painn_edge_bwd_dst_tiled(p, xh, vec, geom_b, g.plan_dst, Wt, bias, offset, g_dx, g_dvec)
                grad_geom = ops.gather_rows(parts[0] if parts.size(0) == 1 else parts.sum(0), g.plan_dst.pos_of)
            else:
                parts = ops.painn_edge_bwd_dst(p, xh, vec, geom, g, Wt, bias, offset, g_dx, g_dvec)
                grad_geom = parts[0] if parts.size(0) == 1 else parts.sum(0)
        if need[0] or need[1]:
            if ctx.tiled and g.plan_src is not None:
                grad_xh, grad_vec = ops.painn_edge_bwd_src_tiled(p, xh, vec, geom_s, g.plan_src, Wt, bias, offset, g_dx, g_dvec)
            else:
                grad_xh, grad_vec = ops.painn_edge_bwd_src(p, xh, vec, geom, g, Wt, bias, offset, g_dx, g_dvec)
        if need[3] or need[4]:
            grad_W, grad_b = ops.painn_edge_bwd_w(p, xh, vec, geom, g, offset, g_dx, g_dvec)
        return grad_xh, grad_vec, grad_geom, grad_W, grad_b, None, None, None


def painn_edge(xh, vec, geom, Wt, bias, offset, g: RowGraph, p):
    """Fused gather -> filter -> message -> segmented reduction.  Returns ``dx [R,F]``, ``dvec [R,3,F]``."""
    return _PaiNNEdge.apply(xh, vec, geom, Wt, bias, offset, g, p)


# ----------------------------------------------------------------------------------------------------
# node-side dense layers on the tensor cores (fused path)
# ----------------------------------------------------------------------------------------------------
_SPLIT_CACHE = {}


def _split_cached(w: Tensor, transposed: bool):
    """(hi, lo) TF32 split of a weight (or of its transpose), cached per live tensor object and version (addresses and
    ids are recycled by the allocator, so the entry keeps a weak reference and is only trusted for the same object)."""
    import weakref
    key = (id(w), transposed)
    hit = _SPLIT_CACHE.get(key)
    if hit is not None and hit[0]() is w and hit[1] == w._version:
        return hit[2], hit[3]
    hi, lo = ops.split_tf32(w.detach().t().contiguous() if transposed else w.detach())
    if len(_SPLIT_CACHE) > 1024:
        _SPLIT_CACHE.clear()
    _SPLIT_CACHE[key] = (weakref.ref(w), w._version, hi, lo)
    return hi, lo


class _LinearTC(Function):
    """``x @ W^T + b`` through ``hn_gemm_tf32x3`` (tcgen05, 3xTF32).  Backward: the data gradient uses the same kernel
    with the transposed weight; weight / bias gradients (only when parameters require grad) are plain torch GEMMs."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        hi, lo = _split_cached(weight, False)
        x2 = x.reshape(-1, x.size(-1))
        out = ops.gemm_tf32x3(x2, hi, lo, None if bias is None else bias.detach().contiguous())
        return out.view(tuple(x.shape[:-1]) + (weight.size(0),))

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        gx = gw = gb = None
        g2 = g.reshape(-1, g.size(-1))
        if ctx.needs_input_grad[0]:
            if ops.gemm_supported(weight.size(0), weight.size(1)):
                hi, lo = _split_cached(weight, True)
                gx = ops.gemm_tf32x3(g2.contiguous(), hi, lo, None).view(x.shape)
            else:
                gx = (g2 @ weight).view(x.shape)
        if ctx.needs_input_grad[1]:
            gw = g2.t() @ x.reshape(-1, x.size(-1))
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = g2.sum(0)
        return gx, gw, gb


def linear(x: Tensor, weight: Tensor, bias: Optional[Tensor], tensor_cores: bool) -> Tensor:
    """nn.Linear forward; on the fused path (fp32, K % 32 == 0, N % 64 == 0) it runs on the tensor cores."""
    if tensor_cores and x.dtype == torch.float32 and ops.gemm_supported(weight.size(1), weight.size(0)) and x.numel() > 0:
        return _LinearTC.apply(x, weight, bias)
    return torch.nn.functional.linear(x, weight, bias)


# ----------------------------------------------------------------------------------------------------
# composite (any-order differentiable) formulation
# ----------------------------------------------------------------------------------------------------
def edge_geometry_composite(pos: Tensor, cell: Optional[Tensor], g: RowGraph) -> Tensor:
    ps = gather_rows(pos, g.seg_src)
    pr = gather_rows(pos, g.seg_dst_atom)
    D = ps - pr
    if cell is not None:
        ce = gather_rows(cell.reshape(-1, 9), g.seg_edge_graph).view(-1, 3, 3)
        S = g.shift[:, :3].to(pos.dtype) * g.sign
        D = D + (S.unsqueeze(2) * ce).sum(1)
    d = D.norm(dim=-1)
    d = torch.where(d <= 1.0e-6, torch.full_like(d, 1.0e-6), d)
    return torch.cat([D / d[:, None], d[:, None]], dim=1)


def painn_edge_composite_flat(xh, vec, geom, Wt, bias, radial_basis, g: RowGraph):
    """Same contract as ``painn_edge`` (``xh`` is the flat ``[rows, 3F]`` buffer of ``graph.xh_sources``) built from
    differentiable pieces; materialises per-edge tensors, so it is meant for training-size batches."""
    F3 = xh.size(1)
    F = F3 // 3
    M = Wt.size(0)
    P = gather_rows(xh, g.seg_xh)
    V = gather_rows(vec.reshape(-1, F3), g.seg_src).view(-1, 3, F)
    emb = radial_basis(geom[:, 3])
    phi = torch.zeros((g.n_edges, F3), dtype=xh.dtype, device=xh.device)
    for m in range(M):
        idx = g.module_edges(m)
        if idx.numel():
            phi = phi.index_copy(0, idx, emb[idx] @ Wt[m] + bias[m])
    a, b, c = torch.split(P * phi, F, dim=-1)
    m_vec = V * (b * (1 / math.sqrt(3.0)))[:, None, :] + c[:, None, :] * geom[:, :3, None]
    m_vec = m_vec * (1 / math.sqrt(F))
    live = (g.edge_mod >= 0).to(xh.dtype)[:, None]
    dx = segment_sum(a * live, g.seg_dst)
    dvec = segment_sum((m_vec * live[:, :, None]).reshape(-1, F3), g.seg_dst).view(-1, 3, F)
    return dx, dvec
