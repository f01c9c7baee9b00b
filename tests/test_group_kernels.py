"""Row-group edge kernels (csrc/hn_edge_group.cu) and their piecewise-polynomial filter table (hermnet_b200/filter_table.py):
table accuracy against the exact float64 Gaussian sum, GroupPlan invariants, and forward / destination-major backward
against the row-per-warp kernels of csrc/hn_edge.cu (which evaluate the reference formula term by term)."""
import numpy as np
import pytest
import torch

import hermnet_b200 as H
from hermnet_b200 import filter_table, ops
from hermnet_b200.graph import GROUP_ROWS
from tests.test_tiled_plan import CASES, _edge_inputs, _model, _system


@pytest.mark.parametrize("K", [16, 20, 50, 128])
def test_filter_table_is_closer_to_the_exact_sum_than_the_fp32_reference_formula(K):
    offset = torch.linspace(0, 1, K)
    coeff = -0.5 / (offset[1] - offset[0]).item() ** 2
    torch.manual_seed(K)
    Wt = torch.randn(2, K, 12) / K ** 0.5
    tab = filter_table.filter_table(Wt, offset, coeff)
    assert tab.shape == (2, K - 1, filter_table.NCOEF, 12) and tab.dtype == torch.float32
    u = torch.rand(50000) * 0.9999
    kc = (u * (K - 1)).long().clamp(0, K - 2)
    s = 2 * (K - 1) * (u - offset[kc]) - 1
    c = tab[1][kc]
    p, q = c[:, -1], torch.zeros_like(c[:, -1])
    for n in range(filter_table.NCOEF - 2, -1, -1):
        q = q * s[:, None] + p
        p = p * s[:, None] + c[:, n]
    du = u.double()[:, None] - offset.double()[None, :]
    G = torch.exp(coeff * du ** 2)
    exact, dexact = G @ Wt[1].double(), (G * 2 * coeff * du) @ Wt[1].double()
    d32 = u[:, None] - offset[None, :]
    g32 = torch.exp(coeff * d32 * d32)
    ref, dref = g32 @ Wt[1], (g32 * 2 * coeff * d32) @ Wt[1]
    err = float((p.double() - exact).abs().max() / exact.abs().max())
    derr = float((q.double() * 2 * (K - 1) - dexact).abs().max() / dexact.abs().max())
    assert err < 3e-7 and derr < 1e-6, (err, derr)
    assert err < 2 * float((ref.double() - exact).abs().max() / exact.abs().max()) + 1e-7
    assert derr < 2 * float((dref.double() - dexact).abs().max() / dexact.abs().max()) + 2e-7


def _check_group_plan(g, d, rc, K):
    plan = g.plan_grp
    assert plan is not None
    live = g.row_mod.long()[g.edge_row.long()] >= 0
    assert plan.n_slots == int(live.sum()) and int(plan.gptr[-1]) == plan.n_slots and plan.gptr.numel() == plan.n_groups + 1
    eid = plan.eid.long()
    assert torch.unique(eid).numel() == eid.numel() and bool(live[eid].all())
    pos_of = plan.pos_of.long()
    assert bool((pos_of[~live] == plan.n_slots).all()) and torch.equal(eid[pos_of[live]], torch.nonzero(live).squeeze(1))
    lens = (plan.gptr[1:] - plan.gptr[:-1]).long()
    group = torch.repeat_interleave(torch.arange(plan.n_groups, device=lens.device), lens)
    meta = plan.meta.long()
    rows = plan.group_rows.long()[group * GROUP_ROWS + meta[:, 1]]
    assert torch.equal(rows, g.edge_row.long()[eid]) and torch.equal(meta[:, 0], g.col.long()[eid])
    assert torch.equal(plan.group_mod.long()[group], g.row_mod.long()[rows])
    assert bool((meta[:, 2] == (g.row_xoff[g.edge_row.long()] + g.col.long())[eid]).all())
    key = group * K + meta[:, 3]
    assert bool((key[1:] >= key[:-1]).all())                      # sorted by (group, interval)
    u = d[eid] * torch.tensor(1.0 / rc, dtype=torch.float32, device=d.device)
    kc = torch.where(u < 1, (u * float(K - 1)).long().clamp(0, K - 2), torch.full_like(meta[:, 3], K - 1))
    assert torch.equal(kc, meta[:, 3])


def _compare_group_ops(model, g, pos, cell, tol=2e-5):
    p, geom, xh, vec, Wt, bias, off, g_dx, g_dvec = _edge_inputs(model, g, pos, cell)
    coef = filter_table.filter_table(Wt, off, model.radial_basis.rbf.coeff)
    geom_g = ops.gather_rows(geom, g.plan_grp.eid)

    def close(a, b, what):
        scale = float(b.abs().max()) + 1e-12
        assert float((a - b).abs().max()) <= tol * scale, (what, float((a - b).abs().max()), scale)

    dx0, dv0 = ops.painn_edge_fwd(p, xh, vec, geom, g, Wt, bias, off)
    dx1, dv1 = ops.painn_edge_fwd_group(p, xh, vec, geom_g, g.plan_grp, coef, bias, off)
    close(dx1, dx0, "dx")
    close(dv1, dv0, "dvec")
    gg0 = ops.painn_edge_bwd_dst(p, xh, vec, geom, g, Wt, bias, off, g_dx, g_dvec).sum(0)
    gg1 = ops.gather_rows(ops.painn_edge_bwd_dst_group(p, xh, vec, geom_g, g.plan_grp, coef, bias, off, g_dx, g_dvec).sum(0),
                          g.plan_grp.pos_of)
    close(gg1, gg0, "g_geom")


@pytest.mark.parametrize("kind,elems,zs,F,K,n_side", CASES)
def test_group_plan_invariants_and_emulated_equivalence(emu, kind, elems, zs, F, K, n_side):
    pos, Z, cell = _system(min(n_side, 5), zs, 11)
    model = _model(kind, elems, F, K, "cpu")
    g = model.build_graph(pos, Z, cell)
    d = ops.edge_geom_fwd(pos[g.perm].contiguous(), cell, g)[:, 3]
    _check_group_plan(g, d, model.rc, K)
    _compare_group_ops(model, g, pos, cell, tol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,elems,zs,F,K,n_side", CASES + [("HVNet", ["Cr", "Fe", "Ni"], [24, 26, 28], 256, 128, 8)])
def test_group_kernels_match_row_kernels(kind, elems, zs, F, K, n_side):
    pos, Z, cell = _system(n_side + 3, zs, 21)
    dev = "cuda:0"
    pos, Z, cell = pos.to(dev), Z.to(dev), cell.to(dev)
    model = _model(kind, elems, F, K, dev)
    g = model.build_graph(pos, Z, cell)
    d = ops.edge_geom_fwd(pos[g.perm].contiguous(), cell, g)[:, 3]
    _check_group_plan(g, d, model.rc, K)
    _compare_group_ops(model, g, pos, cell)


@pytest.mark.gpu
def test_group_kernels_with_a_stale_plan():
    """Plan built for one configuration, kernels run after every atom moved by up to 0.4 A (intervals change, edges cross
    the cutoff): the interval is re-derived from the current distance, results must not depend on the plan."""
    dev = "cuda:0"
    pos, Z, cell = _system(9, [3, 13, 14, 8], 31)
    pos, Z, cell = pos.to(dev), Z.to(dev), cell.to(dev)
    model = _model("HVNet", ["Li", "Al", "Si", "O"], 128, 128, dev)
    g = model.build_graph(pos, Z, cell)
    gen = torch.Generator().manual_seed(1)
    moved = pos + (torch.rand(pos.shape, generator=gen).to(dev) - 0.5) * 0.8
    d_new = ops.edge_geom_fwd(moved[g.perm].contiguous(), cell, g)[:, 3]
    assert int((d_new >= 5.0).sum()) > 100
    _compare_group_ops(model, g, moved, cell)


@pytest.mark.gpu
def test_group_path_is_the_default_and_matches_the_row_path_end_to_end():
    dev = "cuda:0"
    pos, Z, cell = _system(10, [3, 13, 14, 8], 41)
    pos, Z, cell = pos.to(dev), Z.to(dev), cell.to(dev)
    model = _model("HVNet", ["Li", "Al", "Si", "O"], 128, 128, dev, layers=3)
    model.builder.tile_plans = False
    out = {}
    for grp in (True, False):
        model.builder.group_plans = grp
        data = H.Data(pos=pos.clone().requires_grad_(True), atomic_number=Z, cell=cell.clone().requires_grad_(True))
        e = model(data)
        assert (data.graph.plan_grp is not None) == grp
        gp, gc = torch.autograd.grad(e.sum(), [data.pos, data.cell])
        out[grp] = (e.detach(), gp, gc)
    assert float((out[True][0] - out[False][0]).abs().max()) <= 1e-5 * float(out[False][0].abs().max())
    assert float((out[True][1] - out[False][1]).abs().max()) <= 1e-4 * max(1.0, float(out[False][1].abs().max()))
    assert float((out[True][2] - out[False][2]).abs().max()) <= 1e-4 * max(1.0, float(out[False][2].abs().max()))
