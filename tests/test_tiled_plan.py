"""TilePlan (graph.build_tile_plans) invariants and the tiled edge path against the row-per-warp path.

CPU part: host logic over the emulator (tests/emulator.py re-states csrc/hn_edge_tiled.cu slot by slot).
GPU part (``-m gpu``): the real tiled kernels against the row kernels of csrc/hn_edge.cu on identical inputs --
forward, both backward passes, a stale plan (graph reused after the atoms moved) and HPNet / HTNet row layouts.
"""
import numpy as np
import pytest
import torch

import hermnet_b200 as H
from hermnet_b200 import functional as Fn
from hermnet_b200 import ops
from hermnet_b200.graph import TILE_ROWS, edge_windows


def _system(n_side, elems_z, seed, a=2.3, jitter=0.1):
    rng = np.random.default_rng(seed)
    g = np.stack(np.meshgrid(*[np.arange(n_side)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float64)
    pos = (g * a + rng.normal(0, jitter, g.shape)).astype(np.float32)
    Z = rng.choice(np.array(elems_z), size=len(pos))
    cell = (np.eye(3) * n_side * a).astype(np.float32)[None]
    return torch.from_numpy(pos), torch.from_numpy(Z).long(), torch.from_numpy(cell)


def _model(kind, elems, F, K, dev, layers=2, seed=7):
    torch.manual_seed(seed)
    m = getattr(H, kind)(elems=elems, rc=5.0, num_layers=layers, hidden_channels=F, num_rbf=K).to(dev).eval()
    for p in m.parameters():
        p.requires_grad_(False)
    m.builder.tile_plans = True     # opt-in while the row kernels are still as fast (HERMNET_B200_TILED=1)
    m.builder.group_plans = True    # likewise HERMNET_B200_GROUP=1
    return m


def _check_plan(g, d, rc, K):
    dst, src = g.plan_dst, g.plan_src
    assert dst is not None and src is not None
    E = g.n_edges
    live = (g.row_mod.long()[g.edge_row.long()] >= 0)
    win = edge_windows(d, rc, K, dst.n_windows)
    for plan, per_tile in ((dst, dst.n_windows + 1), (src, g.n_modules * (dst.n_windows + 1))):
        assert plan.bptr.numel() == plan.n_tiles * per_tile + 1 and int(plan.bptr[-1]) == plan.n_pad
        sizes = plan.bptr[1:] - plan.bptr[:-1]
        assert bool((sizes % 2 == 0).all())
        meta = plan.meta.long()
        real = meta[:, 1] < TILE_ROWS
        assert bool((meta[~real, 1] == TILE_ROWS).all())
        eids = plan.eid.long()[real]
        assert eids.numel() == int(live.sum()) and torch.unique(eids).numel() == eids.numel()   # every live edge once
        assert bool(live[eids].all())
        pos_of = plan.pos_of.long()
        assert bool((pos_of[~live] == plan.n_pad).all())
        assert torch.equal(plan.eid.long()[pos_of[live]], torch.nonzero(live).squeeze(1))
        bucket = torch.repeat_interleave(torch.arange(sizes.numel(), device=sizes.device), sizes.long())
        assert torch.equal(bucket[real] % (dst.n_windows + 1), win[eids])                        # window of the bucket
        assert bool((meta[:, 2] == (g.row_xoff[g.edge_row.long()] + g.col.long())[plan.eid.long()]).all())
    meta = dst.meta.long()
    real = meta[:, 1] < TILE_ROWS
    sizes = dst.bptr[1:] - dst.bptr[:-1]
    bucket = torch.repeat_interleave(torch.arange(sizes.numel(), device=sizes.device), sizes.long())
    tile = bucket // (dst.n_windows + 1)
    rows = dst.tile_rows.long()[tile * TILE_ROWS + meta[:, 1].clamp(max=TILE_ROWS - 1)]
    assert torch.equal(rows[real], g.edge_row.long()[dst.eid.long()[real]])
    assert torch.equal(dst.tile_mod.long()[tile[real]], g.row_mod.long()[rows[real]])
    assert torch.equal(meta[real, 0], g.col.long()[dst.eid.long()[real]])
    ms = src.meta.long()
    reals = ms[:, 1] < TILE_ROWS
    sizes = src.bptr[1:] - src.bptr[:-1]
    bucket = torch.repeat_interleave(torch.arange(sizes.numel(), device=sizes.device), sizes.long())
    tile = bucket // (g.n_modules * (dst.n_windows + 1))
    mod = (bucket // (dst.n_windows + 1)) % g.n_modules
    e = src.eid.long()[reals]
    assert torch.equal((tile * TILE_ROWS + ms[:, 1])[reals], g.col.long()[e])
    assert torch.equal(ms[reals, 0], g.edge_row.long()[e])
    assert torch.equal(mod[reals], g.row_mod.long()[g.edge_row.long()[e]])


def _edge_inputs(model, g, pos, cell, seed=3):
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


def _compare_ops(model, g, pos, cell, tol=2e-5):
    p, geom, xh, vec, Wt, bias, off, g_dx, g_dvec = _edge_inputs(model, g, pos, cell)
    geom_b = ops.gather_rows(geom, g.plan_dst.eid)
    geom_s = ops.gather_rows(geom, g.plan_src.eid)
    dx0, dv0 = ops.painn_edge_fwd(p, xh, vec, geom, g, Wt, bias, off)
    dx1, dv1 = ops.painn_edge_fwd_tiled(p, xh, vec, geom_b, g.plan_dst, Wt, bias, off)

    def close(a, b, what):
        scale = float(b.abs().max()) + 1e-12
        assert float((a - b).abs().max()) <= tol * scale, (what, float((a - b).abs().max()), scale)

    close(dx1, dx0, "dx")
    close(dv1, dv0, "dvec")
    gg0 = ops.painn_edge_bwd_dst(p, xh, vec, geom, g, Wt, bias, off, g_dx, g_dvec).sum(0)
    gg1 = ops.gather_rows(ops.painn_edge_bwd_dst_tiled(p, xh, vec, geom_b, g.plan_dst, Wt, bias, off, g_dx, g_dvec).sum(0),
                          g.plan_dst.pos_of)
    close(gg1, gg0, "g_geom")
    gx0, gv0 = ops.painn_edge_bwd_src(p, xh, vec, geom, g, Wt, bias, off, g_dx, g_dvec)
    gx1, gv1 = ops.painn_edge_bwd_src_tiled(p, xh, vec, geom_s, g.plan_src, Wt, bias, off, g_dx, g_dvec)
    close(gx1, gx0, "grad_xh")
    close(gv1, gv0, "grad_vec")


CASES = [("HVNet", ["Li", "Al", "Si", "O"], [3, 13, 14, 8], 128, 128, 6),
         ("HVNet", ["H", "O"], [1, 8, 6], 64, 20, 5),           # an element unknown to the model -> inactive rows
         ("HPNet", ["Li", "Si", "O"], [3, 14, 8], 64, 32, 5),
         ("HTNet", ["H", "O"], [1, 8], 128, 50, 5)]


@pytest.mark.parametrize("kind,elems,zs,F,K,n_side", CASES)
def test_tile_plan_invariants_and_emulated_equivalence(emu, kind, elems, zs, F, K, n_side):
    pos, Z, cell = _system(min(n_side, 5), zs, 11)
    model = _model(kind, elems, F, K, "cpu")
    g = model.build_graph(pos, Z, cell)
    d = ops.edge_geom_fwd(pos[g.perm].contiguous(), cell, g)[:, 3]
    _check_plan(g, d, model.rc, K)
    _compare_ops(model, g, pos, cell, tol=1e-5)


def test_tiled_model_equals_row_model_on_cpu_emulation(emu):
    pos, Z, cell = _system(5, [1, 8], 5)
    model = _model("HVNet", ["H", "O"], 64, 32, "cpu")
    out = {}
    for tiled in (True, False):
        model.builder.tile_plans = tiled
        data = H.Data(pos=pos.clone().requires_grad_(True), atomic_number=Z, cell=cell.clone().requires_grad_(True))
        e = model(data)
        assert (data.graph.plan_dst is not None) == tiled
        gp, gc = torch.autograd.grad(e.sum(), [data.pos, data.cell])
        out[tiled] = (e.detach(), gp, gc)
    for a, b in zip(out[True], out[False]):
        assert float((a - b).abs().max()) <= 1e-5 * (float(b.abs().max()) + 1e-9)


# ------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("kind,elems,zs,F,K,n_side", CASES + [("HVNet", ["Cr", "Fe", "Ni"], [24, 26, 28], 256, 128, 8)])
def test_tiled_kernels_match_row_kernels(kind, elems, zs, F, K, n_side):
    pos, Z, cell = _system(n_side + 3, zs, 21)
    dev = "cuda:0"
    pos, Z, cell = pos.to(dev), Z.to(dev), cell.to(dev)
    model = _model(kind, elems, F, K, dev)
    g = model.build_graph(pos, Z, cell)
    d = ops.edge_geom_fwd(pos[g.perm].contiguous(), cell, g)[:, 3]
    _check_plan(g, d, model.rc, K)
    _compare_ops(model, g, pos, cell)


@pytest.mark.gpu
def test_tiled_kernels_with_a_stale_plan_and_beyond_cutoff_edges():
    """Graph (and plan) built for one configuration, evaluated after every atom moved by up to 0.4 A: some bands leave
    their window and some edges cross the cutoff -- the kernels must fall back per edge and still agree."""
    dev = "cuda:0"
    pos, Z, cell = _system(9, [3, 13, 14, 8], 31)
    pos, Z, cell = pos.to(dev), Z.to(dev), cell.to(dev)
    model = _model("HVNet", ["Li", "Al", "Si", "O"], 128, 128, dev)
    g = model.build_graph(pos, Z, cell)
    gen = torch.Generator().manual_seed(1)
    moved = pos + (torch.rand(pos.shape, generator=gen).to(dev) - 0.5) * 0.8
    d_new = ops.edge_geom_fwd(moved[g.perm].contiguous(), cell, g)[:, 3]
    win_old = edge_windows(ops.edge_geom_fwd(pos[g.perm].contiguous(), cell, g)[:, 3], 5.0, 128, g.plan_dst.n_windows)
    win_new = edge_windows(d_new, 5.0, 128, g.plan_dst.n_windows)
    assert int((win_old != win_new).sum()) > 1000 and int((d_new >= 5.0).sum()) > 100
    _compare_ops(model, g, moved, cell)


@pytest.mark.gpu
def test_tiled_path_is_the_default_and_matches_row_path_end_to_end():
    dev = "cuda:0"
    pos, Z, cell = _system(10, [3, 13, 14, 8], 41)
    pos, Z, cell = pos.to(dev), Z.to(dev), cell.to(dev)
    model = _model("HVNet", ["Li", "Al", "Si", "O"], 128, 128, dev, layers=3)
    out = {}
    for tiled in (True, False):
        model.builder.tile_plans = tiled
        data = H.Data(pos=pos.clone().requires_grad_(True), atomic_number=Z, cell=cell.clone().requires_grad_(True))
        e = model(data)
        assert (data.graph.plan_dst is not None) == tiled
        gp, gc = torch.autograd.grad(e.sum(), [data.pos, data.cell])
        out[tiled] = (e.detach(), gp, gc)
    e1, e0 = out[True][0], out[False][0]
    assert float((e1 - e0).abs().max()) <= 1e-5 * float(e0.abs().max())
    assert float((out[True][1] - out[False][1]).abs().max()) <= 1e-4 * max(1.0, float(out[False][1].abs().max()))
    assert float((out[True][2] - out[False][2]).abs().max()) <= 1e-4 * max(1.0, float(out[False][2].abs().max()))
