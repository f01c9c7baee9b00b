"""Tensor-core edge kernels (csrc/hn_edge_tc.cu, the default for F == 128): tile-plan invariants (CPU emulation and
GPU), and forward / destination-major / source-major backward against a dense float64 evaluation of the reference
formula (rmnet.py:55-73 with the full K-term Gaussian sum of rmnet.py:168-193) and against the row-per-warp kernels on
identical inputs.  Covers HVNet / HPNet / HTNet row layouts, inactive rows, K = 20 / 128 / 200, graphs re-used after the
atoms moved (windows wider than one chunk, edges beyond the cutoff), the vec == NULL first-layer variants and the
element-table view of the first layer."""
import numpy as np
import pytest
import torch

from hermnet_b200 import ops, tileplan
from tests.util import edge_inputs as _edge_inputs, frozen_model as _model, lattice_system as _system


def check_plan(plan, g, geom, K, inv_rc):
    """Every live row-edge is in exactly one tile; tiles hold <= 64 edges of one block / group, sorted by
    (local % G, local); the window of every tile covers the 12-wide band of each of its in-range edges."""
    NG = ops.tc_groups()
    nt = plan.n_tiles
    info = plan.tile_info[:nt].cpu().long()
    win = plan.tile_win[:nt].cpu().long()
    erec = plan.erec.cpu().long()
    bt = plan.blk_tile.cpu().long()
    bi = plan.blk_info.cpu().long()
    row = g.edge_row.cpu().long()
    live = g.row_mod.cpu().long()[row] >= 0
    n = int(info[:, 1].sum()) if nt else 0
    assert n == int(live.sum())
    assert int(bt[-1]) == nt and bool((bt[1:] >= bt[:-1]).all())
    if plan.blk_order is not None:       # the processing order is a permutation of the blocks
        assert torch.equal(torch.sort(plan.blk_order.cpu().long()).values, torch.arange(plan.n_blocks))
    if nt == 0:
        return
    assert int(info[:, 1].max()) <= 64 and int(info[:, 1].min()) >= 1
    assert bool((info[:, 0] == torch.cumsum(info[:, 1], 0) - info[:, 1]).all())
    eids = erec[:n, 3]
    assert torch.equal(torch.sort(eids).values, torch.nonzero(live).squeeze(1))
    tile_of = torch.repeat_interleave(torch.arange(nt), info[:, 1])
    blk_of_tile = torch.repeat_interleave(torch.arange(plan.n_blocks), bt[1:] - bt[:-1])
    blk = blk_of_tile[tile_of]
    loc = erec[:n, 2]
    assert bool((loc >= 0).all()) and bool((loc < bi[blk, 2]).all())
    if plan.kind == "dst":
        assert torch.equal(plan.blk_xoff.cpu().long()[blk] + erec[:n, 1], erec[:n, 0])
        assert torch.equal(bi[blk, 0] + loc * bi[blk, 1], row[eids])
        assert torch.equal(erec[:n, 1], g.col.cpu().long()[eids])
        assert torch.equal(info[tile_of, 3], g.row_mod.cpu().long()[row[eids]])
    else:
        assert torch.equal(bi[blk, 0] + loc, g.col.cpu().long()[eids])
        assert torch.equal(erec[:n, 0], row[eids])
        assert torch.equal(info[tile_of, 3], g.row_mod.cpu().long()[row[eids]])
    assert torch.equal(erec[:n, 1 if plan.kind == "src" else 0], (g.row_xoff.cpu()[row[eids]] + g.col.cpu().long()[eids]))
    # order inside a tile, even count
    key = ((loc % NG) << 16) | loc
    same = tile_of[1:] == tile_of[:-1]
    assert bool((key[1:][same] >= key[:-1][same]).all())
    split = torch.zeros(nt, dtype=torch.long)
    for gq in range(NG - 1):
        split += torch.zeros(nt, dtype=torch.long).index_add_(0, tile_of, ((loc % NG) <= gq).long()) << (8 * gq)
    assert torch.equal(split, info[:, 2])
    # windows
    u = geom.cpu()[eids, 3] * inv_rc
    kc = torch.clamp((u * (K - 1)).to(torch.int64), max=K - 1)
    inr = u < 1
    lo, hi = win[tile_of, 0], win[tile_of, 0] + 32 * win[tile_of, 1]
    assert bool(((kc - 5).clamp(min=0)[inr] >= lo[inr]).all())
    assert bool(((kc + 6).clamp(max=K - 1)[inr] < hi[inr]).all())
    assert bool((win[:, 0] % 8 == 0).all()) and bool((win[:, 1] >= 1).all())


@pytest.mark.parametrize("kind,elems,zs,K", [("HVNet", ["Li", "Al", "Si", "O"], [3, 13, 14, 8], 128),
                                             ("HPNet", ["Li", "O"], [3, 8], 50),
                                             ("HTNet", ["H", "O"], [1, 8], 128),
                                             ("HVNet", ["H", "O"], [1, 8, 6], 20)])
def test_plan_invariants_on_cpu_emulation(emu, kind, elems, zs, K):
    pos, Z, cell = _system(4, zs, 5)
    model = _model(kind, elems, 128, K, "cpu")
    g = model.build_graph(pos, Z, cell)
    p, geom, *_ = _edge_inputs(model, g, pos, cell)
    dst, src = tileplan.plans_of(g, geom, p.inv_rc, K, want_src=True)
    for pl in (dst, src):
        pl.update_windows(geom, p.inv_rc, K)
        check_plan(pl, g, geom, K, p.inv_rc)
    # a fresh plan: at most window / 32 chunks per tile
    assert int(dst.tile_win[: dst.n_tiles, 1].max()) <= dst.window // 32 and int(src.tile_win[: src.n_tiles, 1].max()) <= src.window // 32


GPU_CASES = [
    ("HVNet", ["Li", "Al", "Si", "O"], [3, 13, 14, 8], 128, 7),
    ("HVNet", ["H", "O"], [1, 8], 20, 6),
    ("HPNet", ["Li", "O"], [3, 8], 50, 6),
    ("HTNet", ["H", "O"], [1, 8], 128, 6),
    ("HVNet", ["H", "O"], [1, 8, 6], 32, 6),       # atoms of an element the model does not know: inactive rows
    ("HVNet", ["Cr", "Fe", "Ni"], [24, 26, 28], 200, 6),
]


def _close(a, b, what, tol):
    scale = float(b.abs().max()) + 1e-12
    err = float((a - b).abs().max())
    assert err <= tol * scale, (what, err, scale)


def _run_tc(g, p, geom, xh, vec, Wt, bias, off, g_dx, g_dvec, null_vec=False):
    dst, src = tileplan.plans_of(g, geom, p.inv_rc, p.num_rbf, want_src=True)
    dst.update_windows(geom, p.inv_rc, p.num_rbf)
    src.update_windows(geom, p.inv_rc, p.num_rbf)
    check_plan(dst, g, geom, p.num_rbf, p.inv_rc)
    check_plan(src, g, geom, p.num_rbf, p.inv_rc)
    wsplit, wscale = ops.tc_split_weights(Wt)
    v = None if null_vec else vec
    dx, dv = ops.tc_edge_fwd(p, dst, xh, v, geom, wsplit, wscale, bias, off, p.n_rows)
    gg = ops.tc_edge_bwd_dst(p, dst, xh, v, geom, wsplit, wscale, bias, off, g_dx, g_dvec)[0]
    gxh, gvec = ops.tc_edge_bwd_src(p, src, xh, vec, geom, wsplit, wscale, bias, off, g_dx, g_dvec)
    return dx, dv, gg, gxh, gvec


def _run_row(g, p, geom, xh, vec, Wt, bias, off, g_dx, g_dvec, null_vec=False):
    ops.edge_set_variant("row")
    try:
        v = None if null_vec else vec
        dx, dv = ops.painn_edge_fwd(p, xh, v, geom, g, Wt, bias, off)
        gg = ops.painn_edge_bwd_dst(p, xh, v, geom, g, Wt, bias, off, g_dx, g_dvec).sum(0)
        gxh, gvec = ops.painn_edge_bwd_src(p, xh, vec, geom, g, Wt, bias, off, g_dx, g_dvec)
    finally:
        ops.edge_set_variant("auto")
    return dx, dv, gg, gxh, gvec


def _compare(model, g, pos, cell, tol=2e-5):
    from tests.test_sweep_kernels import _dense_reference
    args = _edge_inputs(model, g, pos, cell)
    row = _run_row(g, *args)
    tc = _run_tc(g, *args)
    ref = _dense_reference(g, *args)
    names = ("dx", "dvec", "g_geom", "grad_xh", "grad_vec")
    for n, a, b, r in zip(names, tc, row, ref):
        _close(a, b.view_as(a), n + " tc vs row", tol)
        _close(a.double(), r.view_as(a), n + " tc vs float64", tol)
        # the tensor-core path keeps the terms the 12-wide band drops: never (materially) worse than the row kernels
        eq = float((a.double() - r.view_as(a)).abs().max())
        er = float((b.view_as(a).double() - r.view_as(a)).abs().max())
        assert eq <= 3.0 * er + 2e-6 * float(r.abs().max()), (n, eq, er)
    # vec == NULL variants against the row kernels' NULL variants
    row0 = _run_row(g, *args, null_vec=True)
    tc0 = _run_tc(g, *args, null_vec=True)
    for n, a, b in zip(names[:3], tc0, row0):
        _close(a, b.view_as(a), n + " (vec = NULL) tc vs row", tol)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,elems,zs,K,n_side", GPU_CASES)
def test_tc_kernels_match_row_kernels_and_float64(kind, elems, zs, K, n_side):
    dev = "cuda:0"
    pos, Z, cell = _system(n_side, zs, 5)
    pos, Z, cell = pos.to(dev), Z.to(dev), cell.to(dev)
    model = _model(kind, elems, 128, K, dev)
    g = model.build_graph(pos, Z, cell)
    _compare(model, g, pos, cell)


@pytest.mark.gpu
def test_tc_kernels_after_the_atoms_moved():
    """Plan built for one configuration, kernels run for another: windows of several chunks, edges beyond the cutoff."""
    dev = "cuda:0"
    pos, Z, cell = _system(8, [3, 13, 14, 8], 31)
    pos, Z, cell = pos.to(dev), Z.to(dev), cell.to(dev)
    model = _model("HVNet", ["Li", "Al", "Si", "O"], 128, 128, dev)
    g = model.build_graph(pos, Z, cell)
    p, geom0, *_ = _edge_inputs(model, g, pos, cell)
    tileplan.plans_of(g, geom0, p.inv_rc, p.num_rbf, want_src=True)      # plans of the ORIGINAL geometry
    gen = torch.Generator().manual_seed(1)
    moved = pos + (torch.rand(pos.shape, generator=gen).to(dev) - 0.5) * 1.6
    d_new = ops.edge_geom_fwd(moved[g.perm].contiguous(), cell, g)[:, 3]
    assert int((d_new >= 5.0).sum()) > 100
    _compare(model, g, moved, cell)
    assert int(g._lazy["tc_dst"].tile_win[:, 1].max()) > 1


@pytest.mark.gpu
def test_tc_kernels_long_rows_dense_system():
    dev = "cuda:0"
    rng = np.random.default_rng(3)
    n_side, a = 8, 1.45
    grid = np.stack(np.meshgrid(*[np.arange(n_side)] * 3, indexing="ij"), -1).reshape(-1, 3)
    pos = torch.from_numpy((grid * a + rng.normal(0, 0.05, grid.shape)).astype(np.float32)).to(dev)
    Z = torch.from_numpy(rng.choice(np.array([1, 8]), size=len(grid))).long().to(dev)
    cell = torch.from_numpy((np.eye(3) * n_side * a).astype(np.float32)[None]).to(dev)
    model = _model("HVNet", ["H", "O"], 128, 128, dev)
    g = model.build_graph(pos, Z, cell)
    assert g.n_edges / g.n_atoms > 120
    _compare(model, g, pos, cell)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["triclinic_multi_image", "cluster_nonpbc_capped", "batch3_mixed"])
def test_tc_kernels_on_the_golden_geometries(name):
    from tests import util
    case = util.load_case(name)
    dev = "cuda:0"
    model = _model("HVNet", case["cfg"]["elems"], 128, 24, dev)
    pos, Z = case["pos"].to(dev), case["Z"].to(dev)
    cell = None if case["cell"] is None else case["cell"].to(dev)
    g = model.builder.from_positions(pos, Z, cell, case["batch"].to(dev))
    _compare(model, g, pos, cell)


@pytest.mark.gpu
def test_model_energy_forces_tc_equals_quad_and_is_deterministic():
    """Whole model (first-layer element table included) through the tensor-core kernels vs the tile-sweep kernels."""
    import hermnet_b200 as H
    dev = "cuda:0"
    pos, Z, cell = _system(9, [3, 13, 14, 8], 2)
    pos, Z, cell = pos.to(dev), Z.to(dev), cell.to(dev)
    model = _model("HVNet", ["Li", "Al", "Si", "O"], 128, 128, dev, layers=3)
    out = {}
    for variant in ("quad", "tc", "tc"):
        ops.edge_set_variant(variant)
        try:
            d = H.Data(pos=pos.clone().requires_grad_(True), atomic_number=Z, cell=cell.clone().requires_grad_(True))
            e = model(d)
            f, gc = torch.autograd.grad(e.sum(), [d.pos, d.cell])
        finally:
            ops.edge_set_variant("auto")
        out.setdefault(variant, []).append((e.detach(), f, gc))
    (e0, f0, c0), (e1, f1, c1), (e2, f2, c2) = out["quad"][0], out["tc"][0], out["tc"][1]
    assert float((e1 - e0).abs().max() / e0.abs().max()) < 1e-5
    assert float((f1 - f0).abs().max()) < 1e-4 * max(1.0, float(f0.abs().max()))
    assert float((c1 - c0).abs().max()) < 1e-3 * max(1.0, float(c0.abs().max()))
    assert torch.equal(e1, e2) and torch.equal(f1, f2) and torch.equal(c1, c2)      # no atomics: bit-deterministic
