"""Fused node side of an HVNet layer (functional._XProjHV / _NodeUpdateHV over csrc/hn_node.cu + the GEMM epilogues of
csrc/hn_gemm.cu) against the plain torch formulation of the same layer -- energies, forces and cell gradients."""
import numpy as np
import pytest
import torch

import hermnet_b200 as H
from tests import util
from tests.util import lattice_system as _system


def _run(model, pos, Z, cell, fused_node):
    model.fused_node = fused_node
    data = H.Data(pos=pos.clone().requires_grad_(True), atomic_number=Z, cell=cell.clone().requires_grad_(True))
    e = model(data)
    gp, gc = torch.autograd.grad(e.sum(), [data.pos, data.cell])
    return e.detach(), gp, gc


def _model(elems, F, K, dev, layers=2):
    torch.manual_seed(3)
    m = H.HVNet(elems=elems, rc=5.0, num_layers=layers, hidden_channels=F, num_rbf=K).to(dev).eval()
    for p in m.parameters():
        p.requires_grad_(False)
    return m


def _compare(model, pos, Z, cell, tol_e=1e-5, tol_f=1e-4):
    e1, f1, c1 = _run(model, pos, Z, cell, True)
    e0, f0, c0 = _run(model, pos, Z, cell, False)
    assert float((e1 - e0).abs().max()) <= tol_e * max(1.0, float(e0.abs().max()))
    assert float((f1 - f0).abs().max()) <= tol_f * max(1.0, float(f0.abs().max()))
    assert float((c1 - c0).abs().max()) <= tol_f * max(1.0, float(c0.abs().max()))


@pytest.mark.parametrize("elems,zs,F,K", [(["H", "O"], [1, 8], 64, 32), (["H", "O", "C"], [1, 8, 7], 64, 20)])
def test_fused_node_path_matches_torch_path_on_cpu_emulation(emu, elems, zs, F, K):
    """zs may contain an element the model does not know (7) and the model an element that is absent (C)."""
    pos, Z, cell = _system(4, zs, 13)
    model = _model(elems, F, K, "cpu")
    assert model._fused_node_path(model.hermconvs[0], object(), model.build_graph(pos, Z, cell))
    _compare(model, pos, Z, cell)


def test_fused_node_path_is_skipped_when_parameters_need_gradients(emu):
    pos, Z, cell = _system(3, [1, 8], 13)
    model = _model(["H", "O"], 64, 32, "cpu")
    for p in model.parameters():
        p.requires_grad_(True)
    assert not model._fused_node_path(model.hermconvs[0], object(), model.build_graph(pos, Z, cell))


@pytest.mark.gpu
@pytest.mark.parametrize("elems,zs,F,K,n_side", [(["Li", "Al", "Si", "O"], [3, 13, 14, 8], 128, 128, 9),
                                                 (["H", "O", "C"], [1, 8, 7], 64, 20, 7),
                                                 (["Cr", "Fe"], [24, 26], 256, 64, 6)])
def test_fused_node_path_matches_torch_path(elems, zs, F, K, n_side):
    pos, Z, cell = _system(n_side, zs, 17)
    dev = "cuda:0"
    model = _model(elems, F, K, dev, layers=3)
    _compare(model, pos.to(dev), Z.to(dev), cell.to(dev))


@pytest.mark.gpu
def test_fused_node_path_on_the_golden_water_box():
    case = util.load_case("c1_hvnet")
    model, _ = util.make_model(case["kind"], case["cfg"], case["seed"], "cuda:0")
    for p in model.parameters():
        p.requires_grad_(False)
    data = util.make_data(case, "cuda:0", with_edges=False)
    assert model._fused_node_path(model.hermconvs[0], object(), model.build_graph(data.pos.detach(), data.atomic_number, data.cell.detach()))
    e, f, gc = util.energy_forces(model, data)
    assert util.rel_err(e.cpu(), case["energy"]) < 1e-5
    assert float((f.cpu() - case["forces"]).abs().max()) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("F", [64, 128])
def test_readout_kernels_match_float64(F):
    """hn_readout_{fwd,bwd} (plain fp32 readout MLP, hermnet.py:129) against a float64 evaluation."""
    from hermnet_b200 import ops
    gen = torch.Generator().manual_seed(F)
    n, Hd = 1000, F // 2
    x = torch.randn(n, F, generator=gen).cuda()
    W1 = (torch.randn(Hd, F, generator=gen) / F ** 0.5).cuda()
    b1 = (0.1 * torch.randn(Hd, generator=gen)).cuda()
    W2 = (torch.randn(Hd, generator=gen) / Hd ** 0.5).cuda()
    b2t = torch.tensor([0.3]).cuda()
    b2 = 0.3
    ge = torch.randn(n, generator=gen).cuda()
    e = ops.readout_fwd(x, W1, b1, W2, b2t)
    gx = ops.readout_bwd(x, W1, b1, W2, b2t, ge)
    xd = x.double().requires_grad_(True)
    ed = (torch.nn.functional.silu(xd @ W1.double().t() + b1.double()) / 0.6) @ W2.double() + b2
    (gd,) = torch.autograd.grad((ed * ge.double()).sum(), xd)
    assert float((e.squeeze(1).double() - ed).abs().max()) < 2e-6 * max(1.0, float(ed.abs().max()))
    assert float((gx.double() - gd).abs().max()) < 2e-6 * max(1.0, float(gd.abs().max()))


@pytest.mark.gpu
@pytest.mark.parametrize("n,F", [(1, 128), (1000, 128), (777, 64), (4099, 96), (513, 256), (300, 512)])
def test_layernorm_kernels_match_float64(n, F):
    """hn_layernorm_{fwd,bwd} (the affine-free nn.LayerNorm of rmnet.py:39,52) against float64 autograd."""
    from hermnet_b200 import ops
    g = torch.Generator().manual_seed(n + F)
    x = (torch.randn(n, F, generator=g) * torch.exp(torch.randn(n, 1, generator=g)) + torch.randn(n, 1, generator=g)).cuda()
    gy = torch.randn(n, F, generator=g).cuda()
    xd = x.double().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xd, (F,), None, None, 1e-5)
    (gref,) = torch.autograd.grad(ref, xd, gy.double())
    xhat, mean, rstd = ops.layernorm_fwd(x, 1e-5)
    assert float((xhat.double() - ref).abs().max()) < 2e-6 * max(1.0, float(ref.abs().max()))
    assert torch.allclose(mean.double(), xd.mean(1), rtol=1e-6, atol=1e-6)
    gx = ops.layernorm_bwd(gy, x, mean, rstd)
    assert float((gx.double() - gref).abs().max()) < 5e-6 * max(1.0, float(gref.abs().max()))
    t_hat, t_mean, t_rstd = torch.native_layer_norm(x, (F,), None, None, 1e-5)
    assert float((xhat - t_hat).abs().max()) < 2e-6 * max(1.0, float(t_hat.abs().max()))


def _run_l0(model, pos, Z, cell, on):
    model.layer0_basis = on
    return _run(model, pos, Z, cell, True)


@pytest.mark.parametrize("elems,zs,F,K", [(["H", "O"], [1, 8], 64, 32), (["H", "O", "C"], [1, 8, 7], 64, 20)])
def test_layer0_basis_aggregation_matches_the_edge_kernels_on_cpu_emulation(emu, monkeypatch, elems, zs, F, K):
    """First layer by basis aggregation + GEMM (functional._Layer0HV) == the element-table edge kernels, values and gradients."""
    from hermnet_b200 import ops
    pos, Z, cell = _system(4, zs, 13)
    model = _model(elems, F, K, "cpu")
    calls = []
    real = ops.layer0_basis_fwd
    monkeypatch.setattr(ops, "layer0_basis_fwd", lambda *a, **k: (calls.append(1), real(*a, **k))[1])
    e1, f1, c1 = _run_l0(model, pos, Z, cell, True)
    assert len(calls) == 1
    e0, f0, c0 = _run_l0(model, pos, Z, cell, False)
    assert len(calls) == 1
    assert float((e1 - e0).abs().max()) <= 1e-5 * max(1.0, float(e0.abs().max()))
    assert float((f1 - f0).abs().max()) <= 1e-4 * max(1.0, float(f0.abs().max()))
    assert float((c1 - c0).abs().max()) <= 1e-4 * max(1.0, float(c0.abs().max()))


@pytest.mark.gpu
@pytest.mark.parametrize("elems,zs,F,K,n_side", [(["Li", "Al", "Si", "O"], [3, 13, 14, 8], 128, 128, 9),
                                                 (["H", "O", "C"], [1, 8, 7], 64, 20, 7),
                                                 (["Cr", "Fe"], [24, 26], 256, 64, 6)])
def test_layer0_basis_aggregation_matches_the_edge_kernels(elems, zs, F, K, n_side):
    pos, Z, cell = _system(n_side, zs, 17)
    dev = "cuda:0"
    model = _model(elems, F, K, dev, layers=2)
    pos, Z, cell = pos.to(dev), Z.to(dev), cell.to(dev)
    e1, f1, c1 = _run_l0(model, pos, Z, cell, True)
    e0, f0, c0 = _run_l0(model, pos, Z, cell, False)
    assert float((e1 - e0).abs().max()) <= 1e-5 * max(1.0, float(e0.abs().max()))
    assert float((f1 - f0).abs().max()) <= 1e-4 * max(1.0, float(f0.abs().max()))
    assert float((c1 - c0).abs().max()) <= 1e-4 * max(1.0, float(c0.abs().max()))


@pytest.mark.gpu
def test_layer0_basis_kernels_match_the_emulation():
    """hn_layer0_basis_{fwd,bwd} against the torch restatement (tests/emulator.py) on a real graph, with a live mask."""
    from hermnet_b200 import ops
    from tests import emulator as E
    pos, Z, cell = _system(6, [3, 13, 14, 8], 5)
    dev = "cuda:0"
    model = _model(["Li", "Al", "Si", "O"], 128, 128, dev)
    g = model.build_graph(pos.to(dev), Z.to(dev), cell.to(dev))
    uniq, g0 = model._layer0_tables(g, Z.to(dev)[g.perm])
    geom = ops.edge_geom_fwd(pos.to(dev)[g.perm].contiguous(), cell.to(dev), g)
    p = ops.edge_params(g, g.n_modules, 128, 128, int(model.radial_basis.envelope.p), model.rc, model.radial_basis.rbf.coeff)
    gen = torch.Generator().manual_seed(1)
    live = (torch.rand(g.n_edges, generator=gen) > 0.2).to(torch.uint8).to(dev)
    off = model.radial_basis.rbf.offset
    nz, kp = int(uniq.numel()), ops.layer0_row_len(int(uniq.numel()), 128)
    cpu = lambda t: t.cpu() if torch.is_tensor(t) else t
    import copy
    gc = copy.copy(g0)
    for k in ("rowptr", "col", "row_mod", "edge_row"):
        setattr(gc, k, getattr(g0, k).cpu())
    for flags, lv in ((0, None), (1, live)):
        p.flags = flags
        Sa, Sc = ops.layer0_basis_fwd(p, g0, geom, lv, off, nz, kp)
        Sa_r, Sc_r = E.layer0_basis_fwd(p, gc, geom.cpu(), cpu(lv), off.cpu(), nz, kp)
        act = (g0.row_mod >= 0).cpu()
        assert float((Sa.cpu()[act] - Sa_r[act]).abs().max()) < 2e-5 * max(1.0, float(Sa_r.abs().max()))
        assert float((Sc.cpu()[act] - Sc_r[act]).abs().max()) < 2e-5 * max(1.0, float(Sc_r.abs().max()))
        gSa = torch.randn(Sa.shape, generator=gen).to(dev)
        gSc = torch.randn(Sc.shape, generator=gen).to(dev)
        gg = ops.layer0_basis_bwd(p, g0, geom, lv, off, nz, kp, gSa, gSc)
        gg_r = E.layer0_basis_bwd(p, gc, geom.cpu(), cpu(lv), off.cpu(), nz, kp, gSa.cpu(), gSc.cpu())
        assert float((gg.cpu() - gg_r).abs().max()) < 1e-4 * max(1.0, float(gg_r.abs().max()))
