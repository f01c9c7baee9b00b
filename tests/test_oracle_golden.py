"""The oracle against the committed golden fixtures (tests/golden, produced by the reference's own source run over
oracle/ref_shims.py -- see tests/golden/make_golden.py) and the oracle's internal consistency.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import build as obuild
from oracle import hermnet_oracle as O
from oracle import neighbor_oracle as NO
from tests import util


@pytest.mark.parametrize("name", util.ALL_CASES)
def test_oracle_reproduces_golden(name):
    c = util.load_case(name)
    sd = O.make_state_dict(c["kind"], c["cfg"], c["seed"])
    out = O.energy_and_forces(c["kind"], sd, c["cfg"], c["pos"], c["Z"], c["edge_index"], c["cell"], c["edge_shift"],
                              c["batch"], want_cell_grad=c["cell"] is not None)
    assert util.rel_err(out[0], c["energy"]) < 2e-6
    assert float((out[1] - c["forces"]).abs().max()) < 2e-5
    if c["cell"] is not None:
        assert float((out[2] - c["cell_grad"]).abs().max()) < 2e-4


@pytest.mark.parametrize("name", ["triclinic_multi_image", "batch3_mixed"])
def test_literal_in_subgraph_loop_equals_vectorised(name):
    c = util.load_case(name)
    sd = O.make_state_dict(c["kind"], c["cfg"], c["seed"])
    args = (sd, c["cfg"], c["pos"], c["Z"], c["edge_index"], c["cell"], c["edge_shift"], c["batch"])
    assert torch.equal(O.hvnet_forward(*args, literal_subgraph=True), O.hvnet_forward(*args))


@pytest.mark.parametrize("name", [n for n in util.ALL_CASES if "nonpbc" not in n])
def test_neighbor_oracle_reproduces_golden_edges(name):
    c = util.load_case(name)
    rows, off = [], 0
    for g in range(int(c["batch"].max()) + 1):
        sel = (c["batch"] == g).numpy()
        i, j, S = NO.neighbor_list_pbc(c["pos"].numpy()[sel], c["cell"][g].numpy(), c["cfg"]["rc"])
        rows.append(NO.canonical_edges(np.stack([i + off, j + off]), S))
        off += int(sel.sum())
    got = np.concatenate(rows)
    got = got[np.lexsort((got[:, 4], got[:, 3], got[:, 2], got[:, 1], got[:, 0]))]
    assert np.array_equal(got, c["edges"])


@pytest.mark.parametrize("name", ["c1_hvnet", "triclinic_multi_image"])
def test_c_oracle_equals_numpy_oracle(name):
    c = util.load_case(name)
    i, j, S = obuild.nl_pbc_rows(c["pos"].numpy(), c["cell"][0].numpy(), c["cfg"]["rc"])
    assert np.array_equal(NO.canonical_edges(np.stack([i, j]), S), c["edges"])


def test_neighbor_list_is_symmetric_and_strict():
    c = util.load_case("triclinic_multi_image")
    e = c["edges"]
    fwd = {tuple(r) for r in e.tolist()}
    assert {(s, d, -a, -b, -cc) for (d, s, a, b, cc) in fwd} == fwd     # (i,j,S) <=> (j,i,-S)
    assert not any(d == s and a == b == cc == 0 for (d, s, a, b, cc) in fwd)
    assert any(d == s for (d, s, a, b, cc) in fwd)                       # self-image edges exist (rc > L/2)


def test_htnet_factorised_equals_explicit_triplets():
    c = util.load_case("water24_htnet")
    sd = O.make_state_dict(c["kind"], c["cfg"], c["seed"])
    args = (sd, c["cfg"], c["pos"], c["Z"], c["edge_index"], c["cell"], c["edge_shift"], c["batch"])
    a = O.htnet_forward(*args)
    b = O.htnet_forward(*args, explicit_triplets=True)
    assert util.rel_err(b, a) < 1e-5


def test_hpnet_reduces_to_hvnet_on_one_element():
    rng = np.random.default_rng(3)
    cell = torch.eye(3) * 6.0
    pos = torch.from_numpy(rng.uniform(0, 6, (10, 3)).astype(np.float32))
    Z = torch.full((10,), 14)
    i, j, S = NO.neighbor_list_pbc(pos.numpy(), cell.numpy(), 4.0)
    ei, es = torch.from_numpy(np.stack([i, j])), torch.from_numpy(S.astype(np.float32))
    cfg = dict(elems=["Si"], rc=4.0, num_layers=2, hidden_channels=32, num_rbf=16)
    sd_v = O.make_state_dict("HVNet", cfg, 5)
    sd_p = {k.replace(".mods.Si.", ".mods.Si-Si."): v for k, v in sd_v.items()}
    ev = O.hvnet_forward(sd_v, cfg, pos, Z, ei, cell[None], es)
    ep = O.hpnet_forward(sd_p, cfg, pos, Z, ei, cell[None], es)
    assert torch.equal(ev, ep)


def test_oracle_symmetries_fp64():
    """Rotation / translation / permutation invariance of E and equivariance of F, finite-difference forces (fp64)."""
    rng = np.random.default_rng(9)
    n = 12
    pos = torch.from_numpy(rng.normal(0, 1.5, (n, 3)))
    Z = torch.from_numpy(rng.choice([1, 8], n))
    ei = torch.from_numpy(NO.radius_graph_nonpbc(pos.numpy(), 4.0))
    cfg = dict(elems=["H", "O"], rc=4.0, num_layers=2, hidden_channels=16, num_rbf=12)
    sd = {k: (v.double() if v.is_floating_point() else v) for k, v in O.make_state_dict("HVNet", cfg, 2).items()}
    e0, f0 = O.energy_and_forces("HVNet", sd, cfg, pos, Z, ei)
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    Rm = torch.from_numpy(q)
    e1, f1 = O.energy_and_forces("HVNet", sd, cfg, pos @ Rm.T + 0.7, Z, ei)
    assert torch.allclose(e0, e1, atol=1e-10) and torch.allclose(f0 @ Rm.T, f1, atol=1e-9)
    perm = torch.from_numpy(rng.permutation(n))
    inv = torch.empty_like(perm); inv[perm] = torch.arange(n)
    e2, f2 = O.energy_and_forces("HVNet", sd, cfg, pos[perm], Z[perm], inv[ei])
    assert torch.allclose(e0, e2, atol=1e-10) and torch.allclose(f0[perm], f2, atol=1e-9)
    h = 1e-5
    for (a, k) in [(0, 0), (5, 2)]:
        pp, pm = pos.clone(), pos.clone()
        pp[a, k] += h; pm[a, k] -= h
        ep = O.hvnet_forward(sd, cfg, pp, Z, ei); em = O.hvnet_forward(sd, cfg, pm, Z, ei)
        assert abs(float(-(ep - em) / (2 * h)) - float(f0[a, k])) < 1e-6


def test_binned_neighbour_list_equals_the_brute_force_oracle():
    """The production-cost CPU list of bench.py's baseline leg == the brute-force restatement (same exact test)."""
    from hermnet_b200 import synthetic
    from oracle import neighbor_oracle as NO
    pos, Z, cell = synthetic.cubic_lattice(6, 2.3, ("Li", "Al", "Si", "O"), None, 0.10, 4)
    pos = pos + np.float32(7.5)          # atoms outside the cell: wrapped candidates, original-position test
    a = NO.neighbor_list_pbc(pos, cell, 5.0)
    b = NO.neighbor_list_pbc_binned(pos, cell, 5.0)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    case_pos, _, case_cell = synthetic.water_box(4, seed=0)
    a = NO.neighbor_list_pbc(case_pos, case_cell, 5.0)
    b = NO.neighbor_list_pbc_binned(case_pos, case_cell, 5.0)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
