"""Synthetic periodic systems of SURVEY.md 8(d) (numpy, seeded) -- inputs for tests and bench.py.

All generators return ``(pos float32 [N,3], Z int64 [N], cell float32 [3,3])``.  Positions are NOT
wrapped (the neighbour search accepts positions outside the cell, like ASE's).
"""
from __future__ import annotations

import numpy as np

from .symbols import atomic_numbers


def _random_rotations(rng, n):
    q = rng.standard_normal((n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    a, b, c, d = q.T
    return np.stack([
        np.stack([a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)], -1),
        np.stack([2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b)], -1),
        np.stack([2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d], -1)], 1)


def water_box(n_side: int = 4, seed: int = 0, density_g_cm3: float = 0.997, jitter: float = 0.05):
    """C1 / C3: ``n_side^3`` rigid H2O (0.9572 A, 104.52 deg) on a cubic grid, random orientation,
    N(0, jitter) noise, atom order O,H,H.  n_side=4 -> 192 atoms in a 12.417 A cell; 22 -> 31 944 atoms."""
    rng = np.random.default_rng(seed)
    n_mol = n_side ** 3
    mass_g = 18.01528 / 6.02214076e23
    cell_len = (n_mol * mass_g / density_g_cm3) ** (1 / 3) * 1e8
    a = cell_len / n_side
    g = (np.stack(np.meshgrid(*[np.arange(n_side)] * 3, indexing="ij"), -1).reshape(-1, 3) + 0.5) * a
    half = np.deg2rad(104.52) / 2
    mol = np.array([[0, 0, 0], [0.9572 * np.sin(half), 0.9572 * np.cos(half), 0],
                    [-0.9572 * np.sin(half), 0.9572 * np.cos(half), 0]])
    rot = _random_rotations(rng, n_mol)
    pos = g[:, None, :] + np.einsum("nij,aj->nai", rot, mol)
    pos = pos.reshape(-1, 3) + rng.normal(0, jitter, size=(3 * n_mol, 3))
    Z = np.tile(np.array([8, 1, 1]), n_mol)
    return pos.astype(np.float32), Z.astype(np.int64), (np.eye(3) * cell_len).astype(np.float32)


def cubic_lattice(n_side: int, a: float, species, probs=None, jitter: float = 0.10, seed: int = 0):
    """C2 / C4: simple-cubic sites with spacing ``a``, N(0, jitter) noise, iid species."""
    rng = np.random.default_rng(seed)
    idx = np.stack(np.meshgrid(*[np.arange(n_side)] * 3, indexing="ij"), -1).reshape(-1, 3)
    pos = idx * a + rng.normal(0, jitter, size=(n_side ** 3, 3))
    zs = np.array([atomic_numbers[s] for s in species])
    Z = rng.choice(zs, size=n_side ** 3, p=probs)
    return pos.astype(np.float32), Z.astype(np.int64), (np.eye(3) * n_side * a).astype(np.float32)


def fcc_alloy(n_cells: int, a: float = 3.6, species=("Cr", "Mn", "Fe", "Co", "Ni"), jitter: float = 0.05,
              seed: int = 5):
    """C5: fcc ``n_cells^3`` conventional cells, iid species."""
    rng = np.random.default_rng(seed)
    idx = np.stack(np.meshgrid(*[np.arange(n_cells)] * 3, indexing="ij"), -1).reshape(-1, 1, 3)
    basis = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]])[None]
    pos = ((idx + basis) * a).reshape(-1, 3)
    pos = pos + rng.normal(0, jitter, size=pos.shape)
    zs = np.array([atomic_numbers[s] for s in species])
    Z = rng.choice(zs, size=pos.shape[0])
    return pos.astype(np.float32), Z.astype(np.int64), (np.eye(3) * n_cells * a).astype(np.float32)


def config(name: str, scale: float = 1.0):
    """Named BASELINE.json configurations.  ``scale`` < 1 shrinks the box edge (tests)."""
    if name == "C1":
        return water_box(4, seed=0), dict(kind="HVNet", elems=["H", "O"], rc=5.0, num_layers=3,
                                          hidden_channels=128, num_rbf=128)
    if name == "C2":  # one graph of the batch; seeds 100..131 give the 32 graphs
        return cubic_lattice(max(2, int(round(16 * scale))), 2.3, ("Li", "Si", "O"), (1 / 3, 1 / 6, 1 / 2), 0.10, 100), \
            dict(kind="HPNet", elems=["Li", "Si", "O"], rc=5.0, num_layers=3, hidden_channels=128, num_rbf=128)
    if name == "C3":
        return water_box(max(2, int(round(22 * scale))), seed=3), \
            dict(kind="HTNet", elems=["H", "O"], rc=5.0, num_layers=3, hidden_channels=128, num_rbf=128)
    if name == "C4":
        return cubic_lattice(max(3, int(round(100 * scale))), 2.3, ("Li", "Al", "Si", "O"), None, 0.10, 4), \
            dict(kind="HVNet", elems=["Li", "Al", "Si", "O"], rc=5.0, num_layers=3, hidden_channels=128, num_rbf=128)
    if name == "C5":
        return fcc_alloy(max(2, int(round(40 * scale))), 3.6), \
            dict(kind="HTNet", elems=["Cr", "Mn", "Fe", "Co", "Ni"], rc=6.0, num_layers=3, hidden_channels=256,
                 num_rbf=128)
    raise ValueError(name)
