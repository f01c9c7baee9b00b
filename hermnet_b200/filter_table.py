"""Piecewise-polynomial table of the radial filter  f_c(u) = sum_k W[k][c] exp(coeff (u - offset_k)^2)  used by the
row-group edge kernels (csrc/hn_edge_group.cu).

Reference arithmetic: ``rbf_proj(envelope * GaussianSmearing(d/rc))`` (HermNet/rmnet.py:55,168-172; PyG GaussianSmearing
with ``offset = linspace(0, 1, K)``).  The Gaussian sum is a smooth function of the single scalar ``u``; on every grid
interval ``[offset_kc, offset_kc+1)`` it is represented by the degree-9 polynomial in ``s = 2 (u - offset_kc)(K-1) - 1``
that interpolates it at the 10 Chebyshev nodes.  The basis-function polynomials are fitted in float64 from the model's
actual ``offset`` buffer and ``coeff`` (all basis functions within 9 grid steps contribute; the rest are < e^-40), and
the table for a weight matrix is their float64 contraction with ``W`` rounded once to float32.  Measured against the
exact float64 sum the float32 Horner evaluation is within 8e-8 (values) / 1.8e-7 (derivative) of max|f| -- closer than
the reference's own float32 evaluation of the K exponentials (2.1e-7 / 2.8e-7); see tests/test_filter_table.py.
"""
from __future__ import annotations

import weakref

import numpy as np
import torch

DEGREE = 9
NCOEF = DEGREE + 1
HALF = 9          # basis functions kc-8 .. kc+9 contribute to interval kc

_BASIS_CACHE = {}
_TABLE_CACHE = {}


def basis_polynomials(offset: torch.Tensor, coeff: float) -> np.ndarray:
    """``B[kc, j, n]``: monomial coefficient ``n`` (in ``s``) of ``exp(coeff (u - offset_k)^2)``, ``k = kc - HALF + 1 + j``,
    on interval ``kc`` (zero rows where ``k`` falls outside ``[0, K)``).  float64, cached per (offset, coeff)."""
    from numpy.polynomial import chebyshev as C
    off = offset.detach().double().cpu().numpy()
    key = (off.tobytes(), float(coeff))
    hit = _BASIS_CACHE.get(key)
    if hit is not None:
        return hit
    K = off.shape[0]
    nodes = np.cos(np.pi * (np.arange(NCOEF) + 0.5) / NCOEF)
    B = np.zeros((K - 1, 2 * HALF, NCOEF))
    for kc in range(K - 1):
        u = off[kc] + ((nodes + 1.0) * 0.5) / (K - 1)
        for j in range(2 * HALF):
            k = kc - HALF + 1 + j
            if 0 <= k < K:
                B[kc, j] = C.cheb2poly(C.chebfit(nodes, np.exp(coeff * (u - off[k]) ** 2), DEGREE))
    if len(_BASIS_CACHE) > 16:
        _BASIS_CACHE.clear()
    _BASIS_CACHE[key] = B
    return B


def filter_table(Wt: torch.Tensor, offset: torch.Tensor, coeff: float) -> torch.Tensor:
    """``Wt [M, K, C]`` (rbf_proj.weight transposed, as the edge kernels take it) -> ``coef [M, K-1, NCOEF, C]`` float32."""
    M, K, Cc = Wt.shape
    B = torch.from_numpy(basis_polynomials(offset, coeff)).to(Wt.device)                 # [K-1, 2*HALF, NCOEF] f64
    out = torch.empty((M, K - 1, NCOEF, Cc), dtype=torch.float32, device=Wt.device)
    for m in range(M):
        Wp = torch.zeros((K + 2 * HALF, Cc), dtype=torch.float64, device=Wt.device)
        Wp[HALF - 1:HALF - 1 + K] = Wt[m].detach().double()
        win = Wp.unfold(0, 2 * HALF, 1)[: K - 1]                                          # [K-1, C, 2*HALF] (view)
        out[m] = torch.einsum("kjn,kcj->knc", B, win).to(torch.float32)
    return out


def cached_filter_table(weights, offset: torch.Tensor, coeff: float) -> torch.Tensor:
    """Table for the stacked ``rbf_proj`` weights of one layer (``weights``: list of ``[3F, K]`` parameters), rebuilt only
    when one of them changed (same weak-reference / version test as the TF32 weight splits)."""
    key = tuple(id(w) for w in weights)
    hit = _TABLE_CACHE.get(key)
    if hit is not None and all(r() is w and v == w._version for r, v, w in zip(hit[0], hit[1], weights)):
        return hit[2]
    Wt = torch.stack([w.detach().t() for w in weights]).contiguous()
    table = filter_table(Wt, offset, coeff)
    if len(_TABLE_CACHE) > 256:
        _TABLE_CACHE.clear()
    _TABLE_CACHE[key] = ([weakref.ref(w) for w in weights], [w._version for w in weights], table)
    return table
