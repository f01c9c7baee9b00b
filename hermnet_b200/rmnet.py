"""Parameter containers of the modified-PaiNN sub-network and the radial basis.

API / ``state_dict`` compatible with ``/root/reference/HermNet/rmnet.py``:
``PaiNNModule.message_layer.{x_proj.0,x_proj.2,rbf_proj,x_layernorm}``, ``PaiNNModule.update_layer.{vec_proj,
xvec_proj.0,xvec_proj.2}``, ``RadialBasis.rbf.offset`` ...  Unlike the reference, the modules do not run the
edge path themselves: the model classes (``hermnet.py``) evaluate the node-side GEMMs on type slices with these
weights and hand the edge side to the fused CUDA kernels.
"""
from __future__ import annotations

import math

import numpy as np
import torch
from torch import nn

__all__ = ["PaiNNModule", "PaiNNMessage", "PaiNNUpdate", "ScaledSiLU", "AtomEmbedding", "RadialBasis",
           "PolynomialEnvelope", "ExponentialEnvelope", "GaussianSmearing", "SphericalBesselBasis", "BernsteinBasis"]


class ScaledSiLU(nn.Module):
    """rmnet.py:110-117."""

    def __init__(self):
        super().__init__()
        self.scale_factor = 1 / 0.6

    def forward(self, x):
        return nn.functional.silu(x) * self.scale_factor


class PaiNNMessage(nn.Module):
    """Weights of rmnet.py:35-49.  ``node_features`` is the node-side half of ``forward`` (rmnet.py:52);
    the edge-side half (rmnet.py:55-73) is ``hermnet_b200.functional.painn_edge``."""

    def __init__(self, hidden_channels, num_rbf):
        super().__init__()
        self.hidden_channels = hidden_channels
        self.x_proj = nn.Sequential(nn.Linear(hidden_channels, hidden_channels), ScaledSiLU(),
                                    nn.Linear(hidden_channels, hidden_channels * 3))
        self.rbf_proj = nn.Linear(num_rbf, hidden_channels * 3)
        self.inv_sqrt_3 = 1 / math.sqrt(3.0)
        self.inv_sqrt_h = 1 / math.sqrt(hidden_channels)
        self.x_layernorm = nn.LayerNorm(hidden_channels)

    def node_features(self, x):
        return self.x_proj(self.x_layernorm(x))


class PaiNNUpdate(nn.Module):
    """rmnet.py:79-107.  ``vdot`` overrides the <v1,v2> term (HTNet's triadic inner product)."""

    def __init__(self, hidden_channels):
        super().__init__()
        self.hidden_channels = hidden_channels
        self.vec_proj = nn.Linear(hidden_channels, hidden_channels * 2, bias=False)
        self.xvec_proj = nn.Sequential(nn.Linear(hidden_channels * 2, hidden_channels), ScaledSiLU(),
                                       nn.Linear(hidden_channels, hidden_channels * 3))
        self.inv_sqrt_2 = 1 / math.sqrt(2.0)
        self.inv_sqrt_h = 1 / math.sqrt(hidden_channels)

    def forward(self, x, vec, vdot=None, lin=None):
        """``lin(input, nn.Linear)`` lets the model route the dense layers to the tensor-core GEMM."""
        if lin is None:
            lin = lambda t, layer: layer(t)
        F = self.hidden_channels
        v1, v2 = torch.split(lin(vec, self.vec_proj), F, dim=-1)
        if vdot is None:
            vdot = (v1 * v2).sum(dim=1) * self.inv_sqrt_h
        vnorm = torch.sqrt(torch.sum(v2 ** 2, dim=-2) + 1e-8)
        h = self.xvec_proj[1](lin(torch.cat([x, vnorm], dim=-1), self.xvec_proj[0]))
        a1, a2, a3 = torch.split(lin(h, self.xvec_proj[2]), F, dim=-1)
        return (a1 + a2 * vdot) * self.inv_sqrt_2, a3.unsqueeze(1) * v1


class PaiNNModule(nn.Module):
    """rmnet.py:11-32 container (``message_layer`` + ``update_layer``)."""

    def __init__(self, hidden_channels=512, num_rbf=128):
        super().__init__()
        self.num_rbf = num_rbf
        self.message_layer = PaiNNMessage(hidden_channels, num_rbf)
        self.update_layer = PaiNNUpdate(hidden_channels)
        self.inv_sqrt_2 = 1 / math.sqrt(2.0)

    def node_update(self, x, vec, dx, dvec, vdot=None, lin=None):
        """rmnet.py:24-32 on the rows of one sub-network: residual, rescale, update block.  Returns (vec, x)."""
        x = (x + dx) * self.inv_sqrt_2
        vec = vec + dvec
        dx2, dvec2 = self.update_layer(x, vec, vdot, lin)
        return vec + dvec2, x + dx2


class AtomEmbedding(nn.Module):
    """rmnet.py:120-131 (unused by the models, kept for namespace compatibility)."""

    def __init__(self, emb_size, num_elements):
        super().__init__()
        self.emb_size = emb_size
        self.embeddings = nn.Embedding(num_elements, emb_size)
        nn.init.uniform_(self.embeddings.weight, a=-math.sqrt(3), b=math.sqrt(3))

    def forward(self, Z):
        return self.embeddings(Z - 1)


class GaussianSmearing(nn.Module):
    """PyG ``GaussianSmearing`` restated [upstream]: buffer ``offset = linspace(start, stop, K)``."""

    def __init__(self, start=0.0, stop=5.0, num_gaussians=50):
        super().__init__()
        offset = torch.linspace(start, stop, num_gaussians)
        self.coeff = -0.5 / (offset[1] - offset[0]).item() ** 2
        self.register_buffer("offset", offset)
        self._linspace = (float(start), float(stop), int(num_gaussians))

    def forward(self, dist):
        dist = dist.view(-1, 1) - self.offset.view(1, -1)
        return torch.exp(self.coeff * dist.pow(2))

    def is_standard(self) -> bool:
        """True iff ``offset`` still equals linspace(0, 1, K) -- the precondition of the banded fused kernel."""
        start, stop, k = self._linspace
        return start == 0.0 and stop == 1.0 and bool(
            torch.equal(self.offset.detach().cpu().float(), torch.linspace(0.0, 1.0, k)))


class PolynomialEnvelope(nn.Module):
    """rmnet.py:175-193."""

    def __init__(self, exponent):
        super().__init__()
        assert exponent > 0
        self.p = exponent
        self.a = -(self.p + 1) * (self.p + 2) / 2
        self.b = self.p * (self.p + 2)
        self.c = -self.p * (self.p + 1) / 2

    def forward(self, d_scaled):
        env = 1 + self.a * d_scaled ** self.p + self.b * d_scaled ** (self.p + 1) + self.c * d_scaled ** (self.p + 2)
        return torch.where(d_scaled < 1, env, torch.zeros_like(d_scaled))


class ExponentialEnvelope(nn.Module):
    """rmnet.py:196-208."""

    def forward(self, d_scaled):
        env = torch.exp(-(d_scaled ** 2) / ((1 - d_scaled) * (1 + d_scaled)))
        return torch.where(d_scaled < 1, env, torch.zeros_like(d_scaled))


class SphericalBesselBasis(nn.Module):
    """rmnet.py:211-233."""

    def __init__(self, num_radial: int, cutoff: float):
        super().__init__()
        self.norm_const = math.sqrt(2 / (cutoff ** 3))
        self.frequencies = nn.Parameter(math.pi * torch.arange(1, num_radial + 1).float(), requires_grad=True)

    def forward(self, d_scaled):
        return self.norm_const / d_scaled[:, None] * torch.sin(self.frequencies * d_scaled[:, None])


class BernsteinBasis(nn.Module):
    """rmnet.py:236-275."""

    def __init__(self, num_radial: int, pregamma_initial: float = 0.45264):
        super().__init__()
        from scipy.special import binom
        self.register_buffer("prefactor", torch.tensor(binom(num_radial - 1, np.arange(num_radial)), dtype=torch.float),
                             persistent=False)
        self.pregamma = nn.Parameter(torch.tensor(pregamma_initial, dtype=torch.float), requires_grad=True)
        self.softplus = nn.Softplus()
        exp1 = torch.arange(num_radial)
        self.register_buffer("exp1", exp1[None, :], persistent=False)
        self.register_buffer("exp2", (num_radial - 1 - exp1)[None, :], persistent=False)

    def forward(self, d_scaled):
        gamma = self.softplus(self.pregamma)
        exp_d = torch.exp(-gamma * d_scaled)[:, None]
        return self.prefactor * (exp_d ** self.exp1) * ((1 - exp_d) ** self.exp2)


class RadialBasis(nn.Module):
    """rmnet.py:134-172: ``envelope(d/rc)[:, None] * rbf(d/rc)``; same constructor, errors and state keys."""

    def __init__(self, num_radial, cutoff, rbf={"name": "gaussian"}, envelope={"name": "polynomial", "exponent": 5}):
        super().__init__()
        self.inv_cutoff = 1 / cutoff
        self.cutoff = float(cutoff)
        self.num_radial = num_radial
        env_name = envelope["name"].lower()
        env_hparams = {k: v for k, v in envelope.items() if k != "name"}
        if env_name == "polynomial":
            self.envelope = PolynomialEnvelope(**env_hparams)
        elif env_name == "exponential":
            self.envelope = ExponentialEnvelope(**env_hparams)
        else:
            raise ValueError(f"Unknown envelope function '{env_name}'.")
        rbf_name = rbf["name"].lower()
        rbf_hparams = {k: v for k, v in rbf.items() if k != "name"}
        if rbf_name == "gaussian":
            self.rbf = GaussianSmearing(start=0, stop=1, num_gaussians=num_radial, **rbf_hparams)
        elif rbf_name == "spherical_bessel":
            self.rbf = SphericalBesselBasis(num_radial=num_radial, cutoff=cutoff, **rbf_hparams)
        elif rbf_name == "bernstein":
            self.rbf = BernsteinBasis(num_radial=num_radial, **rbf_hparams)
        else:
            raise ValueError(f"Unknown radial basis function '{rbf_name}'.")
        self._fusable = None

    def forward(self, d):
        d_scaled = d * self.inv_cutoff
        return self.envelope(d_scaled)[:, None] * self.rbf(d_scaled)

    def fusable(self) -> bool:
        """The fused edge kernel evaluates Gaussian RBF x polynomial envelope in registers (the default and
        benchmarked combination); every other combination runs through the composite formulation."""
        if self._fusable is None:
            self._fusable = (isinstance(self.rbf, GaussianSmearing) and isinstance(self.envelope, PolynomialEnvelope)
                             and float(self.envelope.p).is_integer() and self.num_radial >= 2 and self.rbf.is_standard())
        return self._fusable

    def _load_from_state_dict(self, *args, **kwargs):
        self._fusable = None
        return super()._load_from_state_dict(*args, **kwargs)
