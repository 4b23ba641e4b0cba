"""Parameter containers with the reference's layer names (reference src/models/net/layers.py).

The arithmetic of these layers runs inside the fused native trunk (str2str_b200/csrc); the classes exist so
that `state_dict()` keys, shapes and constructor signatures match the reference and a reference `.pth` loads
with `strict=True`.  Initialisation follows the reference's schemes (layers.py:34-125) with torch's own
truncated normal instead of scipy's.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

_TRUNC_STD = 0.8796256610342398  # std of a unit normal truncated to [-2, 2]


class Linear(nn.Linear):
    def __init__(self, in_dim: int, out_dim: int, bias: bool = True, init: str = "default"):
        super().__init__(in_dim, out_dim, bias=bias)
        with torch.no_grad():
            if bias:
                self.bias.fill_(0)
            if init in ("default", "relu"):
                std = math.sqrt((1.0 if init == "default" else 2.0) / max(1, in_dim)) / _TRUNC_STD
                nn.init.trunc_normal_(self.weight, std=std, a=-2 * std, b=2 * std)
            elif init == "final":
                self.weight.fill_(0.0)
            elif init == "gating":
                self.weight.fill_(0.0)
                if bias:
                    self.bias.fill_(1.0)
            elif init == "glorot":
                nn.init.xavier_uniform_(self.weight, gain=1)
            elif init == "normal":
                nn.init.kaiming_normal_(self.weight, nonlinearity="linear")
            else:
                raise ValueError("Invalid init string.")


class _FusedOnly(nn.Module):
    def forward(self, *args, **kwargs):  # pragma: no cover
        raise RuntimeError(
            f"{type(self).__name__} has no stand-alone kernel: it runs fused inside TranslationIPA.forward "
            "(str2str_b200/csrc/api.cu do_trunk)"
        )


class NodeTransition(_FusedOnly):
    def __init__(self, dim):
        super().__init__()
        self.dim = dim
        self.linear_1 = Linear(dim, dim, init="relu")
        self.linear_2 = Linear(dim, dim, init="relu")
        self.linear_3 = Linear(dim, dim, init="final")
        self.ln = nn.LayerNorm(dim)


class EdgeTransition(_FusedOnly):
    def __init__(self, node_embed_size, edge_embed_in, edge_embed_out, num_layers=2, node_dilation=2):
        super().__init__()
        bias_embed_size = node_embed_size // node_dilation
        self.initial_embed = Linear(node_embed_size, bias_embed_size, init="relu")
        hidden = bias_embed_size * 2 + edge_embed_in
        layers = []
        for _ in range(num_layers):
            layers += [Linear(hidden, hidden, init="relu"), nn.ReLU()]
        self.trunk = nn.Sequential(*layers)
        self.final_layer = Linear(hidden, edge_embed_out, init="final")
        self.layer_norm = nn.LayerNorm(edge_embed_out)


class TorsionAngleHead(_FusedOnly):
    def __init__(self, in_dim, n_torsion_angles, eps=1e-8):
        super().__init__()
        self.linear_1 = Linear(in_dim, in_dim, init="relu")
        self.linear_2 = Linear(in_dim, in_dim, init="relu")
        self.linear_3 = Linear(in_dim, in_dim, init="final")  # registered, never used (reference quirk)
        self.linear_final = Linear(in_dim, n_torsion_angles * 2, init="final")
        self.eps = eps


class BackboneUpdate(_FusedOnly):
    def __init__(self, c_s):
        super().__init__()
        self.c_s = c_s
        self.linear = Linear(c_s, 6, init="final")
