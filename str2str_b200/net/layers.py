"""The reference's trunk layers (reference src/models/net/layers.py): same names, constructor signatures, `state_dict()`
keys and `forward` signatures, so a reference `.pth` loads with `strict=True` and each module can be called on its own.

Inside `TranslationIPA.forward` these layers run fused in the native trunk (str2str_b200/csrc/api.cu: do_trunk); called
directly, each `forward` goes through its own C-ABI entry point (s2s_node_transition / s2s_edge_transition /
s2s_torsion_head / s2s_backbone_update) with the weights of the owning `DenoisingNet`'s native context.  Initialisation
follows the reference's schemes (layers.py:34-125) with torch's own truncated normal instead of scipy's.
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


class _Native(nn.Module):
    """Base of the trunk layers: finds the owning DenoisingNet's native engine and this layer's block index."""

    _s2s_block = None  # set by TranslationIPA.__init__

    def _engine(self, x: torch.Tensor, rows: int):
        from .ipa import _engine_of

        eng = _engine_of(self).native(x.device)
        eng.reserve(1, max(rows, 1))  # shape only: B * L >= rows
        return eng


class NodeTransition(_Native):
    def __init__(self, dim):
        super().__init__()
        self.dim = dim
        self.linear_1 = Linear(dim, dim, init="relu")
        self.linear_2 = Linear(dim, dim, init="relu")
        self.linear_3 = Linear(dim, dim, init="final")
        self.ln = nn.LayerNorm(dim)

    def forward(self, s):
        """LN(s + linear_3(relu(linear_2(relu(linear_1(s))))))   (layers.py:138-145)"""
        x = s.to(torch.float32).contiguous()
        return self._engine(x, x.numel() // self.dim).node_transition(self._s2s_block, x)


class EdgeTransition(_Native):
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

    def forward(self, node_embed, edge_embed):
        """[B,L,256], [B,L,L,128] -> [B,L,L,128] (layers.py:170-185; the pair tensor is bf16 inside the library, the result
        is returned in the dtype of `edge_embed`)."""
        from .ipa import _engine_of

        B, L = node_embed.shape[:2]
        eng = _engine_of(self).native(node_embed.device)
        eng.reserve(B, L)
        ones = torch.ones(B, L, device=node_embed.device, dtype=torch.float32)
        out = eng.edge_transition(self._s2s_block, node_embed.to(torch.float32).contiguous(), edge_embed.to(torch.bfloat16).contiguous(), ones)
        return out.to(edge_embed.dtype)


class TorsionAngleHead(_Native):
    def __init__(self, in_dim, n_torsion_angles, eps=1e-8):
        super().__init__()
        self.in_dim = in_dim
        self.linear_1 = Linear(in_dim, in_dim, init="relu")
        self.linear_2 = Linear(in_dim, in_dim, init="relu")
        self.linear_3 = Linear(in_dim, in_dim, init="final")  # registered, never used (reference quirk)
        self.linear_final = Linear(in_dim, n_torsion_angles * 2, init="final")
        self.eps = eps

    def forward(self, s):
        """unit-norm (clamp eps) linear_final(s + linear_2(relu(linear_1(s))))   (layers.py:199-213)"""
        x = s.to(torch.float32).contiguous()
        return self._engine(x, x.numel() // self.in_dim).torsion_head(x)


class BackboneUpdate(_Native):
    def __init__(self, c_s):
        super().__init__()
        self.c_s = c_s
        self.linear = Linear(c_s, 6, init="final")

    def forward(self, s: torch.Tensor):
        """[*, N_res, C_s] -> [*, N_res, 6]   (layers.py:232-241)"""
        x = s.to(torch.float32).contiguous()
        return self._engine(x, x.numel() // self.c_s).backbone_update(self._s2s_block, x)
