"""Drop-in `InvariantPointAttention` / `TranslationIPA` (reference src/models/net/ipa.py:31-387).

Same constructor kwargs, parameter names and forward signatures; the arithmetic is the native library's.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.nn as nn

from ..rigid import Rigid
from .layers import BackboneUpdate, EdgeTransition, Linear, NodeTransition, TorsionAngleHead

_SUPPORTED = dict(c_s=256, c_z=128, c_hidden=256, no_heads=8, no_qk_points=8, no_v_points=12)


def _engine_of(module):
    root = getattr(module, "_s2s_root", None)
    root = root() if root is not None else None
    if root is None:
        raise RuntimeError(
            f"{type(module).__name__} must be attached to a str2str_b200 DenoisingNet (which owns the native "
            "context with all 274 weight tensors) before it can run"
        )
    return root


class InvariantPointAttention(nn.Module):
    def __init__(self, c_s: int, c_z: int, c_hidden: int, no_heads: int, no_qk_points: int, no_v_points: int,
                 inf: float = 1e5, eps: float = 1e-8):
        super().__init__()
        got = dict(c_s=c_s, c_z=c_z, c_hidden=c_hidden, no_heads=no_heads, no_qk_points=no_qk_points, no_v_points=no_v_points)
        if got != _SUPPORTED or inf != 1e5 or eps != 1e-8:
            raise ValueError(f"the sm_100a kernels are specialised for {_SUPPORTED}, inf=1e5, eps=1e-8; got {got}")
        self.c_s, self.c_z, self.c_hidden, self.no_heads = c_s, c_z, c_hidden, no_heads
        self.no_qk_points, self.no_v_points, self.inf, self.eps = no_qk_points, no_v_points, inf, eps
        hc = c_hidden * no_heads
        self.linear_q = Linear(c_s, hc)
        self.linear_kv = Linear(c_s, 2 * hc)
        self.linear_q_points = Linear(c_s, no_heads * no_qk_points * 3)
        self.linear_kv_points = Linear(c_s, no_heads * (no_qk_points + no_v_points) * 3)
        self.linear_b = Linear(c_z, no_heads)
        self.down_z = Linear(c_z, c_z // 4)
        self.head_weights = nn.Parameter(torch.full((no_heads,), 0.541324854612918))
        self.linear_out = Linear(no_heads * (c_z // 4 + c_hidden + no_v_points * 4), c_s, init="final")
        self._s2s_block = None

    def forward(self, s: torch.Tensor, z: Optional[torch.Tensor], r: Rigid, mask: torch.Tensor,
                _offload_inference: bool = False, _z_reference_list: Optional[Sequence[torch.Tensor]] = None) -> torch.Tensor:
        if _offload_inference:
            z = _z_reference_list[0]
        net = _engine_of(self)
        eng = net.native(s.device)
        B, L = mask.shape
        eng.reserve(B, L)  # shape only: the relative-position table of the embedder is not this module's to re-plan
        f32 = lambda x: x.to(torch.float32).contiguous()
        return eng.ipa(self._s2s_block, f32(s), z.to(torch.bfloat16).contiguous(), f32(r.get_rots().get_quats()),
                       f32(r.get_trans()), f32(mask))


class TranslationIPA(nn.Module):
    def __init__(self, c_s: int, c_z: int, coordinate_scaling: float, no_ipa_blocks: int, skip_embed_size: int,
                 transformer_num_heads: int = 4, transformer_num_layers: int = 2, c_hidden: int = 256, no_heads: int = 8,
                 no_qk_points: int = 8, no_v_points: int = 12, dropout: float = 0.0):
        super().__init__()
        if (no_ipa_blocks, skip_embed_size, transformer_num_heads, transformer_num_layers) != (4, 64, 4, 2) or coordinate_scaling != 0.1:
            raise ValueError("the sm_100a trunk is specialised for 4 blocks, skip 64, 4x2 transformer, coordinate_scaling 0.1")
        self.coordinate_scaling = coordinate_scaling
        self.num_blocks = no_ipa_blocks
        self.trunk = nn.ModuleDict()
        for b in range(no_ipa_blocks):
            ipa = InvariantPointAttention(c_s=c_s, c_z=c_z, c_hidden=c_hidden, no_heads=no_heads,
                                          no_qk_points=no_qk_points, no_v_points=no_v_points)
            ipa._s2s_block = b
            self.trunk[f"ipa_{b}"] = ipa
            self.trunk[f"ipa_ln_{b}"] = nn.LayerNorm(c_s)
            self.trunk[f"skip_embed_{b}"] = Linear(c_s, skip_embed_size, init="final")
            d = c_s + skip_embed_size
            layer = nn.TransformerEncoderLayer(d_model=d, nhead=transformer_num_heads, dim_feedforward=d)
            self.trunk[f"transformer_{b}"] = nn.TransformerEncoder(layer, transformer_num_layers, enable_nested_tensor=False)
            self.trunk[f"linear_{b}"] = Linear(d, c_s, init="final")
            self.trunk[f"node_transition_{b}"] = NodeTransition(c_s)
            self.trunk[f"bb_update_{b}"] = BackboneUpdate(c_s)
            if b < no_ipa_blocks - 1:
                self.trunk[f"edge_transition_{b}"] = EdgeTransition(node_embed_size=c_s, edge_embed_in=c_z, edge_embed_out=c_z)
                self.trunk[f"edge_transition_{b}"]._s2s_block = b
            self.trunk[f"node_transition_{b}"]._s2s_block = b
            self.trunk[f"bb_update_{b}"]._s2s_block = b
        self.torsion_pred = TorsionAngleHead(c_s, 1)

    def forward(self, node_embed, edge_embed, batch):
        net = _engine_of(self)
        eng = net.native(node_embed.device)
        f32 = lambda x: x.to(torch.float32).contiguous()
        node_mask, fixed = f32(batch["residue_mask"]), f32(batch["fixed_mask"])
        B, L = node_mask.shape
        eng.reserve(B, L)  # shape only (the trunk reads no residue indices)
        init = f32(batch["rigids_t"])
        out7, psi = eng.trunk(f32(node_embed), edge_embed.to(torch.bfloat16).contiguous(), init, node_mask, fixed, None)
        return {"in_rigids": Rigid.from_tensor_7(init), "out_rigids": Rigid.from_tensor_7(out7), "psi": psi}
