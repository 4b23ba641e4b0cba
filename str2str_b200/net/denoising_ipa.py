"""Drop-in `EmbeddingModule` / `DenoisingNet` (reference src/models/net/denoising_ipa.py:49-211).

`DenoisingNet.forward(batch, as_tensor_7)` keeps the reference's batch-dict keys and output dict; it owns the
native context (one per device) built from its own `state_dict()`, so a reference checkpoint loaded with
`load_state_dict` is what the kernels run.
"""
from __future__ import annotations

import weakref
from typing import Dict

import torch
import torch.nn as nn

from ..engine import NativeEngine
from ..rigid import Rigid

_EMB_SUPPORTED = dict(init_embed_size=32, node_embed_size=256, edge_embed_size=128, num_bins=22, min_bin=1e-5, max_bin=20.0)


class EmbeddingModule(nn.Module):
    def __init__(self, init_embed_size: int, node_embed_size: int, edge_embed_size: int, num_bins: int = 22,
                 min_bin: float = 1e-5, max_bin: float = 20.0, self_conditioning: bool = True):
        super().__init__()
        got = dict(init_embed_size=init_embed_size, node_embed_size=node_embed_size, edge_embed_size=edge_embed_size,
                   num_bins=num_bins, min_bin=min_bin, max_bin=max_bin)
        if got != _EMB_SUPPORTED or not self_conditioning:
            raise ValueError(f"the sm_100a embedder is specialised for {_EMB_SUPPORTED}, self_conditioning=True; got {got}")
        node_in = init_embed_size + 1 + init_embed_size
        edge_in = 2 * (init_embed_size + 1) + init_embed_size + num_bins
        self.self_conditioning = self_conditioning

        def mlp(d_in, d):
            return nn.Sequential(nn.Linear(d_in, d), nn.ReLU(), nn.Linear(d, d), nn.ReLU(), nn.Linear(d, d), nn.LayerNorm(d))

        self.node_embed = mlp(node_in, node_embed_size)
        self.edge_embed = mlp(edge_in, edge_embed_size)

    def forward(self, residue_idx, t, fixed_mask, self_conditioning_ca):
        from .ipa import _engine_of

        eng = _engine_of(self).native(self_conditioning_ca.device)
        B, L = residue_idx.shape
        residue_idx = residue_idx.long().contiguous()
        eng.reserve(B, L, residue_idx)
        f32 = lambda x: x.to(torch.float32).contiguous()
        ones = torch.ones(B, L, device=residue_idx.device, dtype=torch.float32)
        node, z = eng.embed(f32(t), residue_idx, f32(fixed_mask), f32(self_conditioning_ca), ones)
        return node, z.float()


class DenoisingNet(nn.Module):
    def __init__(self, embedder: nn.Module, translator: nn.Module, pair_kernels: int = 1, node_gemm: int = 1):
        """`pair_kernels` / `node_gemm` (not reference kwargs; both default to the tensor-core path that bench.py times):
        0 selects the SIMT pair kernels / the exact-fp32 FFMA node GEMMs that the tests use as on-device cross-checks."""
        super().__init__()
        self.embedder = embedder
        self.translator = translator
        self._native: Dict[str, NativeEngine] = {}
        self._opts = dict(pair_kernels=pair_kernels, node_gemm=node_gemm)
        ref = weakref.ref(self)
        for m in self.modules():
            object.__setattr__(m, "_s2s_root", ref)

    # -- native context management ---------------------------------------------------------------------
    def native(self, device) -> NativeEngine:
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("str2str_b200.DenoisingNet runs on CUDA only: move the module and the batch to a GPU")
        if device.index is None:  # "cuda" and "cuda:<current>" are one device: one engine (and one set of options) for both spellings
            device = torch.device("cuda", torch.cuda.current_device())
        key = str(device)
        if key not in self._native:
            self._native[key] = NativeEngine(self.state_dict(), device, **self._opts)
        return self._native[key]

    def set_option(self, key: str, value: int):
        self._opts[key] = value
        for e in self._native.values():
            e.set_option(key, value)

    def _invalidate(self):
        # the engines (and with them every workspace / weight image a captured CUDA graph may point into) are dropped; samplers
        # notice through NativeEngine.token, which is unique per engine instance
        self._native = {}

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._invalidate()
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._invalidate()
        return out

    # -- forward ---------------------------------------------------------------------------------------
    def forward(self, batch, as_tensor_7=False):
        dev = batch["rigids_t"].device
        eng = self.native(dev)
        ridx = batch["residue_idx"].contiguous()
        B, L = ridx.shape
        if ridx.dtype != torch.int64:
            ridx = ridx.long()
        # always re-plan from the actual indices (a min/max of [B, L] integers): other callers share this engine's
        # relative-position table, and a tensor's address says nothing about its contents
        eng.reserve(B, L, ridx)
        f32 = lambda x: x.to(torch.float32).contiguous()
        fixed = f32(batch["fixed_mask"])
        gt_psi = f32(batch["torsion_angles_sin_cos"][..., 2, :])
        out7, psi = eng.net_forward(f32(batch["rigids_t"]), f32(batch["sc_ca_t"]), f32(batch["t"]), ridx,
                                    f32(batch["residue_mask"]), fixed, gt_psi)
        aatype = batch["aatype"].long().contiguous() if "aatype" in batch else None
        atom37, atom14 = eng.backbone_atoms(out7, psi, aatype)
        rigids = out7 if as_tensor_7 else Rigid.from_tensor_7(out7)
        return {"rigids": rigids, "psi": psi, "atom37": atom37, "atom14": atom14}
