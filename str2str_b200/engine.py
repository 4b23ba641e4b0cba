"""Thin owner of one native context (s2s_ctx): weights, workspace, stage calls.

Everything here is plumbing — tensors in, device pointers to the C ABI, tensors out.  No arithmetic.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from .backbone_constants import BACKBONE_MASK, BACKBONE_POS, BB_FRAME_VALID, PSI_FRAME


def host_tables():
    """Constant tables computed with the reference's own torch expressions so they are bit-identical
    (denoising_ipa.py:26-30,39-41; geo_utils.py:49-53)."""
    tfreq = torch.exp(torch.arange(16, dtype=torch.float) * -(math.log(10000) / (16 - 1)))
    pdenom = 2056 ** (2 * torch.arange(16)[None] / 32)
    bin_lower = torch.linspace(1e-5, 20.0, 22)
    pos = np.asarray(BACKBONE_POS, dtype=np.float32).reshape(21, 15)
    mask = np.asarray(BACKBONE_MASK, dtype=np.float32).reshape(21, 5)
    psi = np.asarray(PSI_FRAME, dtype=np.float32)
    table = np.concatenate(
        [pos, mask, psi[:, :3, :3].reshape(21, 9), psi[:, :3, 3], np.asarray(BB_FRAME_VALID, dtype=np.float32)[:, None]], axis=1
    )
    assert table.shape == (21, 33)
    return (
        tfreq.float().contiguous(),
        pdenom.float().reshape(16).contiguous(),
        bin_lower.float().contiguous(),
        torch.from_numpy(np.ascontiguousarray(table)),
    )


class NativeEngine:
    """One s2s_ctx bound to one CUDA device."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device: torch.device, pair_kernels: int = 1, node_gemm: int = 0):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("str2str_b200 runs on CUDA devices only (no CPU fallback)")
        tabs = host_tables()
        with torch.cuda.device(self.device):
            self.ctx = self.lib.s2s_create(*[t.data_ptr() for t in tabs])
            if not self.ctx:
                raise RuntimeError("s2s_create failed: " + self.lib.s2s_last_error().decode())
            # keep fp32 contiguous device copies alive for the lifetime of the context
            self.params = {k: v.detach().to(self.device, torch.float32).contiguous() for k, v in state_dict.items()}
            for k, v in self.params.items():
                _lib.check(self.lib.s2s_set_param(self.ctx, k.encode(), v.data_ptr(), v.numel()))
            _lib.check(self.lib.s2s_finalize(self.ctx, _lib.stream()))
            self.set_option("pair_kernels", pair_kernels)
            self.set_option("node_gemm", node_gemm)
        self.shape = None
        self.generation = 0  # bumped whenever the workspace is (re)allocated: invalidates captured CUDA graphs

    def __del__(self):
        ctx, self.ctx = getattr(self, "ctx", None), None
        if ctx:
            self.lib.s2s_destroy(ctx)

    def set_option(self, key: str, value: int):
        _lib.check(self.lib.s2s_set_option(self.ctx, key.encode(), int(value)))

    def reserve(self, B: int, L: int, residue_idx: torch.Tensor):
        """Size the workspace; needs min/max residue-index offset (one host sync, setup time only)."""
        lo, hi = int(residue_idx.min()), int(residue_idx.max())
        key = (B, L, lo - hi, hi - lo)
        if self.shape is not None and self.shape[0] >= B and self.shape[1] >= L and self.shape[2:] == key[2:]:
            return
        with torch.cuda.device(self.device):
            _lib.check(self.lib.s2s_reserve(self.ctx, B, L, lo - hi, hi - lo, _lib.stream()))
        self.shape = key
        self.generation += 1

    # -- stage calls; all inputs fp32/int64 contiguous CUDA tensors -------------------------------------
    def net_forward(self, rigids_t, sc_ca, t, residue_idx, residue_mask, fixed_mask, gt_psi, out_rigids=None, out_psi=None):
        B, L = residue_idx.shape
        if out_rigids is None:
            out_rigids = torch.empty(B, L, 7, device=self.device, dtype=torch.float32)
        if out_psi is None:
            out_psi = torch.empty(B, L, 2, device=self.device, dtype=torch.float32)
        p = _lib.ptr
        _lib.check(self.lib.s2s_net_forward(self.ctx, B, L, p(rigids_t), p(sc_ca), p(t), p(residue_idx), p(residue_mask),
                                            p(fixed_mask), p(gt_psi), p(out_rigids), p(out_psi), _lib.stream()))
        return out_rigids, out_psi

    def trunk(self, node, z_bf16, rigids_t, residue_mask, fixed_mask, gt_psi: Optional[torch.Tensor]):
        B, L = residue_mask.shape
        out_rigids = torch.empty(B, L, 7, device=self.device, dtype=torch.float32)
        out_psi = torch.empty(B, L, 2, device=self.device, dtype=torch.float32)
        p = _lib.ptr
        _lib.check(self.lib.s2s_trunk(self.ctx, B, L, p(node), p(z_bf16), p(rigids_t), p(residue_mask), p(fixed_mask),
                                      p(gt_psi), p(out_rigids), p(out_psi), _lib.stream()))
        return out_rigids, out_psi

    def embed(self, t, residue_idx, fixed_mask, sc_ca, residue_mask):
        B, L = residue_idx.shape
        node = torch.empty(B, L, 256, device=self.device, dtype=torch.float32)
        z = torch.empty(B, L, L, 128, device=self.device, dtype=torch.bfloat16)
        p = _lib.ptr
        _lib.check(self.lib.s2s_embed(self.ctx, B, L, p(t), p(residue_idx), p(fixed_mask), p(sc_ca), p(residue_mask),
                                      p(node), p(z), _lib.stream()))
        return node, z

    def ipa(self, blk, node, z_bf16, quat, trans_nm, residue_mask):
        B, L = residue_mask.shape
        out = torch.empty(B, L, 256, device=self.device, dtype=torch.float32)
        p = _lib.ptr
        _lib.check(self.lib.s2s_ipa(self.ctx, blk, B, L, p(node), p(z_bf16), p(quat), p(trans_nm), p(residue_mask), p(out),
                                    _lib.stream()))
        return out

    def edge_transition(self, blk, node, z_bf16, residue_mask):
        B, L = residue_mask.shape
        out = torch.empty_like(z_bf16)
        p = _lib.ptr
        _lib.check(self.lib.s2s_edge_transition(self.ctx, blk, B, L, p(node), p(z_bf16), p(residue_mask), p(out), _lib.stream()))
        return out

    def backbone_atoms(self, rigids7, psi, aatype, want_atom14=True):
        lead = rigids7.shape[:-1]
        rows = int(np.prod(lead))
        atom37 = torch.empty(*lead, 37, 3, device=self.device, dtype=torch.float32)
        atom14 = torch.empty(*lead, 14, 3, device=self.device, dtype=torch.float32) if want_atom14 else None
        p = _lib.ptr
        _lib.check(self.lib.s2s_backbone_atoms(self.ctx, rows, p(rigids7), p(psi), p(aatype), p(atom37), p(atom14), _lib.stream()))
        return atom37, atom14
