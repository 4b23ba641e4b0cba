"""Thin owner of one native context (s2s_ctx): weights, workspace, stage calls.

Everything here is plumbing — tensors in, device pointers to the C ABI, tensors out.  No arithmetic.
"""
from __future__ import annotations

import itertools
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


_ENGINE_UIDS = itertools.count(1)


class NativeEngine:
    """One s2s_ctx bound to one CUDA device."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device: torch.device, pair_kernels: int = 1, node_gemm: int = 1, **options: int):
        self.uid = next(_ENGINE_UIDS)  # process-wide unique: (uid, generation) identifies one workspace allocation
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
            _lib.check(self.lib.s2s_finalize(self.ctx, _lib.stream(self.device)))
            self.set_option("pair_kernels", pair_kernels)
            self.set_option("node_gemm", node_gemm)
            for key, value in options.items():  # further s2s_set_option keys (et_pair, embed_table, chain, ...)
                self.set_option(key, value)
        self.shape = None
        self.generation = 0  # bumped whenever the workspace is (re)allocated: invalidates captured CUDA graphs

    @property
    def token(self):
        """Identity of the current workspace allocation.  Captured CUDA graphs hold raw pointers into it (and into this
        engine's weight images), so a graph is valid only while the token it was captured under is still current; the
        uid part makes tokens of different engines (e.g. after load_state_dict rebuilt the context) never compare equal."""
        return (self.uid, self.generation)

    def __del__(self):
        ctx, self.ctx = getattr(self, "ctx", None), None
        if ctx:
            self.lib.s2s_destroy(ctx)

    def set_option(self, key: str, value: int):
        _lib.check(self.lib.s2s_set_option(self.ctx, key.encode(), int(value)))

    def reserve(self, B: int, L: int, residue_idx: Optional[torch.Tensor] = None):
        """Size the workspace for (B, L) and plan the relative-position table from the min/max residue-index offset (one host
        sync, setup time only).  `residue_idx=None` is a shape-only call: the table already planned stays (callers that run a
        sub-module without residue indices, e.g. InvariantPointAttention.forward).  Any L is accepted (the library pads
        chain lengths to its tile size internally).  `generation` is bumped whenever the allocation or the table changes:
        both are baked into captured CUDA graphs."""
        if residue_idx is None:
            span = None if self.shape is None else self.shape[2]
        else:
            span = int(residue_idx.max()) - int(residue_idx.min())
        if self.shape is not None and self.shape[0] >= B and self.shape[1] >= L and self.shape[2] == span:
            return
        with torch.cuda.device(self.device):
            if span is None:
                _lib.check(self.lib.s2s_reserve(self.ctx, B, L, 1, 0, _lib.stream(self.device)))  # d_max < d_min: keep / default table
            else:
                _lib.check(self.lib.s2s_reserve(self.ctx, B, L, -span, span, _lib.stream(self.device)))
        old = self.shape or (0, 0, None)
        self.shape = (max(B, old[0]), max(L, old[1]), span)
        self.generation += 1

    # -- stage calls; all inputs fp32/int64 contiguous CUDA tensors -------------------------------------
    def net_forward(self, rigids_t, sc_ca, t, residue_idx, residue_mask, fixed_mask, gt_psi, out_rigids=None, out_psi=None):
        B, L = residue_idx.shape
        if out_rigids is None:
            out_rigids = torch.empty(B, L, 7, device=self.device, dtype=torch.float32)
        if out_psi is None:
            out_psi = torch.empty(B, L, 2, device=self.device, dtype=torch.float32)
        p, pi = _lib.ptr, _lib.ptr_i64
        _lib.check(self.lib.s2s_net_forward(self.ctx, B, L, p(rigids_t), p(sc_ca), p(t), pi(residue_idx), p(residue_mask),
                                            p(fixed_mask), p(gt_psi), p(out_rigids), p(out_psi), _lib.stream(self.device)))
        return out_rigids, out_psi

    def trunk(self, node, z_bf16, rigids_t, residue_mask, fixed_mask, gt_psi: Optional[torch.Tensor]):
        B, L = residue_mask.shape
        out_rigids = torch.empty(B, L, 7, device=self.device, dtype=torch.float32)
        out_psi = torch.empty(B, L, 2, device=self.device, dtype=torch.float32)
        p, pb = _lib.ptr, _lib.ptr_bf16
        _lib.check(self.lib.s2s_trunk(self.ctx, B, L, p(node), pb(z_bf16), p(rigids_t), p(residue_mask), p(fixed_mask),
                                      p(gt_psi), p(out_rigids), p(out_psi), _lib.stream(self.device)))
        return out_rigids, out_psi

    def embed(self, t, residue_idx, fixed_mask, sc_ca, residue_mask):
        B, L = residue_idx.shape
        node = torch.empty(B, L, 256, device=self.device, dtype=torch.float32)
        z = torch.empty(B, L, L, 128, device=self.device, dtype=torch.bfloat16)
        p, pi, pb = _lib.ptr, _lib.ptr_i64, _lib.ptr_bf16
        _lib.check(self.lib.s2s_embed(self.ctx, B, L, p(t), pi(residue_idx), p(fixed_mask), p(sc_ca), p(residue_mask),
                                      p(node), pb(z), _lib.stream(self.device)))
        return node, z

    def ipa(self, blk, node, z_bf16, quat, trans_nm, residue_mask):
        B, L = residue_mask.shape
        out = torch.empty(B, L, 256, device=self.device, dtype=torch.float32)
        p, pb = _lib.ptr, _lib.ptr_bf16
        _lib.check(self.lib.s2s_ipa(self.ctx, blk, B, L, p(node), pb(z_bf16), p(quat), p(trans_nm), p(residue_mask), p(out),
                                    _lib.stream(self.device)))
        return out

    def edge_transition(self, blk, node, z_bf16, residue_mask):
        B, L = residue_mask.shape
        out = torch.empty_like(z_bf16)
        p, pb = _lib.ptr, _lib.ptr_bf16
        _lib.check(self.lib.s2s_edge_transition(self.ctx, blk, B, L, p(node), pb(z_bf16), p(residue_mask), pb(out), _lib.stream(self.device)))
        return out

    def node_transition(self, blk, s):
        """NodeTransition.forward of block `blk` on [..., 256] rows."""
        out = torch.empty_like(s)
        _lib.check(self.lib.s2s_node_transition(self.ctx, blk, s.numel() // 256, _lib.ptr(s), _lib.ptr(out), _lib.stream(self.device)))
        return out

    def torsion_head(self, s):
        out = torch.empty(*s.shape[:-1], 2, device=self.device, dtype=torch.float32)
        _lib.check(self.lib.s2s_torsion_head(self.ctx, s.numel() // 256, _lib.ptr(s), _lib.ptr(out), _lib.stream(self.device)))
        return out

    def backbone_update(self, blk, s):
        out = torch.empty(*s.shape[:-1], 6, device=self.device, dtype=torch.float32)
        _lib.check(self.lib.s2s_backbone_update(self.ctx, blk, s.numel() // 256, _lib.ptr(s), _lib.ptr(out), _lib.stream(self.device)))
        return out

    def backbone_atoms(self, rigids7, psi, aatype, want_atom14=True):
        lead = rigids7.shape[:-1]
        rows = int(np.prod(lead))
        atom37 = torch.empty(*lead, 37, 3, device=self.device, dtype=torch.float32)
        atom14 = torch.empty(*lead, 14, 3, device=self.device, dtype=torch.float32) if want_atom14 else None
        p = _lib.ptr
        _lib.check(self.lib.s2s_backbone_atoms(self.ctx, rows, p(rigids7), p(psi), _lib.ptr_i64(aatype), p(atom37), p(atom14),
                                               _lib.stream(self.device)))
        return atom37, atom14
