"""Drop-in `FrameDiffuser` (reference src/models/score/frame.py:21-255) on the fused SE(3) kernels.

Same method signatures and return types (`Rigid` objects / tensor_7 / fp64 score tensors).  Schedule scalars
are evaluated on the host with the reference's torch expressions (a [B]-sized computation), everything
per-residue runs in csrc/rigid.cu through the C ABI.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from .. import _lib
from ..rigid import Rigid
from . import r3, so3


def schedule_rows(trans_diffuser: r3.R3Diffuser, rot_diffuser: so3.SO3Diffuser, t_cpu: torch.Tensor):
    """Per-decoy scalars for s2s_se3_step: [B,8] fp32 (see include/str2str_b200.h) and the sigma buckets."""
    t = t_cpu.detach().to("cpu", torch.float32)
    sig = rot_diffuser.sigma(t)
    idx = rot_diffuser.sigma_idx(sig)
    sigma_q = rot_diffuser.discrete_sigma[idx]
    g_rot = rot_diffuser.diffusion_coef(t)
    beta = trans_diffuser.marginal_b_t(t)
    b_t = trans_diffuser.b_t(t)
    rows = torch.stack([t, sigma_q, g_rot, g_rot ** 2, torch.exp(-0.5 * beta), 1.0 - torch.exp(-beta), b_t, torch.sqrt(b_t)], -1)
    return rows.float().contiguous(), idx


class FrameDiffuser:
    def __init__(self, trans_diffuser: Optional[r3.R3Diffuser] = None, rot_diffuser: Optional[so3.SO3Diffuser] = None,
                 min_t: float = 0.001):
        if trans_diffuser is None or rot_diffuser is None:
            raise ValueError("the fused SE(3) kernels diffuse rotations and translations together; both diffusers are required")
        if (trans_diffuser.min_b, trans_diffuser.max_b, trans_diffuser.coordinate_scaling) != (0.1, 20.0, 0.1):
            raise ValueError("R3Diffuser must use min_b=0.1, max_b=20.0, coordinate_scaling=0.1 (configs/model/diffusion.yaml:45-48)")
        self.trans_diffuser, self.rot_diffuser, self.min_t = trans_diffuser, rot_diffuser, min_t

    # ------------------------------------------------------------------------------------------------
    def _sched(self, t: torch.Tensor, dt: float, device):
        rows, idx = schedule_rows(self.trans_diffuser, self.rot_diffuser, t)
        B = rows.shape[0]
        sd = torch.tensor([[dt, np.sqrt(dt)]] * B, dtype=torch.float64)
        return rows.to(device), sd.to(device), idx

    @staticmethod
    def _tensor7(r) -> torch.Tensor:
        t7 = r.to_tensor_7() if isinstance(r, Rigid) else r
        return t7.to(torch.float32).contiguous()

    def score(self, rigids_0: Rigid, rigids_t: Rigid, t: torch.Tensor, mask: torch.Tensor = None):
        r0, rt = self._tensor7(rigids_0), self._tensor7(rigids_t)
        B, L = r0.shape[:2]
        dev = r0.device
        sf, sd, _ = self._sched(t, 1.0, dev)
        m = torch.ones(B, L, device=dev) if mask is None else mask.to(dev, torch.float32).contiguous()
        rs = torch.empty(B, L, 3, device=dev, dtype=torch.float64)
        ts = torch.empty(B, L, 3, device=dev, dtype=torch.float64)
        p, pd = _lib.ptr, _lib.ptr_f64
        _lib.check(_lib.load().s2s_se3_step(B, L, p(rt), p(r0), p(m), None, p(sf), pd(sd), None, None, C.c_float(1.0), 1, 1,
                                            pd(rs), pd(ts), None, _lib.stream(dev)))
        if mask is None or mask.dtype != torch.float64:
            rs, ts = rs.float(), ts.float()  # without fp64 masks the reference's scores stay fp32
        return {"trans_score": ts, "rot_score": rs}

    def reverse(self, rigids_t: Rigid, rot_score: torch.Tensor, trans_score: torch.Tensor, t: torch.Tensor, dt: float,
                diffuse_mask: torch.Tensor = None, center_trans: bool = True, noise_scale: float = 1.0,
                probability_flow: bool = True, rot_noise: torch.Tensor = None, trans_noise: torch.Tensor = None):
        """`rot_noise` / `trans_noise` ([B,L,3] N(0,1)) are optional injected draws for SDE mode; if absent they
        are drawn with torch.randn on the device (the reference draws them inside so3/r3.reverse)."""
        if not center_trans:
            raise ValueError("center_trans=False is never used by the reference sampler and has no kernel")
        rt = self._tensor7(rigids_t)
        B, L = rt.shape[:2]
        dev = rt.device
        sf, sd, _ = self._sched(t, dt, dev)
        rs = rot_score.to(dev, torch.float64).contiguous()
        ts = trans_score.to(dev, torch.float64).contiguous()
        dm = None if diffuse_mask is None else diffuse_mask.to(dev, torch.float32).contiguous()
        if not probability_flow:
            rot_noise = torch.randn(B, L, 3, device=dev) if rot_noise is None else rot_noise.to(dev, torch.float32).contiguous()
            trans_noise = torch.randn(B, L, 3, device=dev) if trans_noise is None else trans_noise.to(dev, torch.float32).contiguous()
        ones = torch.ones(B, L, device=dev)
        out = torch.empty(B, L, 7, device=dev, dtype=torch.float32)
        p, pd = _lib.ptr, _lib.ptr_f64
        _lib.check(_lib.load().s2s_se3_step(B, L, p(rt), None, p(ones), p(dm), p(sf), pd(sd), p(rot_noise), p(trans_noise),
                                            C.c_float(noise_scale), int(probability_flow), 2, pd(rs), pd(ts), p(out), _lib.stream(dev)))
        return Rigid.from_tensor_7(out)

    def score_and_reverse(self, rigids_0_7, rigids_t_7, residue_mask, diffuse_mask, sched_f, sched_d, out,
                          noise_scale=1.0, probability_flow=True, rot_noise=None, trans_noise=None):
        """Fused fast path used by the sampler: one launch, scores never touch HBM."""
        B, L = rigids_t_7.shape[:2]
        p = _lib.ptr
        _lib.check(_lib.load().s2s_se3_step(B, L, p(rigids_t_7), p(rigids_0_7), p(residue_mask), p(diffuse_mask), p(sched_f),
                                            _lib.ptr_f64(sched_d), p(rot_noise), p(trans_noise), C.c_float(noise_scale),
                                            int(probability_flow), 0, None, None, p(out), _lib.stream(rigids_t_7.device)))
        return out

    @staticmethod
    def decoy_noise(shape, device, seed: int, first_decoy: int, stream_id: int, uniform: bool = False) -> torch.Tensor:
        """[B, ...] draws keyed by global decoy id (s2s_rng_fill): decoy b of this call reads Philox subsequence first_decoy + b
        of `seed`, so its values do not depend on the batch / rank / world size it is sampled in."""
        out = torch.empty(*shape, device=device, dtype=torch.float32)
        B = shape[0]
        _lib.check(_lib.load().s2s_rng_fill(_lib.ptr(out), B, out.numel() // B, int(seed) & (2 ** 64 - 1), int(first_decoy),
                                            int(stream_id), int(uniform), _lib.stream(out.device)))
        return out

    def forward_marginal(self, rigids_0: Rigid, t: torch.Tensor, diffuse_mask: torch.Tensor = None, as_tensor_7: bool = True,
                         noise=None, seed: Optional[int] = None, first_decoy: int = 0):
        """`noise` = (axis [B,L,3] N(0,1), u [B,L] U[0,1), trans [B,L,3] N(0,1)) optionally injected.  With `seed` the three
        draws are keyed by global decoy id (`decoy_noise`); otherwise they come from torch's device generator in the
        reference's order (so3.py:259,262; r3.py:66).  Scores of the perturbation, which the sampler discards
        (diffusion_module.py:274-279), are returned as None."""
        rot0 = rigids_0.get_rots().get_rot_mats().to(torch.float32).contiguous()
        x0 = rigids_0.get_trans().to(torch.float32).contiguous()
        B, L = x0.shape[:2]
        dev = x0.device
        tc = t.detach().to("cpu", torch.float32)
        beta = self.trans_diffuser.marginal_b_t(tc)
        sf = torch.stack([torch.exp(-0.5 * beta), torch.sqrt(1 - torch.exp(-beta))], -1).float().contiguous().to(dev)
        idx = self.rot_diffuser.t_to_idx(tc)
        cdf = torch.from_numpy(np.stack([self.rot_diffuser.cdf_row(int(i)) for i in idx])).to(dev).contiguous()
        omega = self.rot_diffuser.discrete_omega.float().contiguous().to(dev)
        if noise is None:
            if seed is not None:
                noise = (self.decoy_noise((B, L, 3), dev, seed, first_decoy, 0), self.decoy_noise((B, L), dev, seed, first_decoy, 1, True),
                         self.decoy_noise((B, L, 3), dev, seed, first_decoy, 2))
            else:
                noise = (torch.randn(B, L, 3, device=dev), torch.rand(B, L, device=dev), torch.randn(B, L, 3, device=dev))
        ax, u, zt = [n.to(dev, torch.float32).contiguous() for n in noise]
        dm = None if diffuse_mask is None else torch.as_tensor(diffuse_mask).to(dev, torch.float32).contiguous()
        out = torch.empty(B, L, 7, device=dev, dtype=torch.float32)
        p = _lib.ptr
        _lib.check(_lib.load().s2s_se3_perturb(B, L, p(rot0), p(x0), p(dm), p(sf), _lib.ptr_f64(cdf), p(omega), p(ax), p(u), p(zt), p(out),
                                               _lib.stream(dev)))
        rigids_t = out if as_tensor_7 else Rigid.from_tensor_7(out)
        return {"rigids_t": rigids_t, "trans_score": None, "rot_score": None,
                "trans_score_scaling": self.trans_diffuser.score_scaling(t), "rot_score_scaling": None}

    def sample_prior(self, shape, device, reference_rigids: Rigid = None, diffuse_mask: torch.Tensor = None,
                     as_tensor_7: bool = False, noise=None, seed: Optional[int] = None, first_decoy: int = 0):
        """frame.py:212-255 without reference rigids (`backward_only: true`): rotation ~ IGSO3(t = 1) about a uniform axis
        (so3.py:244-272 via sample_ref at t = 1: the identity composed with the sampled rotation vector), translation ~
        N(0, 1) / coordinate_scaling (r3.py:76-77).  `noise` = (axis, u, trans) as in forward_marginal."""
        if reference_rigids is not None or diffuse_mask is not None:
            raise ValueError("sample_prior with reference_rigids is a motif-scaffolding path the sampler never takes")
        B, L = shape
        zero = Rigid.from_tensor_4x4(torch.eye(4, device=device).expand(B, L, 4, 4))
        if noise is None and seed is None:
            noise = (torch.randn(B, L, 3, device=device), torch.rand(B, L, device=device), torch.randn(B, L, 3, device=device))
        elif noise is None:
            noise = (self.decoy_noise((B, L, 3), device, seed, first_decoy, 0), self.decoy_noise((B, L), device, seed, first_decoy, 1, True),
                     self.decoy_noise((B, L, 3), device, seed, first_decoy, 2))
        # identity frames perturbed at t = 1 give the rotation; the translation is then replaced by the prior draw itself
        out = self.forward_marginal(zero, torch.ones(B), None, as_tensor_7=True, noise=noise)["rigids_t"]
        out[..., 4:] = noise[2].to(device, torch.float32) / self.trans_diffuser.coordinate_scaling
        return {"rigids_t": out if as_tensor_7 else Rigid.from_tensor_7(out)}
