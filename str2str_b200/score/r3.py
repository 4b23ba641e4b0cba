"""Drop-in `R3Diffuser` (reference src/models/score/r3.py): VP-SDE schedule scalars on the host.

The per-residue arithmetic (score, reverse step, perturbation) runs in the fused SE(3) kernels
(csrc/rigid.cu); this class supplies the schedule values, computed with the reference's own torch
expressions so they are bit-identical.
"""
from __future__ import annotations

import torch


class R3Diffuser:
    def __init__(self, min_b: float = 0.1, max_b: float = 20.0, coordinate_scaling: float = 1.0):
        self.min_b, self.max_b, self.coordinate_scaling = min_b, max_b, coordinate_scaling

    def scale(self, x):
        return x * self.coordinate_scaling

    def unscale(self, x):
        return x / self.coordinate_scaling

    def b_t(self, t: torch.Tensor):
        if torch.any(t < 0) or torch.any(t > 1):
            raise ValueError(f"Invalid t={t}")
        return self.min_b + t * (self.max_b - self.min_b)

    def diffusion_coef(self, t):
        return torch.sqrt(self.b_t(t))

    def marginal_b_t(self, t):
        return t * self.min_b + 0.5 * (t ** 2) * (self.max_b - self.min_b)

    def conditional_var(self, t, use_torch=False):
        return 1.0 - torch.exp(-self.marginal_b_t(t))

    def score_scaling(self, t: torch.Tensor):
        return 1.0 / torch.sqrt(self.conditional_var(t))

    def sample_prior(self, shape, device=None):
        return torch.randn(size=shape, device=device)
