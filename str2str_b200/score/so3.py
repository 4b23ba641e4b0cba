"""Drop-in `SO3Diffuser` (reference src/models/score/so3.py:133-371): IGSO(3) schedule and tables on the host.

sigma(t), its 1000-bucket quantisation (np.digitize, integer — bit exact by construction: same ops as the
reference) and the CDF rows used for sampling are host-side; the series, scores and compositions run in
csrc/rigid.cu.  CDF rows are built lazily per sigma bucket with the reference's fp64 numpy formula
(so3.py:21-62,171-183) instead of the full 1000x1000 table, and cached in `cache_dir` like the reference does.
"""
from __future__ import annotations

import math
import os
from typing import Dict

import numpy as np
import torch


class SO3Diffuser:
    def __init__(self, cache_dir: str = "./cache", schedule: str = "logarithmic", min_sigma: float = 0.1,
                 max_sigma: float = 1.5, num_sigma: int = 1000, num_omega: int = 1000, use_cached_score: bool = False,
                 eps: float = 1e-6):
        if schedule != "logarithmic":
            raise ValueError(f"Unrecognize schedule {schedule}")
        if use_cached_score:
            raise ValueError("use_cached_score=True is a training-time lookup the inference path never takes (diffusion.yaml:57)")
        if num_omega != 1000 or eps != 1e-6:
            raise ValueError("the SE(3) kernels are specialised for num_omega=1000, eps=1e-6")
        self.schedule, self.min_sigma, self.max_sigma, self.num_sigma = schedule, min_sigma, max_sigma, num_sigma
        self.num_omega, self.use_cached_score, self.eps = num_omega, use_cached_score, eps
        self.discrete_omega = torch.linspace(0, np.pi, steps=num_omega + 1)[1:]
        rp = lambda x: str(x).replace(".", "_")
        self.cache_dir = os.path.join(
            cache_dir, f"eps_{num_sigma}_omega_{num_omega}_min_sigma_{rp(min_sigma)}_max_sigma_{rp(max_sigma)}_schedule_{schedule}")
        self._cdf_rows: Dict[int, np.ndarray] = {}

    @property
    def discrete_sigma(self):
        return self.sigma(torch.linspace(0.0, 1.0, self.num_sigma))

    def sigma(self, t: torch.Tensor):
        if torch.any(t < 0) or torch.any(t > 1):
            raise ValueError(f"Invalid t={t}")
        return torch.log(t * math.exp(self.max_sigma) + (1 - t) * math.exp(self.min_sigma))

    def sigma_idx(self, sigma: torch.Tensor):
        return torch.as_tensor(np.digitize(sigma.cpu().numpy(), self.discrete_sigma) - 1, dtype=torch.long)

    def t_to_idx(self, t: torch.Tensor):
        return self.sigma_idx(self.sigma(t))

    def diffusion_coef(self, t: torch.Tensor):
        return torch.sqrt(2 * (math.exp(self.max_sigma) - math.exp(self.min_sigma)) * self.sigma(t) / torch.exp(self.sigma(t)))

    def cdf_row(self, idx: int) -> np.ndarray:
        """Row `idx` of the reference's `_cdf` table (fp64 [num_omega])."""
        if idx not in self._cdf_rows:
            path = os.path.join(self.cache_dir, f"cdf_row_{idx}.npy")
            row = None
            if os.path.exists(path):
                try:
                    row = np.load(path)
                    if row.shape != (self.num_omega,):
                        row = None
                except (OSError, ValueError, EOFError):
                    row = None  # unreadable / partly written by another process: recompute
            if row is None:
                omega = self.discrete_omega.numpy()
                sig = self.discrete_sigma.numpy()[idx]
                ls = np.arange(1000)[None]
                om = omega[..., None]
                f = ((2 * ls + 1) * np.exp(-ls * (ls + 1) * sig ** 2 / 2) * np.sin(om * (ls + 1 / 2)) / np.sin(om / 2)).sum(-1)
                row = (f * (1.0 - np.cos(omega)) / np.pi).cumsum() / self.num_omega * np.pi
                try:  # write to a private temporary file, then rename: readers never see a partial row
                    os.makedirs(self.cache_dir, exist_ok=True)
                    tmp = f"{path}.{os.getpid()}.tmp.npy"
                    np.save(tmp, row)
                    os.replace(tmp, path)
                except OSError:
                    pass
            self._cdf_rows[idx] = row
        return self._cdf_rows[idx]
