"""Ensemble metrics on C-alpha coordinates: the step after sampling, as batched tensor reductions that run
on whatever device holds the sampled coordinates (the reference works in numpy with a Python-level histogram per distance
channel — `np.apply_along_axis` over L(L-1)/2 columns).

Mirrors src/metrics/metrics.py of the reference:
  adjacent_ca_distance :12-23   distance_matrix_ca :26-37   pairwise_distance_ca :40-53   radius_of_gyration :56-80
  validity (_steric_clash) :83-124   bonding_validity :127-141   js_pwd :144-168   js_rg :198-216
`js_tica` (:171-195) needs `deeptime` (TICA), which is not installed: not provided.

All arithmetic is fp64 like numpy's.  Histograms follow `np.histogram` for uniform bins exactly (same index formula, same
edge corrections, right edge inclusive, out-of-range values dropped, a degenerate range widened by +-0.5), and the
Jensen-Shannon distance follows `scipy.spatial.distance.jensenshannon` (natural log, inputs normalised), so the rounded
results equal the reference's (tests/golden/metrics_*.npz, produced by the unmodified reference functions).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

PSEUDO_C = 1e-6


def _t(x) -> torch.Tensor:
    x = torch.as_tensor(x)
    assert x.dim() in (2, 3), f"CA coords should be 2D or 3D, got {tuple(x.shape)}"
    return x.to(torch.float64)


def adjacent_ca_distance(coords) -> torch.Tensor:
    """(..., L, 3) -> (..., L-1) distances between consecutive C-alpha atoms."""
    x = _t(coords)
    d = x[..., :-1, :] - x[..., 1:, :]
    return torch.sqrt((d * d).sum(-1))


def distance_matrix_ca(coords) -> torch.Tensor:
    """(..., L, 3) -> (..., L, L)."""
    x = _t(coords)
    d = x[..., None, :, :] - x[..., None, :]
    return torch.sqrt((d * d).sum(-1))


def pairwise_distance_ca(coords, k: int = 1) -> torch.Tensor:
    """(..., L, 3) -> (..., D): upper triangle (offset k) of the distance matrix in np.triu_indices order; only the D pairs
    are formed (the reference builds the full L x L matrix first)."""
    x = _t(coords)
    L = x.shape[-2]
    row, col = torch.triu_indices(L, L, offset=k, device=x.device)
    d = x[..., col, :] - x[..., row, :]   # the reference's matrix element [i, j] is |x_j - x_i| with the same operand order
    return torch.sqrt((d * d).sum(-1))


def radius_of_gyration(coords, masses=None) -> torch.Tensor:
    x = _t(coords)
    n = x.shape[-2]
    if masses is None:
        m = torch.ones(n, dtype=torch.float64, device=x.device)
    else:
        m = torch.as_tensor(masses, dtype=torch.float64, device=x.device)
        assert m.dim() == 1 and m.shape[0] == n, f"masses {tuple(m.shape)} != number of particles {n}"
    w = m / m.sum()
    c = x - x.mean(-2, keepdim=True)
    return ((c * c).sum(-1) * w).sum(-1) ** 0.5


def steric_clash(coords, ca_vdw_radius: float = 1.7, allowable_overlap: float = 0.4, k_exclusion: int = 0) -> torch.Tensor:
    """Number of C-alpha pairs closer than 2 r_vdw - overlap per conformation (reference `_steric_clash`)."""
    x = _t(coords)
    assert not torch.isnan(x).any(), "coords should not contain nan"
    assert k_exclusion >= 0, "k_exclusion should be non-negative"
    pwd = pairwise_distance_ca(x, k=k_exclusion + 1)
    assert pwd.dim() == 2, f"pwd should be 2D, got {tuple(pwd.shape)}"
    return (pwd < 2 * ca_vdw_radius - allowable_overlap).sum(-1)


def _round4(v) -> float:
    return float(np.around(float(v), decimals=4))


def validity(ca_coords_dict: Dict[str, object], **clash_kwargs) -> Dict[str, float]:
    """Fraction of conformations without any steric clash."""
    return {k: _round4(1.0 - (steric_clash(v, **clash_kwargs) > 0).double().mean()) for k, v in ca_coords_dict.items()}


def bonding_validity(ca_coords_dict: Dict[str, object], ref_key: str = "target", eps: float = 1e-6) -> Dict[str, float]:
    """Fraction of conformations whose consecutive C-alpha distances all stay below the reference ensemble's maximum."""
    adj = {k: adjacent_ca_distance(v) for k, v in ca_coords_dict.items()}
    thres = adj[ref_key].max() + 1e-6
    return {k: _round4((v < thres.to(v.device)).all(-1).sum().item() / len(v)) for k, v in adj.items()}


def histogram_columns(values: torch.Tensor, lo: torch.Tensor, hi: torch.Tensor, n_bins: int, weights: Optional[torch.Tensor] = None) -> torch.Tensor:
    """np.histogram(values[:, d], bins=n_bins, range=(lo[d], hi[d]), weights=weights) for every column d at once.
    values [B, D], lo / hi [D] -> counts [n_bins, D] (fp64)."""
    B, D = values.shape
    lo, hi = lo.clone(), hi.clone()
    same = lo == hi
    lo[same] -= 0.5
    hi[same] += 0.5
    keep = (values >= lo) & (values <= hi)
    idx = (((values - lo) / (hi - lo)) * n_bins).to(torch.int64)  # numpy's guess: ((a - first) / (last - first)) * n_bins, truncated
    idx = torch.where(idx == n_bins, idx - 1, idx).clamp_(0, n_bins - 1)
    # np.linspace(lo, hi, n_bins + 1): arange * step + start with step = (hi - lo) / n_bins, last edge set to hi
    edges = torch.arange(n_bins + 1, dtype=torch.float64, device=values.device)[:, None] * ((hi - lo) / n_bins)[None, :] + lo[None, :]
    edges[-1] = hi
    cols = torch.arange(D, device=values.device).expand(B, D)
    idx = idx - (values < edges[idx, cols]).to(torch.int64)
    idx = idx.clamp_(0, n_bins - 1)
    idx = idx + ((values >= edges[idx + 1, cols]) & (idx != n_bins - 1)).to(torch.int64)
    w = torch.ones(B, dtype=torch.float64, device=values.device) if weights is None else torch.as_tensor(weights, dtype=torch.float64, device=values.device)
    out = torch.zeros(n_bins, D, dtype=torch.float64, device=values.device)
    out.index_put_((idx[keep], cols[keep]), w[:, None].expand(B, D)[keep], accumulate=True)
    return out


def jensenshannon(p: torch.Tensor, q: torch.Tensor) -> torch.Tensor:
    """scipy.spatial.distance.jensenshannon(p, q, axis=0): sqrt of the JS divergence (natural log) of the normalised columns."""
    p = p / p.sum(0, keepdim=True)
    q = q / q.sum(0, keepdim=True)
    m = (p + q) / 2.0

    def rel_entr(a, b):
        return torch.where(a > 0, a * torch.log(a / b), torch.zeros_like(a))

    js = (rel_entr(p, m).sum(0) + rel_entr(q, m).sum(0)) / 2.0
    return torch.sqrt(js)


def _weights(ca_coords_dict, weights):
    weights = dict(weights or {})
    for k, v in ca_coords_dict.items():
        weights.setdefault(k, np.ones(len(v)))
    return weights


def js_pwd(ca_coords_dict: Dict[str, object], ref_key: str = "target", n_bins: int = 50, pwd_offset: int = 3, weights=None) -> Dict[str, float]:
    """Mean over distance channels of the JS distance between the per-channel distance histograms of each ensemble and of
    the reference ensemble (bins span the reference's min..max per channel)."""
    pwd = {k: pairwise_distance_ca(v, k=pwd_offset) for k, v in ca_coords_dict.items()}
    weights = _weights(ca_coords_dict, weights)
    lo, hi = pwd[ref_key].min(0).values, pwd[ref_key].max(0).values
    binned = {k: histogram_columns(v, lo.to(v.device), hi.to(v.device), n_bins, weights[k]) + PSEUDO_C for k, v in pwd.items()}
    res = {k: _round4(jensenshannon(v, binned[ref_key].to(v.device)).mean()) for k, v in binned.items() if k != ref_key}
    res[ref_key] = 0.0
    return res


def js_rg(ca_coords_dict: Dict[str, object], ref_key: str = "target", n_bins: int = 50, weights=None) -> Dict[str, float]:
    """JS distance between the radius-of-gyration histograms (bins span the reference's min..max)."""
    rg = {k: radius_of_gyration(v) for k, v in ca_coords_dict.items()}
    weights = _weights(ca_coords_dict, weights)
    lo, hi = rg[ref_key].min().reshape(1), rg[ref_key].max().reshape(1)
    binned = {k: histogram_columns(v[:, None], lo.to(v.device), hi.to(v.device), n_bins, weights[k]) + PSEUDO_C for k, v in rg.items()}
    res = {k: _round4(jensenshannon(v, binned[ref_key].to(v.device)).mean()) for k, v in binned.items() if k != ref_key}
    res[ref_key] = 0.0
    return res


def evaluate_ensembles(ca_coords_dict: Dict[str, object], ref_key: str = "target") -> Dict[str, Dict[str, float]]:
    """The four deeptime-free metrics of the reference's evaluation script in one call."""
    return {"validity": validity(ca_coords_dict), "bonding_validity": bonding_validity(ca_coords_dict, ref_key=ref_key),
            "js_pwd": js_pwd(ca_coords_dict, ref_key=ref_key), "js_rg": js_rg(ca_coords_dict, ref_key=ref_key)}
