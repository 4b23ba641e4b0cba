"""Deterministic synthetic weights and inputs for parity tests and the benchmark.

No trained Str2Str checkpoint is reachable offline (the reference README links a Google-Drive file), and a
freshly constructed reference net is the identity on frames because every ``init="final"`` matrix is zero
(reference src/models/net/layers.py:52-54,121-122).  Tests and bench therefore use the state dict built
here: same keys and shapes as the reference ``DenoisingNet.state_dict()``, values drawn
from per-tensor seeded CPU generators so that they do not depend on creation order, machine or the
reference being importable.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, List, Tuple

import torch

# configs/model/diffusion.yaml:20-40 of the reference
C_S, C_Z, C_HIDDEN, N_HEADS, N_QK_PTS, N_V_PTS = 256, 128, 256, 8, 8, 12
N_BLOCKS, SKIP_DIM, TFM_HEADS, TFM_LAYERS, INIT_DIM, N_BINS = 4, 64, 4, 2, 32, 22
D_TFM = C_S + SKIP_DIM


def param_spec() -> List[Tuple[str, Tuple[int, ...], str]]:
    """(name, shape, kind) for every tensor of the reference state dict, in a fixed order.

    kind: 'lecun' | 'relu' | 'final' (zero-initialised in the reference) | 'bias' | 'ln_w' | 'ln_b' | 'headw'.
    """
    spec: List[Tuple[str, Tuple[int, ...], str]] = []

    def lin(name, out_d, in_d, kind="lecun"):
        spec.append((name + ".weight", (out_d, in_d), kind))
        spec.append((name + ".bias", (out_d,), "bias"))

    def ln(name, d):
        spec.append((name + ".weight", (d,), "ln_w"))
        spec.append((name + ".bias", (d,), "ln_b"))

    node_in = INIT_DIM + 1 + INIT_DIM
    edge_in = 2 * (INIT_DIM + 1) + INIT_DIM + N_BINS
    for base, d_in, d in (("embedder.node_embed", node_in, C_S), ("embedder.edge_embed", edge_in, C_Z)):
        lin(base + ".0", d, d_in)
        lin(base + ".2", d, d)
        lin(base + ".4", d, d)
        ln(base + ".5", d)
    t = "translator.trunk."
    for b in range(N_BLOCKS):
        ipa = f"{t}ipa_{b}."
        spec.append((ipa + "head_weights", (N_HEADS,), "headw"))
        lin(ipa + "linear_q", N_HEADS * C_HIDDEN, C_S)
        lin(ipa + "linear_kv", 2 * N_HEADS * C_HIDDEN, C_S)
        lin(ipa + "linear_q_points", N_HEADS * N_QK_PTS * 3, C_S)
        lin(ipa + "linear_kv_points", N_HEADS * (N_QK_PTS + N_V_PTS) * 3, C_S)
        lin(ipa + "linear_b", N_HEADS, C_Z)
        lin(ipa + "down_z", C_Z // 4, C_Z)
        lin(ipa + "linear_out", C_S, N_HEADS * (C_Z // 4 + C_HIDDEN + N_V_PTS * 4), "final")
        ln(f"{t}ipa_ln_{b}", C_S)
        lin(f"{t}skip_embed_{b}", SKIP_DIM, C_S, "final")
        for layer in range(TFM_LAYERS):
            tl = f"{t}transformer_{b}.layers.{layer}."
            spec.append((tl + "self_attn.in_proj_weight", (3 * D_TFM, D_TFM), "lecun"))
            spec.append((tl + "self_attn.in_proj_bias", (3 * D_TFM,), "bias"))
            lin(tl + "self_attn.out_proj", D_TFM, D_TFM)
            lin(tl + "linear1", D_TFM, D_TFM, "relu")
            lin(tl + "linear2", D_TFM, D_TFM)
            ln(tl + "norm1", D_TFM)
            ln(tl + "norm2", D_TFM)
        lin(f"{t}linear_{b}", C_S, D_TFM, "final")
        nt = f"{t}node_transition_{b}."
        lin(nt + "linear_1", C_S, C_S, "relu")
        lin(nt + "linear_2", C_S, C_S, "relu")
        lin(nt + "linear_3", C_S, C_S, "final")
        ln(nt + "ln", C_S)
        lin(f"{t}bb_update_{b}.linear", 6, C_S, "final")
        if b < N_BLOCKS - 1:
            et = f"{t}edge_transition_{b}."
            hid = C_Z + 2 * (C_S // 2)
            lin(et + "initial_embed", C_S // 2, C_S, "relu")
            lin(et + "trunk.0", hid, hid, "relu")
            lin(et + "trunk.2", hid, hid, "relu")
            lin(et + "final_layer", C_Z, hid, "final")
            ln(et + "layer_norm", C_Z)
    tp = "translator.torsion_pred."
    lin(tp + "linear_1", C_S, C_S, "relu")
    lin(tp + "linear_2", C_S, C_S, "relu")
    lin(tp + "linear_3", C_S, C_S, "final")  # registered but unused by forward (layers.py:194,199-213)
    lin(tp + "linear_final", 2, C_S, "final")
    return spec


def make_state_dict(seed: int = 0, final_scale: float = 0.02, bias_scale: float = 0.02) -> Dict[str, torch.Tensor]:
    """Synthetic fp32 CPU state dict with the reference's keys.

    ``final_scale`` sets the std (in units of 1/sqrt(fan_in)) of the matrices the reference zero-initialises;
    0.02 gives a well-conditioned 100-step denoising map, 0.1 is the stress setting.
    """
    sd: Dict[str, torch.Tensor] = {}
    for name, shape, kind in param_spec():
        g = torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
        r = torch.randn(shape, generator=g, dtype=torch.float32)
        if kind in ("lecun", "relu", "final"):
            fan_in = shape[1]
            gain = {"lecun": 1.0, "relu": math.sqrt(2.0), "final": final_scale}[kind]
            # the two 6-/2-wide heads move frames directly; keep them on the same small scale
            v = r * (gain / math.sqrt(fan_in))
        elif kind == "bias":
            v = r * bias_scale
        elif kind == "ln_w":
            v = 1.0 + 0.05 * r
        elif kind == "ln_b":
            v = 0.02 * r
        elif kind == "headw":
            v = 0.541324854612918 + 0.1 * r
        else:  # pragma: no cover
            raise ValueError(kind)
        sd[name] = v.contiguous()
    return sd


def make_backbone(length: int, seed: int = 7):
    """Synthetic protein backbone: Calpha random walk with 3.8 A steps, centred; random unit-quaternion frames.

    Returns (quat [L,4] wxyz, trans [L,3] in Angstrom), fp32 CPU (SURVEY.md §8d synthetic inputs).
    """
    g = torch.Generator().manual_seed(seed)
    steps = torch.randn(length, 3, generator=g)
    steps = 3.8 * steps / steps.norm(dim=-1, keepdim=True)
    ca = torch.cumsum(steps, dim=0)
    ca = ca - ca.mean(dim=0, keepdim=True)
    q = torch.randn(length, 4, generator=g)
    q = q / q.norm(dim=-1, keepdim=True)
    return q.float().contiguous(), ca.float().contiguous()


def make_features(batch: int, length: int, seed: int = 7, n_pad: int = 0, n_fixed: int = 0, random_aatype: bool = False):
    """Feature dict with the reference's keys/dtypes (SURVEY.md §8b 'Batch dict'), all rows identical
    (one protein replicated ``batch`` times, as predict_step does, diffusion_module.py:269-272)."""
    g = torch.Generator().manual_seed(seed + 1000)
    residue_mask = torch.ones(batch, length, dtype=torch.float64)
    if n_pad:
        residue_mask[:, length - n_pad:] = 0.0
    fixed_mask = torch.zeros(batch, length, dtype=torch.float64)
    if n_fixed:
        fixed_mask[:, :n_fixed] = 1.0
    idx = torch.arange(length, dtype=torch.int64)
    if length > 8:
        idx = idx + (idx >= length // 2).long() * 5  # a chain break: offsets are not just i-j
    residue_idx = idx[None].repeat(batch, 1)
    aatype = torch.randint(0, 20, (length,), generator=g) if random_aatype else torch.zeros(length, dtype=torch.int64)
    tors = torch.randn(length, 7, 2, generator=g, dtype=torch.float64)
    tors = tors / tors.norm(dim=-1, keepdim=True)
    return {
        "aatype": aatype[None].repeat(batch, 1).contiguous(),
        "residue_mask": residue_mask,
        "fixed_mask": fixed_mask,
        "residue_idx": residue_idx.contiguous(),
        "torsion_angles_sin_cos": tors[None].repeat(batch, 1, 1, 1).contiguous(),
    }
