"""CPU oracle for the Str2Str denoising hot path  --  TEST INFRASTRUCTURE, NOT THE PRODUCT.

A plain-PyTorch (fp32, CPU) restatement of the reference algorithm for the path BASELINE.json names:
score network forward (embedder + 4-block IPA trunk + frame update + psi head), the SE(3) diffusion step
(IGSO(3)/VP-SDE score, reverse step, perturbation) and the forward-backward sampler loop.  Written as
stateless functions over a flat parameter dict that uses the reference's ``state_dict`` keys.

Pinning: the reference ships no numerical tests for this path, so this file is pinned
against outputs of the reference itself: ``tests/golden/make_golden.py`` imports the unmodified reference
from /root/reference in the build container and stores its outputs in ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks this file against them on every run.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this module.  Every function cites the reference lines (relative to /root/reference) it follows.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

P = Dict[str, torch.Tensor]

# configs/model/diffusion.yaml:20-58
H, C, PQ, PV, NBLK = 8, 256, 8, 12, 4
COORD_SCALE = 0.1
MIN_B, MAX_B = 0.1, 20.0
MIN_SIGMA, MAX_SIGMA, NUM_SIGMA, NUM_OMEGA = 0.1, 1.5, 1000, 1000


# ------------------------------------------------------------------------------------------------
# rotations (src/common/rigid_utils.py, src/common/rotation3d.py)
# ------------------------------------------------------------------------------------------------
def quat_to_rotmat(q: torch.Tensor) -> torch.Tensor:
    """Quadratic form WITHOUT normalisation (rigid_utils.py:187-207, _QTR_MAT :163-185)."""
    a, b, c, d = q.unbind(-1)
    rows = [
        a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c),
        2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b),
        2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d,
    ]
    return torch.stack(rows, -1).reshape(q.shape[:-1] + (3, 3))


def quat_to_rotmat_normalised(q: torch.Tensor) -> torch.Tensor:
    """rotation3d.quaternion_to_matrix :41-70 (scales by 2/|q|^2)."""
    r, i, j, k = q.unbind(-1)
    s = 2.0 / (q * q).sum(-1)
    rows = [
        1 - s * (j * j + k * k), s * (i * j - k * r), s * (i * k + j * r),
        s * (i * j + k * r), 1 - s * (i * i + k * k), s * (j * k - i * r),
        s * (i * k - j * r), s * (j * k + i * r), 1 - s * (i * i + j * j),
    ]
    return torch.stack(rows, -1).reshape(q.shape[:-1] + (3, 3))


def rotmat_to_quat(m: torch.Tensor) -> torch.Tensor:
    """rotation3d.matrix_to_quaternion :102-161: four candidates, pick the one with the largest |component|."""
    m00, m01, m02 = m[..., 0, 0], m[..., 0, 1], m[..., 0, 2]
    m10, m11, m12 = m[..., 1, 0], m[..., 1, 1], m[..., 1, 2]
    m20, m21, m22 = m[..., 2, 0], m[..., 2, 1], m[..., 2, 2]
    q_abs = torch.stack(
        [1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22, 1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22], -1
    ).clamp(min=0).sqrt()
    cand = torch.stack(
        [
            torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], -1),
            torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], -1),
            torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], -1),
            torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], -1),
        ],
        -2,
    )
    cand = cand / (2.0 * q_abs[..., None].clamp(min=0.1))
    pick = q_abs.argmax(-1)
    return torch.gather(cand, -2, pick[..., None, None].expand(pick.shape + (1, 4))).squeeze(-2)


def _sinc_half(angle: torch.Tensor, half: torch.Tensor) -> torch.Tensor:
    """sin(angle/2)/angle with the reference's small-angle series (rotation3d.py:506-518,541-552)."""
    small = angle.abs() < 1e-6
    safe = torch.where(small, torch.ones_like(angle), angle)
    return torch.where(small, 0.5 - angle * angle / 48, torch.sin(half) / safe)


def axis_angle_to_quat(v: torch.Tensor) -> torch.Tensor:
    """rotation3d.axis_angle_to_quaternion :493-522."""
    ang = v.norm(p=2, dim=-1, keepdim=True)
    half = 0.5 * ang
    return torch.cat([torch.cos(half), v * _sinc_half(ang, half)], -1)


def quat_to_axis_angle(q: torch.Tensor) -> torch.Tensor:
    """rotation3d.quaternion_to_axis_angle :525-553 (no sign standardisation: angle in [0, 2pi))."""
    n = q[..., 1:].norm(p=2, dim=-1, keepdim=True)
    half = torch.atan2(n, q[..., :1])
    ang = 2 * half
    return q[..., 1:] / _sinc_half(ang, half)


def axis_angle_to_rotmat(v):
    return quat_to_rotmat_normalised(axis_angle_to_quat(v))  # rotation3d.py:461-474


def rotmat_to_axis_angle(m):
    return quat_to_axis_angle(rotmat_to_quat(m))  # rotation3d.py:477-490


def quat_mul(p: torch.Tensor, q: torch.Tensor) -> torch.Tensor:
    """Hamilton product (rigid_utils.py:229-265)."""
    a1, b1, c1, d1 = p.unbind(-1)
    a2, b2, c2, d2 = q.unbind(-1)
    return torch.stack(
        [
            a1 * a2 - b1 * b2 - c1 * c2 - d1 * d2,
            a1 * b2 + b1 * a2 + c1 * d2 - d1 * c2,
            a1 * c2 - b1 * d2 + c1 * a2 + d1 * b2,
            a1 * d2 + b1 * c2 - c1 * b2 + d1 * a2,
        ],
        -1,
    )


def rot_apply(R: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """R x, written out (rigid_utils.py:84-108)."""
    return (R * x[..., None, :]).sum(-1)


def compose_rotvec(v1: torch.Tensor, v2: torch.Tensor) -> torch.Tensor:
    """so3.py:13-19: fp64 3x3 product of the two rotations, back to an fp32 rotation vector."""
    R = axis_angle_to_rotmat(v1).double() @ axis_angle_to_rotmat(v2).double()
    return rotmat_to_axis_angle(R).to(v1.dtype)


# ------------------------------------------------------------------------------------------------
# embedder (src/models/net/denoising_ipa.py:13-159, src/common/geo_utils.py:44-56)
# ------------------------------------------------------------------------------------------------
def lin(p: P, name: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, p[name + ".weight"], p[name + ".bias"])


def lnorm(p: P, name: str, x: torch.Tensor) -> torch.Tensor:
    return F.layer_norm(x, x.shape[-1:], p[name + ".weight"], p[name + ".bias"], 1e-5)


def time_embedding(t: torch.Tensor, dim: int = 32, max_len: int = 10000) -> torch.Tensor:
    """denoising_ipa.py:34-46."""
    half = dim // 2
    freq = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(max_len) / (half - 1)))
    ang = (t * max_len).float()[:, None] * freq[None]
    return torch.cat([ang.sin(), ang.cos()], 1)


def index_embedding(idx: torch.Tensor, dim: int = 32, max_len: int = 2056) -> torch.Tensor:
    """denoising_ipa.py:13-31 (idx is int64; may be negative for pair offsets)."""
    k = torch.arange(dim // 2)
    ang = idx[..., None] * math.pi / (max_len ** (2 * k[None] / dim))
    return torch.cat([ang.sin(), ang.cos()], -1)


def distogram(pos: torch.Tensor, lo: float = 1e-5, hi: float = 20.0, nb: int = 22) -> torch.Tensor:
    """geo_utils.py:44-56: strict inequalities; the last bin's upper edge is 1e8."""
    d = torch.linalg.norm(pos[..., :, None, :] - pos[..., None, :, :], dim=-1)[..., None]
    lower = torch.linspace(lo, hi, nb)
    upper = torch.cat([lower[1:], lower.new_tensor([1e8])])
    return ((d > lower) * (d < upper)).to(pos.dtype)


def embedder(p: P, residue_idx, t, fixed_mask, sc_ca):
    """EmbeddingModule.forward, denoising_ipa.py:107-159. -> node [B,L,256], edge [B,L,L,128]."""
    B, L = residue_idx.shape
    tf = torch.cat([time_embedding(t)[:, None, :].expand(B, L, -1), fixed_mask[..., None].float()], -1)  # [B,L,33]
    node_in = torch.cat([tf, index_embedding(residue_idx)], -1).float()
    rel = residue_idx[:, :, None] - residue_idx[:, None, :]
    pair_in = torch.cat(
        [
            tf[:, :, None, :].expand(B, L, L, -1),
            tf[:, None, :, :].expand(B, L, L, -1),
            index_embedding(rel),
            distogram(sc_ca),
        ],
        -1,
    ).float()

    def mlp(base, x):
        x = F.relu(lin(p, base + ".0", x))
        x = F.relu(lin(p, base + ".2", x))
        return lnorm(p, base + ".5", lin(p, base + ".4", x))

    return mlp("embedder.node_embed", node_in), mlp("embedder.edge_embed", pair_in)


# ------------------------------------------------------------------------------------------------
# trunk (src/models/net/ipa.py, src/models/net/layers.py)
# ------------------------------------------------------------------------------------------------
def ipa(p: P, pre: str, s, z, quat, trans, mask):
    """InvariantPointAttention.forward, ipa.py:100-268. quat is used un-normalised; trans in nm."""
    B, L, _ = s.shape
    R = quat_to_rotmat(quat)  # [B,L,3,3]
    q = lin(p, pre + "linear_q", s).view(B, L, H, C)
    kv = lin(p, pre + "linear_kv", s).view(B, L, H, 2 * C)
    k, v = kv[..., :C], kv[..., C:]

    def points(name, n):  # x|y|z chunk layout, then to the global frame (ipa.py:144-171)
        raw = lin(p, pre + name, s)
        pts = torch.stack(raw.split(raw.shape[-1] // 3, dim=-1), -1)  # [B,L,H*n,3]
        pts = rot_apply(R[:, :, None], pts) + trans[:, :, None]
        return pts.view(B, L, H, n, 3)

    q_pts = points("linear_q_points", PQ)
    kv_pts = points("linear_kv_points", PQ + PV)
    k_pts, v_pts = kv_pts[..., :PQ, :], kv_pts[..., PQ:, :]

    bias = lin(p, pre + "linear_b", z)  # [B,L,L,H]
    a = torch.einsum("bihc,bjhc->bhij", q, k) * math.sqrt(1.0 / (3 * C))
    a = a + math.sqrt(1.0 / 3) * bias.permute(0, 3, 1, 2)
    d2 = ((q_pts[:, :, None] - k_pts[:, None]) ** 2).sum(-1)  # [B,L,L,H,PQ]
    hw = F.softplus(p[pre + "head_weights"]) * math.sqrt(1.0 / (3 * (PQ * 9.0 / 2)))
    pt = (d2 * hw[:, None]).sum(-1) * (-0.5)  # [B,L,L,H]
    a = a + pt.permute(0, 3, 1, 2)
    a = a + (1e5 * (mask[:, :, None] * mask[:, None, :] - 1))[:, None]
    a = torch.softmax(a, -1)  # [B,H,L,L]

    o = torch.einsum("bhij,bjhc->bihc", a, v).reshape(B, L, H * C)
    o_pt = torch.einsum("bhij,bjhpx->bihpx", a, v_pts)  # global frame
    o_pt = rot_apply(R.transpose(-1, -2)[:, :, None, None], o_pt - trans[:, :, None, None])  # ipa.py:239
    o_norm = torch.sqrt((o_pt ** 2).sum(-1) + 1e-8).reshape(B, L, H * PV)
    o_pt = o_pt.reshape(B, L, H * PV, 3)
    o_pair = torch.einsum("bhij,bijc->bihc", a, lin(p, pre + "down_z", z)).reshape(B, L, -1)
    feats = torch.cat([o, o_pt[..., 0], o_pt[..., 1], o_pt[..., 2], o_norm, o_pair], -1)
    return lin(p, pre + "linear_out", feats)


def transformer_layer(p: P, pre: str, x, key_bias, n_heads: int = 4):
    """One post-norm nn.TransformerEncoderLayer in eval mode (call site ipa.py:312-317,357).
    x: [B,L,D]; key_bias: [B,L] float ADDED to the logits of each key (the reference passes the float
    tensor 1-mask as src_key_padding_mask, which torch treats as an additive mask)."""
    B, L, D = x.shape
    hd = D // n_heads
    qkv = F.linear(x, p[pre + "self_attn.in_proj_weight"], p[pre + "self_attn.in_proj_bias"])
    q, k, v = [u.view(B, L, n_heads, hd).transpose(1, 2) for u in qkv.split(D, dim=-1)]
    att = (q @ k.transpose(-1, -2)) / math.sqrt(hd) + key_bias[:, None, None, :]
    y = (torch.softmax(att, -1) @ v).transpose(1, 2).reshape(B, L, D)
    x = lnorm(p, pre + "norm1", x + lin(p, pre + "self_attn.out_proj", y))
    ff = lin(p, pre + "linear2", F.relu(lin(p, pre + "linear1", x)))
    return lnorm(p, pre + "norm2", x + ff)


def edge_transition(p: P, pre: str, node, edge):
    """EdgeTransition.forward, layers.py:170-185."""
    B, L, _ = node.shape
    n = lin(p, pre + "initial_embed", node)
    x = torch.cat([edge, n[:, :, None, :].expand(B, L, L, -1), n[:, None, :, :].expand(B, L, L, -1)], -1)
    h = F.relu(lin(p, pre + "trunk.0", x))
    h = F.relu(lin(p, pre + "trunk.2", h))
    return lnorm(p, pre + "layer_norm", lin(p, pre + "final_layer", h + x))


def frame_update(quat, trans, upd, m):
    """Rigid.compose_q_update_vec, rigid_utils.py:1042-1066 -> Rotation :590-619, quat_multiply_by_vec :268."""
    zero = torch.zeros_like(upd[..., :1])
    dq = quat_mul(quat, torch.cat([zero, upd[..., :3]], -1)) * m
    new_q = quat + dq
    new_q = new_q / torch.linalg.norm(new_q, dim=-1, keepdim=True)
    new_t = trans + rot_apply(quat_to_rotmat(quat), upd[..., 3:]) * m
    return new_q, new_t


def trunk(p: P, node, edge, rigids_t, node_mask, fixed_mask):
    """TranslationIPA.forward, ipa.py:331-387. Returns (quat, trans[A], psi, node, edge)."""
    t = "translator.trunk."
    diffuse = (1 - fixed_mask) * node_mask
    edge_mask = node_mask[..., None] * node_mask[..., None, :]
    quat, trans = rigids_t[..., :4].float(), rigids_t[..., 4:].float() * COORD_SCALE
    init_node = node
    for b in range(NBLK):
        upd = ipa(p, f"{t}ipa_{b}.", node, edge, quat, trans, node_mask) * node_mask[..., None]
        node = lnorm(p, f"{t}ipa_ln_{b}", node + upd)
        x = torch.cat([node, lin(p, f"{t}skip_embed_{b}", init_node)], -1)
        for layer in range(2):
            x = transformer_layer(p, f"{t}transformer_{b}.layers.{layer}.", x, 1.0 - node_mask)
        node = node + lin(p, f"{t}linear_{b}", x)
        nt = f"{t}node_transition_{b}."
        h = F.relu(lin(p, nt + "linear_1", node))
        h = F.relu(lin(p, nt + "linear_2", h))
        node = lnorm(p, nt + "ln", lin(p, nt + "linear_3", h) + node) * node_mask[..., None]
        upd6 = lin(p, f"{t}bb_update_{b}.linear", node * diffuse[..., None])
        quat, trans = frame_update(quat, trans, upd6, diffuse[..., None])
        if b < NBLK - 1:
            edge = edge_transition(p, f"{t}edge_transition_{b}.", node, edge) * edge_mask[..., None]
    tp = "translator.torsion_pred."  # layers.py:199-213 (linear_3 is unused)
    h = lin(p, tp + "linear_2", F.relu(lin(p, tp + "linear_1", node))) + node
    u = lin(p, tp + "linear_final", h)
    psi = u / torch.sqrt(torch.clamp((u ** 2).sum(-1, keepdim=True), min=1e-8))
    return quat, trans / COORD_SCALE, psi, node, edge


def backbone_atoms(quat, trans, psi, aatype):
    """compute_backbone, all_atom.py:141-173 restricted to what it can produce: N, CA, C, CB (backbone
    group) and O (psi group).  -> atom37 [.,37,3] (slots 0..4 = N,CA,C,CB,O), atom14 [.,14,3]."""
    from str2str_b200.backbone_constants import BACKBONE_MASK, BACKBONE_POS, BB_FRAME_VALID, PSI_FRAME

    pos = torch.tensor(BACKBONE_POS, dtype=torch.float32)[aatype]  # [.,5,3] N CA C O CB
    amask = torch.tensor(BACKBONE_MASK, dtype=torch.float32)[aatype]
    psi_f = torch.tensor(PSI_FRAME, dtype=torch.float32)[aatype]  # [.,4,4]
    valid = torch.tensor(BB_FRAME_VALID, dtype=torch.float32)[aatype]
    R = quat_to_rotmat(quat)
    sin, cos = psi[..., 0], psi[..., 1]
    one, zero = torch.ones_like(sin), torch.zeros_like(sin)
    Rx = torch.stack([one, zero, zero, zero, cos, -sin, zero, sin, cos], -1).reshape(sin.shape + (3, 3))
    Rpsi = psi_f[..., :3, :3] @ Rx  # default frame o torsion rotation (all_atom.py:57-59)
    local_o = rot_apply(Rpsi, pos[..., 3, :]) + psi_f[..., :3, 3]
    local = torch.stack([pos[..., 0, :], pos[..., 1, :], pos[..., 2, :], local_o, pos[..., 4, :]], -2)
    local = local * torch.stack([valid, valid, valid, one, valid], -1)[..., None]
    glob = (rot_apply(R[..., None, :, :], local) + trans[..., None, :]) * amask[..., None]
    atom14 = glob.new_zeros(glob.shape[:-2] + (14, 3))
    atom14[..., :5, :] = glob
    atom37 = glob.new_zeros(glob.shape[:-2] + (37, 3))
    atom37[..., :3, :] = glob[..., :3, :]
    atom37[..., 3, :] = glob[..., 4, :]
    atom37[..., 4, :] = glob[..., 3, :]
    return atom37, atom14


def denoising_net(p: P, feats: Dict[str, torch.Tensor], return_intermediates: bool = False):
    """DenoisingNet.forward, denoising_ipa.py:171-211. 'rigids' is returned as tensor_7 (quat, trans[A])."""
    node_mask = feats["residue_mask"].float()
    fixed = feats["fixed_mask"].float()
    node, edge = embedder(p, feats["residue_idx"], feats["t"], fixed, feats["sc_ca_t"])
    node = node * node_mask[..., None]
    edge = edge * (node_mask[..., None] * node_mask[..., None, :])[..., None]
    quat, trans, psi, node_out, edge_out = trunk(p, node, edge, feats["rigids_t"], node_mask, fixed)
    gt_psi = feats["torsion_angles_sin_cos"][..., 2, :]
    psi = gt_psi * fixed[..., None] + psi * (1 - fixed[..., None])
    aatype = feats["aatype"] if "aatype" in feats else torch.zeros(quat.shape[:-1], dtype=torch.long)
    atom37, atom14 = backbone_atoms(quat, trans, psi.float(), aatype)
    out = {"rigids": torch.cat([quat, trans], -1), "psi": psi, "atom37": atom37, "atom14": atom14}
    if return_intermediates:
        out.update(node_embed=node, edge_embed=edge, node_out=node_out, edge_out=edge_out)
    return out


# ------------------------------------------------------------------------------------------------
# SE(3) diffusion (src/models/score/{so3,r3,frame}.py)
# ------------------------------------------------------------------------------------------------
def sigma_of_t(t: torch.Tensor) -> torch.Tensor:
    """so3.py:216-223 (logarithmic schedule)."""
    return torch.log(t * math.exp(MAX_SIGMA) + (1 - t) * math.exp(MIN_SIGMA))


def discrete_sigma() -> torch.Tensor:
    return sigma_of_t(torch.linspace(0.0, 1.0, NUM_SIGMA))  # so3.py:205-209


def sigma_index(t: torch.Tensor) -> torch.Tensor:
    """so3.py:211-214,236-238: np.digitize(sigma(t), grid) - 1 (integer bucket; must be bit-exact)."""
    return torch.as_tensor(np.digitize(sigma_of_t(t).cpu().numpy(), discrete_sigma().cpu().numpy()) - 1, dtype=torch.long)


def rot_g2(t: torch.Tensor) -> torch.Tensor:
    """g(t)^2 of so3.py:225-234 (uses the un-quantised sigma(t))."""
    s = sigma_of_t(t)
    return torch.sqrt(2 * (math.exp(MAX_SIGMA) - math.exp(MIN_SIGMA)) * s / torch.exp(s)) ** 2


def beta_int(t):
    return t * MIN_B + 0.5 * (t ** 2) * (MAX_B - MIN_B)  # r3.py:40-41


def b_of_t(t):
    return MIN_B + t * (MAX_B - MIN_B)  # r3.py:26-29


def igso3_score_scale(omega: torch.Tensor, sigma: torch.Tensor, n_terms: int = 1000) -> torch.Tensor:
    """d/domega log f(omega; sigma) with the reference's 1000-term fp32 series
    (igso3_expansion so3.py:21-62 with use_torch, score so3.py:85-130). omega [B,L], sigma [B,1]."""
    ls = torch.arange(n_terms)[None, None]
    om, sg = omega[..., None], sigma[..., None]
    f = ((2 * ls + 1) * torch.exp(-ls * (ls + 1) * sg ** 2 / 2) * torch.sin(om * (ls + 1 / 2)) / torch.sin(om / 2)).sum(-1)
    hi = torch.sin(om * (ls + 1 / 2))
    dhi = (ls + 1 / 2) * torch.cos(om * (ls + 1 / 2))
    lo = torch.sin(om / 2)
    dlo = 1 / 2 * torch.cos(om / 2)
    df = ((2 * ls + 1) * torch.exp(-ls * (ls + 1) * sg ** 2 / 2) * (lo * dhi - hi * dlo) / lo ** 2).sum(-1)
    return df / (f + 1e-4)


def rot_score(vec: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """SO3Diffuser.score, so3.py:274-309 (use_cached_score=False)."""
    omega = torch.linalg.norm(vec, dim=-1) + 1e-6
    sigma = discrete_sigma()[sigma_index(t)]
    return igso3_score_scale(omega, sigma[:, None])[..., None] * vec / (omega[..., None] + 1e-6)


def trans_score(x_t, x_0, t):
    """R3Diffuser.score(scale=True), r3.py:133-137. Inputs in Angstrom."""
    tt = t[:, None, None]
    x_t, x_0 = x_t * COORD_SCALE, x_0 * COORD_SCALE
    return -(x_t - torch.exp(-0.5 * beta_int(tt)) * x_0) / (1.0 - torch.exp(-beta_int(tt)))


def diffuser_score(r0_7: torch.Tensor, rt_7: torch.Tensor, t: torch.Tensor, mask: Optional[torch.Tensor]):
    """FrameDiffuser.score, frame.py:109-143. Both frames are quaternion-format (tensor_7)."""
    q0, qt = r0_7[..., :4], rt_7[..., :4]
    q0_inv = rotmat_to_quat(quat_to_rotmat(q0 * q0.new_tensor([1.0, -1.0, -1.0, -1.0]) / (q0 ** 2).sum(-1, keepdim=True)))
    q_t = rotmat_to_quat(quat_to_rotmat(qt))
    rs = rot_score(quat_to_axis_angle(quat_mul(q0_inv, q_t)), t)
    ts = trans_score(rt_7[..., 4:], r0_7[..., 4:], t)
    if mask is not None:
        rs, ts = rs * mask[..., None], ts * mask[..., None]
    return rs, ts


def diffuser_reverse(rt_7, rot_s, trans_s, t, dt, diffuse_mask=None, noise_scale=1.0, probability_flow=True,
                     rot_noise=None, trans_noise=None):
    """FrameDiffuser.reverse(center_trans=True), frame.py:153-210 -> so3.py:333-371, r3.py:79-125.
    Returns tensor_7 (what Rigid.to_tensor_7 gives for the rot-mat Rigid the reference returns)."""
    rotvec_t = rotmat_to_axis_angle(quat_to_rotmat(rt_7[..., :4]))
    x_t = rt_7[..., 4:]
    tt = t[:, None, None]
    half = 0.5 if probability_flow else 1.0
    # rotation (so3.py:357-370)
    g2 = rot_g2(tt)
    perturb = -1.0 * g2 * rot_s * dt * half
    if not probability_flow:
        perturb = perturb + torch.sqrt(g2) * np.sqrt(dt) * (noise_scale * rot_noise)
    rotvec_n = compose_rotvec(rotvec_t, -1.0 * perturb)
    # translation (r3.py:101-124)
    x = x_t * COORD_SCALE
    bt = b_of_t(tt)
    drift = (-0.5 * bt * x - torch.sqrt(bt) ** 2 * trans_s) * dt * half
    if not probability_flow:
        drift = drift + torch.sqrt(bt) * math.sqrt(dt) * (noise_scale * trans_noise)
    x_n = x - drift
    com = x_n.sum(-2) / torch.ones_like(x[..., 0]).sum(-1)[..., None]  # mean over ALL rows (r3.py:117-122)
    x_n = (x_n - com[..., None, :]) / COORD_SCALE
    if diffuse_mask is not None:
        m = diffuse_mask[..., None]
        x_n = m * x_n + (1 - m) * x_t
        rotvec_n = m * rotvec_n + (1 - m) * rotvec_t
    quat = rotmat_to_quat(axis_angle_to_rotmat(rotvec_n).float())  # frame.py:9-15, rigid_utils.py:1203-1215
    return torch.cat([quat, x_n.float()], -1)


_CDF_ROWS: Dict[int, np.ndarray] = {}


def igso3_cdf_row(sigma_idx: int) -> np.ndarray:
    """One row of SO3Diffuser._cdf (so3.py:171-183): fp64 series on the omega grid, density * (1-cos)/pi,
    cumulative sum / num_omega * pi."""
    if sigma_idx not in _CDF_ROWS:
        omega = torch.linspace(0, np.pi, NUM_OMEGA + 1)[1:].numpy()
        sig = discrete_sigma().numpy()[sigma_idx]
        ls = np.arange(1000)[None]
        om = omega[..., None]
        f = ((2 * ls + 1) * np.exp(-ls * (ls + 1) * sig ** 2 / 2) * np.sin(om * (ls + 1 / 2)) / np.sin(om / 2)).sum(-1)
        pdf = f * (1.0 - np.cos(omega)) / np.pi
        _CDF_ROWS[sigma_idx] = pdf.cumsum() / NUM_OMEGA * np.pi
    return _CDF_ROWS[sigma_idx]


def forward_marginal(r0_rot: torch.Tensor, r0_trans: torch.Tensor, t: torch.Tensor, diffuse_mask,
                     axis_noise, u_noise, trans_noise):
    """FrameDiffuser.forward_marginal, frame.py:36-107, with the three random draws passed in
    (order in the reference: randn axis so3.py:259, rand angle so3.py:262, randn_like r3.py:66).
    r0_rot: rotation matrices [B,L,3,3]; returns rigids_t as tensor_7."""
    rot0 = rotmat_to_axis_angle(r0_rot)
    axis = axis_noise / torch.linalg.norm(axis_noise, dim=-1, keepdim=True)
    omega_grid = torch.linspace(0, np.pi, NUM_OMEGA + 1)[1:].numpy()
    idx = sigma_index(t)
    ang = np.stack([np.interp(u_noise[i].numpy(), igso3_cdf_row(int(idx[i])), omega_grid) for i in range(t.shape[0])])
    rot_t = compose_rotvec(rot0, axis * torch.as_tensor(ang, dtype=axis.dtype)[..., None])
    tt = t[:, None, None]
    x0 = r0_trans * COORD_SCALE
    x_t = (trans_noise * torch.sqrt(1 - torch.exp(-beta_int(tt))) + torch.exp(-0.5 * beta_int(tt)) * x0) / COORD_SCALE
    if diffuse_mask is not None:
        m = torch.as_tensor(diffuse_mask, dtype=x_t.dtype)[..., None]
        rot_t = m * rot_t + (1 - m) * rot0
        x_t = m * x_t + (1 - m) * r0_trans
    quat = rotmat_to_quat(axis_angle_to_rotmat(rot_t).float())
    return torch.cat([quat, x_t.float()], -1)


# ------------------------------------------------------------------------------------------------
# sampler loop (src/models/diffusion_module.py:260-334)
# ------------------------------------------------------------------------------------------------
def forward_backward(p: P, feats: Dict[str, torch.Tensor], rigids_t: torch.Tensor, t_delta: float,
                     num_timesteps: int, min_t: float = 0.01, self_conditioning: bool = True,
                     noise_scale: float = 1.0, probability_flow: bool = True, noises=None, trace=None):
    """forward_backward after the perturbation: `rigids_t` [B,L,7] is the perturbed state at T.
    Returns (final rigids tensor_7, psi, atom37).  `trace` (a list) collects per-step sigma indices."""
    T = t_delta if t_delta > 0 else 1.0
    n = int(float(num_timesteps) * T)
    dt = 1.0 / n
    ts = np.linspace(min_t, T, n)[::-1]
    B = rigids_t.shape[0]
    f = dict(feats)
    f["rigids_t"] = rigids_t
    diffuse = (1 - f["fixed_mask"]) * f["residue_mask"]
    f["sc_ca_t"] = torch.zeros_like(rigids_t[..., 4:])
    with torch.no_grad():
        if self_conditioning:
            f["t"] = ts[0] * torch.ones(B)
            f["sc_ca_t"] = denoising_net(p, f)["rigids"][..., 4:]
        for k, t in enumerate(ts):
            f["t"] = t * torch.ones(B)
            out = denoising_net(p, f)
            if t == min_t:
                pred = out["rigids"]
            else:
                if self_conditioning:
                    f["sc_ca_t"] = out["rigids"][..., 4:]
                rs, tsc = diffuser_score(out["rigids"], f["rigids_t"], f["t"], f["residue_mask"])
                if trace is not None:
                    trace.append(int(sigma_index(f["t"])[0]))
                rn, tn = (None, None) if noises is None else noises[k]
                pred = diffuser_reverse(f["rigids_t"], rs, tsc, f["t"], dt, diffuse, noise_scale, probability_flow, rn, tn)
                f["rigids_t"] = pred
        aatype = f["aatype"] if "aatype" in f else torch.zeros(pred.shape[:-1], dtype=torch.long)
        atom37, _ = backbone_atoms(pred[..., :4], pred[..., 4:], out["psi"].float(), aatype)
    return pred, out["psi"], atom37
