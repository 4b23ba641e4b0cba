"""Generate the golden vectors in tests/golden/ by running the UNMODIFIED reference on CPU.

Build-container only (needs /root/reference):   python tests/golden/make_golden.py
The reference has no numerical tests for this path, so these outputs are what pins the
oracle (oracle/str2str_oracle.py) and, through it, the CUDA path.  Weights and inputs come from
str2str_b200.synthetic (seeded, reference-independent); only the OUTPUTS stored here come from the reference.
"""
import os
import sys
import tempfile
from copy import deepcopy

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
import refshim  # noqa: E402

refshim.install()
from src.common.all_atom import compute_backbone  # noqa: E402
from src.common.rigid_utils import Rigid, Rotation  # noqa: E402
from src.models.net.denoising_ipa import DenoisingNet, EmbeddingModule  # noqa: E402
from src.models.net.ipa import TranslationIPA  # noqa: E402
from src.models.score.frame import FrameDiffuser  # noqa: E402
from src.models.score.r3 import R3Diffuser  # noqa: E402
from src.models.score.so3 import SO3Diffuser  # noqa: E402

from str2str_b200 import synthetic  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(8)


def build_reference(final_scale=0.02):
    """Instantiate the reference modules with the kwargs of configs/model/diffusion.yaml:16-58."""
    torch.manual_seed(0)
    np.random.seed(0)
    net = DenoisingNet(
        embedder=EmbeddingModule(init_embed_size=32, node_embed_size=256, edge_embed_size=128, num_bins=22,
                                 min_bin=1e-5, max_bin=20.0, self_conditioning=True),
        translator=TranslationIPA(c_s=256, c_z=128, coordinate_scaling=0.1, no_ipa_blocks=4, skip_embed_size=64,
                                  transformer_num_heads=4, transformer_num_layers=2, c_hidden=256, no_heads=8,
                                  no_qk_points=8, no_v_points=12, dropout=0.0),
    )
    sd = synthetic.make_state_dict(seed=0, final_scale=final_scale)
    assert set(sd) == set(net.state_dict()), set(sd) ^ set(net.state_dict())
    for k, v in net.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
    net.load_state_dict(sd, strict=True)
    net.eval()
    cache = os.path.join(tempfile.gettempdir(), "str2str_igso3_cache")
    diffuser = FrameDiffuser(
        trans_diffuser=R3Diffuser(min_b=0.1, max_b=20.0, coordinate_scaling=0.1),
        rot_diffuser=SO3Diffuser(num_omega=1000, num_sigma=1000, min_sigma=0.1, max_sigma=1.5,
                                 schedule="logarithmic", cache_dir=cache, use_cached_score=False),
        min_t=1e-2,
    )
    return net, diffuser


def rigid_from_quat_trans(q, x):
    """What predict_step builds from rigidgroups_gt_frames (diffusion_module.py:340-345): a rot-mat Rigid."""
    from oracle.str2str_oracle import quat_to_rotmat

    return Rigid(Rotation(rot_mats=quat_to_rotmat(q)), x)


def reference_forward_backward(net, diffuser, feats, rigids_t, t_delta, num_timesteps, min_t=0.01,
                               noise_scale=1.0, probability_flow=True):
    """The loop of diffusion_module.py:260-334 after the perturbation, calling the imported reference."""
    T = t_delta
    n = int(float(num_timesteps) * T)
    dt = 1.0 / n
    ts = np.linspace(min_t, T, n)[::-1]
    B = rigids_t.shape[0]
    _feats = deepcopy(feats)
    _feats["rigids_t"] = rigids_t
    sig_idx = []
    with torch.no_grad():
        diffuse_mask = (1 - _feats["fixed_mask"]) * _feats["residue_mask"]
        _feats["sc_ca_t"] = torch.zeros_like(rigids_t[..., 4:])
        _feats["t"] = ts[0] * torch.ones(B)
        _feats["sc_ca_t"] = net(_feats, as_tensor_7=True)["rigids"][..., 4:]
        for t in ts:
            _feats["t"] = t * torch.ones(B)
            out = net(_feats, as_tensor_7=False)
            if t == min_t:
                rigids_pred = out["rigids"]
            else:
                _feats["sc_ca_t"] = out["rigids"].to_tensor_7()[..., 4:]
                sig_idx.append(int(diffuser.rot_diffuser.t_to_idx(_feats["t"])[0]))
                sc = diffuser.score(rigids_0=out["rigids"], rigids_t=Rigid.from_tensor_7(_feats["rigids_t"]),
                                    t=_feats["t"], mask=_feats["residue_mask"])
                rigids_pred = diffuser.reverse(rigids_t=Rigid.from_tensor_7(_feats["rigids_t"]),
                                               rot_score=sc["rot_score"], trans_score=sc["trans_score"],
                                               t=_feats["t"], dt=dt, diffuse_mask=diffuse_mask, center_trans=True,
                                               noise_scale=noise_scale, probability_flow=probability_flow)
                _feats["rigids_t"] = rigids_pred.to_tensor_7()
        atom37 = compute_backbone(rigids_pred, out["psi"], aatype=_feats["aatype"])[0]
    return rigids_pred.to_tensor_7(), out["psi"], atom37, np.array(sig_idx)


def npz(name, **arrs):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **{k: (v.numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrs.items()})
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def trajectory(net, diffuser, name, B, L, n, n_pad, n_fixed, seed):
    """Full forward-backward trajectory of the unmodified reference from a perturbation drawn at fixed seeds."""
    feats = synthetic.make_features(B, L, seed=seed, n_pad=n_pad, n_fixed=n_fixed, random_aatype=True)
    q, x = synthetic.make_backbone(L, seed=seed)
    r0 = rigid_from_quat_trans(q[None].repeat(B, 1, 1), x[None].repeat(B, 1, 1))
    torch.manual_seed(123)
    np.random.seed(123)
    rigids_t = diffuser.forward_marginal(rigids_0=r0, t=0.5 * torch.ones(B), diffuse_mask=feats["residue_mask"],
                                         as_tensor_7=True)["rigids_t"]
    fin, psi, atom37, sig = reference_forward_backward(net, diffuser, feats, rigids_t, 0.5, 2 * n)
    npz(name, rigids_t=rigids_t, final_rigids=fin, final_psi=psi, final_atom37=atom37[..., :5, :], sigma_idx=sig,
        meta=np.array([B, L, n, n_pad, n_fixed, seed]))


def teacher_forced(name, B, L, n, n_pad, seed, final_scale):
    """Per-step records of the unmodified reference on the STRESS fixture (final_scale = 0.1): the state
    entering every iteration (rigids_t, sc_ca_t, t) and what the reference makes of it (network output, next state), so
    that a candidate can be driven step by step from the reference's own states and compared one step at a time."""
    net, diffuser = build_reference(final_scale)
    feats = synthetic.make_features(B, L, seed=seed, n_pad=n_pad, n_fixed=0, random_aatype=True)
    q, x = synthetic.make_backbone(L, seed=seed)
    r0 = rigid_from_quat_trans(q[None].repeat(B, 1, 1), x[None].repeat(B, 1, 1))
    torch.manual_seed(123)
    np.random.seed(123)
    rigids_t = diffuser.forward_marginal(rigids_0=r0, t=0.5 * torch.ones(B), diffuse_mask=feats["residue_mask"],
                                         as_tensor_7=True)["rigids_t"]
    ts = np.linspace(0.01, 0.5, n)[::-1]
    dt = 1.0 / n
    _f = deepcopy(feats)
    _f["rigids_t"] = rigids_t
    rec = dict(state=[], sc=[], out=[], psi=[], nxt=[])
    with torch.no_grad():
        diffuse_mask = (1 - _f["fixed_mask"]) * _f["residue_mask"]
        _f["sc_ca_t"] = torch.zeros_like(rigids_t[..., 4:])
        _f["t"] = ts[0] * torch.ones(B)
        _f["sc_ca_t"] = net(_f, as_tensor_7=True)["rigids"][..., 4:]
        for t in ts[:-1]:
            _f["t"] = t * torch.ones(B)
            rec["state"].append(_f["rigids_t"].clone()); rec["sc"].append(_f["sc_ca_t"].clone())
            out = net(_f, as_tensor_7=False)
            o7 = out["rigids"].to_tensor_7()
            rec["out"].append(o7.clone()); rec["psi"].append(out["psi"].clone())
            _f["sc_ca_t"] = o7[..., 4:]
            sc = diffuser.score(rigids_0=out["rigids"], rigids_t=Rigid.from_tensor_7(_f["rigids_t"]), t=_f["t"],
                                mask=_f["residue_mask"])
            nxt = diffuser.reverse(rigids_t=Rigid.from_tensor_7(_f["rigids_t"]), rot_score=sc["rot_score"],
                                   trans_score=sc["trans_score"], t=_f["t"], dt=dt, diffuse_mask=diffuse_mask,
                                   center_trans=True, noise_scale=1.0, probability_flow=True).to_tensor_7()
            rec["nxt"].append(nxt.clone())
            _f["rigids_t"] = nxt
    npz(name, ts=ts[:-1], **{k: torch.stack(v) for k, v in rec.items()},
        meta=np.array([B, L, n, n_pad, 0, seed]), final_scale=np.array(final_scale))


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else ""
    if mode == "stress":
        teacher_forced("stress_L64_n25_fs0p1.npz", 2, 64, 25, 3, 13, 0.1)
        return
    net, diffuser = build_reference()
    if mode == "traj128":
        # a chain length the tcgen05 pair kernels take (L % 128 == 0), two decoys, padded tail: 20 denoising steps
        trajectory(net, diffuser, "traj_L128_n20.npz", 2, 128, 20, 4, 0, 9)
        return
    if mode == "traj256":
        # BASELINE cfg 2's own size: L = 256, 100 denoising steps; two decoys with a padded tail
        trajectory(net, diffuser, "traj_L256_n100.npz", 2, 256, 100, 6, 0, 11)
        return
    if mode == "traj512":
        # BASELINE cfg 4's chain length (the long-chain IPA pair kernel, the unfused sequence-transformer attention): L = 512, one
        # decoy with a padded tail, 6 denoising steps
        trajectory(net, diffuser, "traj_L512_n6.npz", 1, 512, 6, 7, 0, 23)
        return
    if mode == "prior":
        # FrameDiffuser.sample_prior (backward_only: true) with the draws it consumes captured: randn [B,L,3] (axis),
        # rand [B,L] (CPU generator, so3.py:262), randn [B,L,3] (translation, r3.py:37-38)
        B, L = 3, 20
        torch.manual_seed(31)
        ax, u, zt = torch.randn(B, L, 3), torch.rand(B, L), torch.randn(B, L, 3)
        torch.manual_seed(31)
        pr = diffuser.sample_prior((B, L), torch.device("cpu"), as_tensor_7=True)["rigids_t"]
        npz("sample_prior.npz", axis=ax, u=u, z=zt, rigids_t=pr, cdf_row_999=diffuser.rot_diffuser._cdf[999].numpy())
        return
    if mode == "ragged":
        # chain lengths that are not a multiple of any tile size (the library pads them internally)
        trajectory(net, diffuser, "traj_L57_n8.npz", 2, 57, 8, 3, 1, 15)
        trajectory(net, diffuser, "traj_L100_n6.npz", 2, 100, 6, 0, 0, 17)
        trajectory(net, diffuser, "traj_L250_n4.npz", 1, 250, 4, 5, 0, 19)
        return

    # --- A. one network forward, small, with padding / fixed residues / chain break / mixed aatype --------
    B, L = 2, 12
    feats = synthetic.make_features(B, L, seed=3, n_pad=2, n_fixed=1, random_aatype=True)
    g = torch.Generator().manual_seed(11)
    q, x = synthetic.make_backbone(L, seed=3)
    rig = torch.cat([q, x], -1)[None].repeat(B, 1, 1)
    rig = rig + 0.3 * torch.randn(rig.shape, generator=g)  # un-normalised quats on purpose (ipa.py:337)
    feats["rigids_t"] = rig.float()
    feats["sc_ca_t"] = (x[None] + 1.5 * torch.randn(B, L, 3, generator=g)).float()
    feats["t"] = torch.tensor([0.3, 0.7])
    with torch.no_grad():
        node, edge = net.embedder(residue_idx=feats["residue_idx"], t=feats["t"],
                                  fixed_mask=feats["fixed_mask"].float(), self_conditioning_ca=feats["sc_ca_t"])
        out = net(feats, as_tensor_7=True)
        # block-0 IPA on the embedder output (module-level check)
        nm = feats["residue_mask"].float()
        node_m, edge_m = node * nm[..., None], edge * (nm[..., None] * nm[..., None, :])[..., None]
        r0 = Rigid.from_tensor_7(rig.clone().float()).apply_trans_fn(lambda v: v * 0.1)
        ipa0 = net.translator.trunk["ipa_0"](node_m, edge_m, r0, nm)
        et0 = net.translator.trunk["edge_transition_0"](node_m, edge_m)
    npz("net_forward_small.npz", rigids_t=feats["rigids_t"], sc_ca_t=feats["sc_ca_t"], t=feats["t"],
        out_rigids=out["rigids"], out_psi=out["psi"], out_atom37=out["atom37"][..., :5, :],
        out_atom14=out["atom14"][..., :5, :], node_embed=node, edge_embed=edge, ipa0=ipa0, et0=et0)

    # --- B. diffuser: score / reverse (ODE and SDE) / forward_marginal ---------------------------------
    B, L = 3, 10
    g = torch.Generator().manual_seed(21)
    t = torch.tensor([0.05, 0.5, 0.95])
    r0 = torch.cat([torch.nn.functional.normalize(torch.randn(B, L, 4, generator=g), dim=-1),
                    5 * torch.randn(B, L, 3, generator=g)], -1)
    rt = torch.cat([torch.nn.functional.normalize(r0[..., :4] + 0.3 * torch.randn(B, L, 4, generator=g), dim=-1),
                    r0[..., 4:] + 2 * torch.randn(B, L, 3, generator=g)], -1)
    mask = torch.ones(B, L, dtype=torch.float64)
    mask[:, -2:] = 0
    fixed = torch.zeros(B, L, dtype=torch.float64)
    fixed[:, 0] = 1
    diffuse = (1 - fixed) * mask
    sc = diffuser.score(rigids_0=Rigid.from_tensor_7(r0, normalize_quats=True), rigids_t=Rigid.from_tensor_7(rt),
                        t=t, mask=mask)
    ode = diffuser.reverse(Rigid.from_tensor_7(rt), sc["rot_score"], sc["trans_score"], t, 0.02, diffuse_mask=diffuse,
                           center_trans=True, noise_scale=1.0, probability_flow=True).to_tensor_7()
    torch.manual_seed(5)
    z_rot = torch.randn_like(sc["rot_score"])
    z_tr = torch.randn_like(sc["trans_score"])
    torch.manual_seed(5)
    sde = diffuser.reverse(Rigid.from_tensor_7(rt), sc["rot_score"], sc["trans_score"], t, 0.02, diffuse_mask=diffuse,
                           center_trans=True, noise_scale=0.7, probability_flow=False).to_tensor_7()
    torch.manual_seed(9)
    ax = torch.randn(B, L, 3)
    u = torch.rand(B, L)
    zt = torch.randn(B, L, 3)
    torch.manual_seed(9)
    from oracle.str2str_oracle import quat_to_rotmat

    fm = diffuser.forward_marginal(Rigid(Rotation(rot_mats=quat_to_rotmat(r0[..., :4])), r0[..., 4:]), t,
                                   diffuse_mask=diffuse, as_tensor_7=True)["rigids_t"]
    sig = diffuser.rot_diffuser.t_to_idx(torch.linspace(0.01, 1.0, 397))
    npz("diffuser_steps.npz", t=t, r0=r0, rt=rt, mask=mask, fixed=fixed, rot_score=sc["rot_score"],
        trans_score=sc["trans_score"], ode=ode, sde=sde, z_rot=z_rot, z_tr=z_tr, fm_axis=ax, fm_u=u, fm_z=zt, fm=fm,
        sigma_idx_grid=sig, cdf_row_500=diffuser.rot_diffuser._cdf[500].numpy())

    # --- C. cfg 1 of BASELINE.json: L=64, B=1, 10 denoise steps, full trajectory --------------------------
    trajectory(net, diffuser, "traj_cfg1_L64_n10.npz", 1, 64, 10, 0, 0, 7)
    trajectory(net, diffuser, "traj_masked_L24_n6.npz", 2, 24, 6, 3, 2, 5)
    trajectory(net, diffuser, "traj_L128_n20.npz", 2, 128, 20, 4, 0, 9)


if __name__ == "__main__":
    main()
