"""The UNMODIFIED reference on the hot path (benchmark / golden tooling only; never imported by the product).

`build_reference` instantiates the reference's own `DenoisingNet` / `FrameDiffuser` (imported through tools/refshim.py from
/root/reference or the staged baseline/_ref/) with the kwargs of configs/model/diffusion.yaml:16-58 and loads the synthetic
state dict; `forward_backward` is the ~75-line sampler closure of `DiffusionLitModule.predict_step`
(src/models/diffusion_module.py:260-334), which cannot be imported (it needs lightning / torchmetrics), restated around the
imported modules: every arithmetic operation is the reference's own code.  Optional hooks time it and bound the number of
forwards (a bounded sample of a long trajectory).
"""
import os
import sys
import tempfile
import time
from copy import deepcopy

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)
import refshim  # noqa: E402


def available() -> bool:
    return refshim.available()


def build_reference(final_scale=0.02, device="cpu"):
    import warnings

    refshim.install()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from src.models.net.denoising_ipa import DenoisingNet, EmbeddingModule
        from src.models.net.ipa import TranslationIPA
        from src.models.score.frame import FrameDiffuser
        from src.models.score.r3 import R3Diffuser
        from src.models.score.so3 import SO3Diffuser

        from str2str_b200 import synthetic

        torch.manual_seed(0)
        np.random.seed(0)
        net = DenoisingNet(
            embedder=EmbeddingModule(init_embed_size=32, node_embed_size=256, edge_embed_size=128, num_bins=22,
                                     min_bin=1e-5, max_bin=20.0, self_conditioning=True),
            translator=TranslationIPA(c_s=256, c_z=128, coordinate_scaling=0.1, no_ipa_blocks=4, skip_embed_size=64,
                                      transformer_num_heads=4, transformer_num_layers=2, c_hidden=256, no_heads=8,
                                      no_qk_points=8, no_v_points=12, dropout=0.0),
        )
        net.load_state_dict(synthetic.make_state_dict(seed=0, final_scale=final_scale), strict=True)
        net = net.to(device).eval()
        cache = os.path.join(tempfile.gettempdir(), "str2str_igso3_cache")
        diffuser = FrameDiffuser(
            trans_diffuser=R3Diffuser(min_b=0.1, max_b=20.0, coordinate_scaling=0.1),
            rot_diffuser=SO3Diffuser(num_omega=1000, num_sigma=1000, min_sigma=0.1, max_sigma=1.5,
                                     schedule="logarithmic", cache_dir=cache, use_cached_score=False),
            min_t=1e-2,
        )
    return net, diffuser


def forward_backward(net, diffuser, feats, rigids_t, t_delta, num_timesteps, min_t=0.01, noise_scale=1.0,
                     probability_flow=True, max_forwards=None, sync=None):
    """diffusion_module.py:260-334 after the perturbation.  Returns (final rigids tensor_7, psi, atom37 or None, seconds per
    forward+step list).  `max_forwards` stops after that many network forwards (bounded sample; atom37 is then None);
    `sync` (e.g. torch.cuda.synchronize) is called around every timed iteration."""
    from src.common.all_atom import compute_backbone
    from src.common.rigid_utils import Rigid

    T = t_delta
    n = int(float(num_timesteps) * T)
    dt = 1.0 / n
    ts = np.linspace(min_t, T, n)[::-1]
    B = rigids_t.shape[0]
    dev = rigids_t.device
    _feats = deepcopy(feats)
    _feats["rigids_t"] = rigids_t
    times, done = [], 0
    tick = (lambda: (sync() if sync else None, time.perf_counter())[1])
    rigids_pred, out = None, None
    with torch.no_grad():
        diffuse_mask = (1 - _feats["fixed_mask"]) * _feats["residue_mask"]
        _feats["sc_ca_t"] = torch.zeros_like(rigids_t[..., 4:])
        _feats["t"] = ts[0] * torch.ones(B, device=dev)
        t0 = tick()
        _feats["sc_ca_t"] = net(_feats, as_tensor_7=True)["rigids"][..., 4:]
        times.append(tick() - t0)
        done += 1
        for t in ts:
            if max_forwards is not None and done >= max_forwards:
                return rigids_pred, None, None, times
            t0 = tick()
            _feats["t"] = t * torch.ones(B, device=dev)
            out = net(_feats, as_tensor_7=False)
            if t == min_t:
                rigids_pred = out["rigids"]
            else:
                _feats["sc_ca_t"] = out["rigids"].to_tensor_7()[..., 4:]
                sc = diffuser.score(rigids_0=out["rigids"], rigids_t=Rigid.from_tensor_7(_feats["rigids_t"]),
                                    t=_feats["t"], mask=_feats["residue_mask"])
                rigids_pred = diffuser.reverse(rigids_t=Rigid.from_tensor_7(_feats["rigids_t"]),
                                               rot_score=sc["rot_score"], trans_score=sc["trans_score"],
                                               t=_feats["t"], dt=dt, diffuse_mask=diffuse_mask, center_trans=True,
                                               noise_scale=noise_scale, probability_flow=probability_flow)
                _feats["rigids_t"] = rigids_pred.to_tensor_7()
            times.append(tick() - t0)
            done += 1
        atom37 = compute_backbone(rigids_pred, out["psi"], aatype=_feats["aatype"])[0]
    return rigids_pred.to_tensor_7(), out["psi"], atom37, times
