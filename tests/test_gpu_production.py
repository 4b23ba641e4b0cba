"""GPU parity tests of the PRODUCTION configuration (tcgen05 pair kernels + tensor-core node track, the defaults) through
the C ABI, directly against the CPU oracle and goldens of the unmodified reference.  Run on the B200 box: pytest -m gpu

What is pinned here (each against the oracle / the reference, never against another kernel of this repo):
  * the tcgen05 EdgeTransition and edge-embedder kernels at L = 128 / 256 / 384 (row tiles inside one i row) and at chain
    lengths that are not a multiple of any tile (flattened tiles + internal padding);
  * the whole network forward at L = 24 ... 300, trajectories at ragged lengths and at BASELINE cfg 2's own size
    (256 residues x 100 denoise steps) against trajectories of the unmodified reference;
  * the stress fixture (final_scale = 0.1) step by step from the reference's own states, with the oracle's
    fp32 reordering noise floor printed beside every number;
  * sample_prior / backward_only, the device RNG (KS test of the IGSO(3) angle marginal), decoy-id keyed seeding;
  * module-level drop-ins (NodeTransition, TorsionAngleHead, BackboneUpdate, EdgeTransition forward).

Tolerances (relative L2 unless stated).  Pair track: the pair MLPs run with bf16 operands and fp32 accumulation and z is
stored in bf16 (DESIGN.md precision plan, SURVEY.md 8d "precision budget"), so against the fp32 oracle a pair-kernel output
carries the bf16 rounding of its output (1.1e-3 rms) plus that of three chained bf16-operand layers: measured 3.0 - 3.5e-3 on
B200; gate 4.5e-3 plus a max-abs bound that a misplaced row (an O(1) error) cannot pass.  EdgeTransition is additionally
compared with the oracle evaluated at the kernel's own rounding points (bf16 weights / hidden activations / n'_j, fp32
per-residue terms), gate 2.5e-3.  C-alpha of network outputs / trajectories: 1e-4 (BASELINE.json north_star).
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import str2str_oracle as O
from str2str_b200 import synthetic

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm())


def load(golden_dir, name):
    return {k: torch.as_tensor(v) for k, v in np.load(os.path.join(golden_dir, name)).items()}


def cuda(d):
    return {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in d.items()}


def make_net(params, pair_kernels=1, node_gemm=1):
    from str2str_b200.net import DenoisingNet, EmbeddingModule, TranslationIPA

    net = DenoisingNet(
        EmbeddingModule(init_embed_size=32, node_embed_size=256, edge_embed_size=128),
        TranslationIPA(c_s=256, c_z=128, coordinate_scaling=0.1, no_ipa_blocks=4, skip_embed_size=64),
        pair_kernels=pair_kernels, node_gemm=node_gemm,
    )
    net.load_state_dict(params, strict=True)
    return net.cuda().eval()


def make_diffuser(tmp="/tmp/str2str_b200_cache"):
    from str2str_b200.score import FrameDiffuser, R3Diffuser, SO3Diffuser

    return FrameDiffuser(R3Diffuser(0.1, 20.0, 0.1), SO3Diffuser(cache_dir=tmp), min_t=1e-2)


class kernel_log:
    """Names of the library kernels launched inside the block (the per-kernel event timing of s2s_profile_*)."""

    def __enter__(self):
        from str2str_b200 import _lib

        self.lib = _lib.load()
        self.lib.s2s_profile_reset()
        self.lib.s2s_profile_enable(1)
        self.names = {}
        return self

    def __exit__(self, *exc):
        torch.cuda.synchronize()
        self.lib.s2s_profile_enable(0)
        buf = C.create_string_buffer(1 << 16)
        n = self.lib.s2s_profile_list(buf, len(buf))
        assert n >= 0
        for ln in buf.value.decode().splitlines():
            name, ms, cnt = ln.split("\t")
            self.names[name] = int(cnt)
        self.lib.s2s_profile_reset()


def module_inputs(B, L, seed, n_pad):
    """Masked node / pair embeddings of realistic scale plus frames: what the trunk hands its sub-modules."""
    g = torch.Generator().manual_seed(seed)
    nm = torch.ones(B, L)
    if n_pad:
        nm[B - 1, L - n_pad:] = 0
    node = torch.randn(B, L, 256, generator=g) * nm[..., None]
    edge = (torch.randn(B, L, L, 128, generator=g) * (nm[..., None] * nm[..., None, :])[..., None]).bfloat16()
    return nm, node, edge


def edge_transition_at_kernel_rounding(p, pre, node, edge):
    """O.edge_transition (layers.py:170-185) evaluated with the roundings pair_tc3.cu documents: bf16 weights, bf16 z / n'_j / h1 /
    h2 operands, fp32 accumulation, the n'_i terms as fp32 per-residue vectors, fp32 LayerNorm, bf16 output."""
    import torch.nn.functional as F

    bf = lambda t: t.bfloat16().float()
    B, L, _ = node.shape
    n = O.lin(p, pre + "initial_embed", node)
    W1, b1 = p[pre + "trunk.0.weight"], p[pre + "trunk.0.bias"]
    W2, b2 = p[pre + "trunk.2.weight"], p[pre + "trunk.2.bias"]
    Wf, bfin = p[pre + "final_layer.weight"], p[pre + "final_layer.bias"]
    z, nj = bf(edge), bf(n)
    u = F.linear(n, W1[:, 128:256], b1)          # per-residue (i) term of layer 1, fp32
    pi = F.linear(n, Wf[:, 128:256], bfin)       # per-residue (i) term of the final layer
    h1 = F.relu(F.linear(z, bf(W1[:, :128])) + F.linear(nj, bf(W1[:, 256:]))[:, None, :, :] + u[:, :, None, :])
    h2 = F.relu(F.linear(bf(h1), bf(W2), b2))
    y = F.linear(bf(h2), bf(Wf)) + F.linear(z, bf(Wf[:, :128])) + F.linear(nj, bf(Wf[:, 256:]))[:, None, :, :] + pi[:, :, None, :]
    return bf(O.lnorm(p, pre + "layer_norm", y))


# ---- (a) the production pair kernels, directly against the oracle -------------------------------------------------------
@pytest.mark.parametrize("L", [128, 256, 384, 24, 57, 100, 160])
def test_edge_transition_tcgen05_vs_oracle(params, L):
    """s2s_edge_transition with the default (tcgen05) kernels against the fp32 oracle on the same bf16 pair tensor:
    L % 128 == 0 (one i row per tile), L % 32 == 0 (flattened tiles: 160), and lengths the library pads (24, 57, 100).
    Gates: see the module docstring (4.5e-3 against fp32, 2.5e-3 against the oracle at the kernel's rounding points, and a
    max-abs bound that a wrong u_i / p_i / n'_j row, an O(1) shift of a whole row, cannot pass)."""
    B = 2 if L <= 256 else 1
    nm, node, edge = module_inputs(B, L, 100 + L, 5 if L >= 57 else 2)
    net = make_net(params)
    eng = net.native("cuda")
    eng.reserve(B, L)
    with kernel_log() as kl:
        out = eng.edge_transition(1, node.cuda().contiguous(), edge.cuda().contiguous(), nm.cuda().contiguous())
    assert kl.names.get("edge_transition", 0) == 1 and "edge_transition_simt" not in kl.names, kl.names
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    pre = "translator.trunk.edge_transition_1."
    em = (nm[..., None] * nm[..., None, :])[..., None]
    ref = O.edge_transition(params, pre, node, edge.float()) * em
    ref_bf = edge_transition_at_kernel_rounding(params, pre, node, edge.float()) * em
    o = out.float().cpu()
    r, mx, r_bf = rel(o, ref), float((o - ref).abs().max()), rel(o, ref_bf)
    print(f"EdgeTransition tcgen05 L={L}: vs fp32 oracle rel {r:.2e}, max|d| {mx:.3f} (LayerNorm outputs, O(1)); "
          f"vs oracle at the kernel's rounding points rel {r_bf:.2e}")
    assert r < 4.5e-3 and mx < 0.08 and r_bf < 2.5e-3
    assert float(o[B - 1, -1].abs().max()) == 0.0  # masked rows are exactly zero


@pytest.mark.parametrize("L", [128, 256, 384, 24, 57, 100, 160])
def test_edge_embedder_tcgen05_vs_oracle(params, L):
    """s2s_embed with the default (tcgen05 pipeline) kernel against the oracle: chain break, fixed residues, per-decoy t."""
    B = 2 if L <= 256 else 1
    f = synthetic.make_features(B, L, seed=200 + L, n_pad=3, n_fixed=2)
    ridx = f["residue_idx"].clone()
    ridx[:, L // 2:] += 37  # chain break: offsets far outside [-L, L]
    q, x = synthetic.make_backbone(L, seed=200 + L)
    g = torch.Generator().manual_seed(L)
    sc = (x[None] + 2.0 * torch.randn(B, L, 3, generator=g)).float()
    t = torch.tensor([0.3, 0.7])[:B]
    net = make_net(params)
    eng = net.native("cuda")
    eng.reserve(B, L, ridx)
    rm = f["residue_mask"].float()
    with kernel_log() as kl:
        node, z = eng.embed(t.cuda(), ridx.cuda(), f["fixed_mask"].float().cuda(), sc.cuda(), rm.cuda())
    assert kl.names.get("edge_embed", 0) == 1 and "edge_embed_simt" not in kl.names, kl.names
    node_o, edge_o = O.embedder(params, ridx, t, f["fixed_mask"].float(), sc)
    node_o = node_o * rm[..., None]
    edge_o = edge_o * (rm[..., None] * rm[..., None, :])[..., None]
    r, mx = rel(z.float(), edge_o), float((z.float().cpu() - edge_o).abs().max())
    print(f"edge embedder tcgen05 L={L}: rel {r:.2e}, max|d| {mx:.3f}; node rel {rel(node, node_o):.2e}")
    assert rel(node, node_o) < 2e-5
    assert r < 4.5e-3 and mx < 0.08  # a wrong distogram bin or relative-position offset shifts a whole row by O(1)


@pytest.mark.parametrize("L", [24, 57, 100, 250])
def test_ipa_block_general_length_vs_oracle(params, L):
    """IPA block (tensor-core projections, fused pair kernel) at chain lengths the library pads internally."""
    from str2str_b200.rigid import Rigid

    B = 2
    nm, node, edge = module_inputs(B, L, 300 + L, 3)
    q, x = synthetic.make_backbone(L, seed=300 + L)
    g = torch.Generator().manual_seed(L)
    rig = torch.cat([q, 0.1 * x], -1)[None].repeat(B, 1, 1) + 0.05 * torch.randn(B, L, 7, generator=g)
    net = make_net(params)
    out = net.translator.trunk["ipa_1"](node.cuda(), edge.cuda(), Rigid.from_tensor_7(rig.cuda()), nm.cuda())
    ref = O.ipa(params, "translator.trunk.ipa_1.", node, edge.float(), rig[..., :4], rig[..., 4:], nm)
    valid = nm.bool()
    r = rel(out.cpu()[valid], ref[valid])
    print(f"ipa (padded internally) L={L}: rel {r:.2e}")
    assert r < 5e-3  # single-pass bf16 projections / logits / P.v by design (tools/precision_probe.py)


# ---- whole forward at arbitrary chain lengths -------------------------------------------------------------------------
@pytest.mark.parametrize("L", [24, 57, 64, 100, 250, 300])
def test_network_forward_any_length_vs_oracle(params, L):
    """DenoisingNet.forward on the production path against the oracle at lengths that are / are not tile multiples, with
    masked tail residues in the caller's batch (they stay keys of the sequence transformer with +1 bias, A.6) next to
    the library's own padding (which must not)."""
    B = 2
    f = synthetic.make_features(B, L, seed=400 + L, n_pad=3, n_fixed=1, random_aatype=True)
    q, x = synthetic.make_backbone(L, seed=400 + L)
    g = torch.Generator().manual_seed(L)
    f["rigids_t"] = (torch.cat([q, x], -1)[None].repeat(B, 1, 1) + 0.2 * torch.randn(B, L, 7, generator=g)).float()
    f["sc_ca_t"] = (x[None] + torch.randn(B, L, 3, generator=g)).float()
    f["t"] = torch.tensor([0.35, 0.6])
    net = make_net(params)
    with torch.no_grad(), kernel_log() as kl:
        out = net(cuda(f), as_tensor_7=True)
    assert kl.names.get("edge_transition", 0) == 3 and kl.names.get("edge_embed", 0) == 1 and kl.names.get("ipa_pair_attention", 0) == 4
    assert not any(k.endswith("_simt") for k in kl.names), kl.names
    ref = O.denoising_net(params, f)
    valid = f["residue_mask"].bool()
    r_ca = rel(out["rigids"].cpu()[valid][:, 4:], ref["rigids"][valid][:, 4:])
    r_q = rel(out["rigids"].cpu()[valid][:, :4], ref["rigids"][valid][:, :4])
    print(f"forward L={L}: C-alpha rel {r_ca:.2e}, quaternion rel {r_q:.2e}")
    assert r_ca < 1e-4 and r_q < 1e-4
    assert rel(out["atom37"].cpu()[valid][:, :5], ref["atom37"][valid][:, :5]) < 1e-4
    assert tuple(out["rigids"].shape) == (B, L, 7) and tuple(out["atom37"].shape) == (B, L, 37, 3)


def _run_traj(golden_dir, params, name, graph=True, gate=1e-4):
    from str2str_b200.rigid import Rigid
    from str2str_b200.sampler import ForwardBackwardSampler, InferenceConfig

    g = load(golden_dir, name)
    B, L, n, n_pad, n_fixed, seed = [int(v) for v in g["meta"]]
    feats = synthetic.make_features(1, L, seed=seed, n_pad=n_pad, n_fixed=n_fixed, random_aatype=True)
    net = make_net(params)
    smp = ForwardBackwardSampler(net, make_diffuser(), InferenceConfig(num_timesteps=2 * n, min_t=0.01), use_cuda_graph=graph)
    q, x = synthetic.make_backbone(L, seed=seed)
    r0 = Rigid.from_tensor_7(torch.cat([q, x], -1)[None].repeat(B, 1, 1).cuda(), normalize_quats=True)
    atom37, fin, psi = smp.forward_backward(cuda(feats), r0, 0.5, rigids_t=g["rigids_t"].cuda(), return_rigids=True)
    valid = synthetic.make_features(B, L, seed=seed, n_pad=n_pad, n_fixed=n_fixed)["residue_mask"].bool()
    ca, ca_ref = fin.cpu()[..., 4:][valid], g["final_rigids"][..., 4:][valid]
    r = rel(ca, ca_ref)
    print(f"{name} graph={graph}: C-alpha rel-L2 {r:.3e}, max|d| {float((ca - ca_ref).abs().max()):.3e} A, launches {smp.launches}")
    assert r < gate
    assert rel(torch.as_tensor(atom37)[valid][:, :5], g["final_atom37"][valid]) < gate
    return r


@pytest.mark.parametrize("name", ["traj_L57_n8.npz", "traj_L100_n6.npz", "traj_L250_n4.npz"])
def test_trajectory_ragged_lengths_vs_reference_golden(golden_dir, params, name):
    """Full trajectories of the UNMODIFIED reference at chain lengths 57 / 100 / 250 (padded tails, a fixed residue)."""
    _run_traj(golden_dir, params, name)


def test_trajectory_cfg2_size_vs_reference_golden(golden_dir, params):
    """BASELINE.json configs[1]'s own size: 256 residues x 100 denoise steps (two decoys, padded tail), CUDA-graph replay as in
    bench.py, against the trajectory of the UNMODIFIED reference: final C-alpha within 1e-4 relative."""
    _run_traj(golden_dir, params, "traj_L256_n100.npz")


def test_trajectory_cfg4_length_vs_reference_golden(golden_dir, params):
    """BASELINE.json configs[3]'s chain length: 512 residues (the long-chain IPA pair kernel with its ring of key blocks, the
    GEMM + softmax + GEMM sequence-transformer attention, flattened EdgeTransition tiles of 4 per i row), 6 denoise steps from the
    trajectory of the UNMODIFIED reference (tests/golden/make_golden.py traj512)."""
    _run_traj(golden_dir, params, "traj_L512_n6.npz")


# ---- (c) stress fixture, teacher-forced --------------------------------------------------------------------------------
def test_stress_fixture_teacher_forced_per_step(golden_dir):
    """final_scale = 0.1 (the reference's own fp32 noise floor exceeds 1e-4 on long trajectories there, so full
    trajectories are not a meaningful gate): every iteration is started from the REFERENCE's state (rigids_t, sc_ca_t, t) and
    the network output and the next state are compared with the reference's, one step at a time.  Beside every error the
    oracle's reordering noise floor is printed: the oracle re-run with 1e-6 relative jitter on every nn.Linear output."""
    from str2str_b200.sampler import InferenceConfig  # noqa: F401  (import check)

    g = load(golden_dir, "stress_L64_n25_fs0p1.npz")
    B, L, n, n_pad, _, seed = [int(v) for v in g["meta"]]
    params = synthetic.make_state_dict(seed=0, final_scale=float(g["final_scale"]))
    feats = synthetic.make_features(B, L, seed=seed, n_pad=n_pad, n_fixed=0, random_aatype=True)
    net = make_net(params)
    d = make_diffuser()
    valid = feats["residue_mask"].bool()
    diffuse = ((1 - feats["fixed_mask"]) * feats["residue_mask"]).float().cuda()
    rmask = feats["residue_mask"].float().cuda()
    worst_out = worst_nxt = worst_floor = 0.0
    gen = torch.Generator().manual_seed(0)
    orig_lin = O.lin
    steps = range(len(g["ts"]))
    for k in steps:
        f = dict(feats, rigids_t=g["state"][k], sc_ca_t=g["sc"][k], t=float(g["ts"][k]) * torch.ones(B))
        with torch.no_grad():
            out = net(cuda(f), as_tensor_7=True)["rigids"]
        e_out = rel(out.cpu()[valid][:, 4:], g["out"][k][valid][:, 4:])
        sf, sd, _ = d._sched(f["t"], 1.0 / n, "cuda")
        nxt = torch.empty(B, L, 7, device="cuda")
        d.score_and_reverse(out, g["state"][k].cuda().contiguous(), rmask, diffuse, sf, sd, nxt)
        e_nxt = rel(nxt.cpu()[valid][:, 4:], g["nxt"][k][valid][:, 4:])
        floor = float("nan")
        if k % 6 == 0:  # the oracle's own noise floor on this step
            def jitter_lin(p, name, x):
                y = orig_lin(p, name, x)
                return y * (1 + 1e-6 * torch.randn(y.shape, generator=gen))
            O.lin = jitter_lin
            try:
                with torch.no_grad():
                    jo = O.denoising_net(params, f)["rigids"]
            finally:
                O.lin = orig_lin
            floor = rel(jo[valid][:, 4:], g["out"][k][valid][:, 4:])
            worst_floor = max(worst_floor, floor)
        print(f"stress step {k:2d} t={float(g['ts'][k]):.3f}: net C-alpha rel {e_out:.2e}, next-state rel {e_nxt:.2e}, oracle jitter floor {floor:.2e}")
        worst_out, worst_nxt = max(worst_out, e_out), max(worst_nxt, e_nxt)
    print(f"stress fixture: worst per-step net error {worst_out:.2e}, next state {worst_nxt:.2e}, oracle noise floor {worst_floor:.2e}")
    assert worst_out < 1e-4 and worst_nxt < 1e-4


# ---- (d) prior sampling, device RNG, decoy-keyed seeding ---------------------------------------------------------------
def test_sample_prior_vs_reference_golden(golden_dir):
    """FrameDiffuser.sample_prior (frame.py:212-255, `backward_only: true`) with the reference's own draws injected."""
    g = load(golden_dir, "sample_prior.npz")
    d = make_diffuser()
    assert np.array_equal(d.rot_diffuser.cdf_row(999), g["cdf_row_999"].numpy())
    B, L = g["u"].shape
    out = d.sample_prior((B, L), torch.device("cuda"), as_tensor_7=True, noise=(g["axis"], g["u"], g["z"]))["rigids_t"].cpu()
    assert rel(out[..., 4:], g["rigids_t"][..., 4:]) < 1e-6
    # the reference's matrix -> quaternion conversion does not fix the sign (A.6): compare rotations, not quaternions
    Ra, Rb = O.quat_to_rotmat(out[..., :4].double()), O.quat_to_rotmat(g["rigids_t"][..., :4].double())
    assert float((Ra - Rb).abs().max()) < 5e-6


def test_backward_only_sampler_runs_from_prior(params):
    """`inference.backward_only: true`: trajectories start from the prior at T = 1 (diffusion_module.py:262-263) and run
    num_timesteps iterations; seeded draws make the run reproducible and independent of batching."""
    from str2str_b200.rigid import Rigid
    from str2str_b200.sampler import ForwardBackwardSampler, InferenceConfig

    L = 40
    feats = cuda(synthetic.make_features(1, L, seed=3))
    q, x = synthetic.make_backbone(L, seed=3)
    net = make_net(params)
    smp = ForwardBackwardSampler(net, make_diffuser(), InferenceConfig(num_timesteps=6, min_t=0.01, backward_only=True))
    r0 = lambda B: Rigid.from_tensor_7(torch.cat([q, x], -1)[None].repeat(B, 1, 1).cuda(), normalize_quats=True)
    a = smp.forward_backward(feats, r0(3), 0.5, seed=11, first_decoy=0)
    b = smp.forward_backward(feats, r0(2), 0.5, seed=11, first_decoy=1)  # decoys 1, 2 of the same job, different batch
    assert a.shape == (3, L, 37, 3) and np.isfinite(a).all()
    assert np.abs(a[1:, :, 1] - b[:, :, 1]).max() < 1e-3 * np.abs(a[:, :, 1]).max()
    assert np.abs(a[0, :, 1] - a[1, :, 1]).max() > 1.0  # different decoys really differ


def test_device_rng_igso3_angle_marginal_ks():
    """The device perturbation (s2s_rng_fill -> s2s_se3_perturb) samples the IGSO(3) rotation angle by inverting the same CDF
    row the reference interpolates (so3.py:262-268): Kolmogorov-Smirnov test of the angle of R_t R_0^T against that row,
    uniformity of the axis, and the Gaussian moments of the translation noise."""
    from scipy import stats

    from str2str_b200.rigid import Rigid

    d = make_diffuser()
    B, L, t = 8, 512, 0.6
    eye = Rigid.from_tensor_4x4(torch.eye(4, device="cuda").expand(B, L, 4, 4))
    out = d.forward_marginal(eye, t * torch.ones(B), None, as_tensor_7=True, seed=2024, first_decoy=0)["rigids_t"].cpu().double()
    qn = out[..., :4] / out[..., :4].norm(dim=-1, keepdim=True)
    omega = (2 * torch.atan2(qn[..., 1:].norm(dim=-1), qn[..., 0].abs())).reshape(-1).numpy()
    idx = int(d.rot_diffuser.t_to_idx(torch.tensor([t]))[0])
    grid, cdf = d.rot_diffuser.discrete_omega.double().numpy(), d.rot_diffuser.cdf_row(idx)
    ks = stats.kstest(omega, lambda w: np.interp(w, grid, cdf))
    print(f"IGSO3 angle KS: D = {ks.statistic:.4f}, p = {ks.pvalue:.3f} over {omega.size} draws (sigma bucket {idx})")
    assert ks.pvalue > 1e-3
    axis = (qn[..., 1:] * torch.sign(qn[..., :1])).reshape(-1, 3)
    axis = axis / axis.norm(dim=-1, keepdim=True)
    assert float(axis.mean(0).abs().max()) < 0.03                      # isotropic
    x = out[..., 4:].reshape(-1).numpy() * 0.1                         # x_t = sqrt(1 - e^-beta) z in scaled units
    var = float(1 - np.exp(-(0.1 * t + 0.5 * 19.9 * t * t)))
    assert abs(x.mean()) < 0.02 and abs(x.var() / var - 1) < 0.05
    assert stats.kstest(x / np.sqrt(var), "norm").pvalue > 1e-3


def test_decoy_keyed_draws_do_not_depend_on_batching():
    """draws are Philox(seed, subsequence = global decoy id): a decoy's perturbation is bit-identical whether it is
    sampled in one batch of 6 or as part of a 2-decoy shard starting at decoy 4."""
    d = make_diffuser()
    full = d.decoy_noise((6, 50, 3), "cuda", 7, 0, 0)
    part = d.decoy_noise((2, 50, 3), "cuda", 7, 4, 0)
    assert torch.equal(full[4:], part)
    assert not torch.equal(full[0], full[1])
    assert not torch.equal(d.decoy_noise((2, 50, 3), "cuda", 7, 4, 2), part)   # another draw of the same decoys
    u = d.decoy_noise((4, 1000), "cuda", 7, 0, 1, uniform=True)
    assert float(u.min()) >= 0.0 and float(u.max()) < 1.0 and abs(float(u.mean()) - 0.5) < 0.02


# ---- engine / graph-cache hygiene -----------------------------------------------------------------------
def test_graph_cache_is_dropped_when_the_engine_is_rebuilt(params):
    """load_state_dict between two forward_backward calls of one sampler rebuilds the native engine (new weight images, new
    workspace); the captured iteration of the first engine must not be replayed on the second."""
    from str2str_b200.rigid import Rigid
    from str2str_b200.sampler import ForwardBackwardSampler, InferenceConfig

    L, B = 64, 2
    net = make_net(params)
    smp = ForwardBackwardSampler(net, make_diffuser(), InferenceConfig(num_timesteps=8, min_t=0.01), use_cuda_graph=True)
    feats = cuda(synthetic.make_features(1, L, seed=1))
    q, x = synthetic.make_backbone(L, seed=1)
    r0 = Rigid.from_tensor_7(torch.cat([q, x], -1)[None].repeat(B, 1, 1).cuda(), normalize_quats=True)
    g = torch.Generator().manual_seed(1)
    rt = (torch.cat([q, x], -1)[None].repeat(B, 1, 1) + 0.3 * torch.randn(B, L, 7, generator=g)).float().cuda()
    a1 = smp.forward_backward(feats, r0, 0.5, rigids_t=rt)
    other = synthetic.make_state_dict(seed=5, final_scale=0.02)
    net.load_state_dict(other, strict=True)
    b1 = smp.forward_backward(feats, r0, 0.5, rigids_t=rt)
    fresh = ForwardBackwardSampler(make_net(other), make_diffuser(), InferenceConfig(num_timesteps=8, min_t=0.01), use_cuda_graph=False)
    b_ref = fresh.forward_backward(feats, r0, 0.5, rigids_t=rt)
    assert np.abs(b1 - b_ref).max() < 1e-3 and np.abs(a1 - b1).max() > 1e-2
    net.load_state_dict(params, strict=True)
    a2 = smp.forward_backward(feats, r0, 0.5, rigids_t=rt)
    assert np.array_equal(a1, a2)


def test_relative_position_table_follows_the_indices(params):
    """The engine's relative-position table is re-planned from the indices of every call: a sub-module call in between (which
    re-plans nothing) and a new residue_idx tensor with a chain break must both give the oracle's pair embedding."""
    from str2str_b200.rigid import Rigid

    B, L = 1, 32
    net = make_net(params)
    f = synthetic.make_features(B, L, seed=2)
    q, x = synthetic.make_backbone(L, seed=2)
    f["rigids_t"] = torch.cat([q, x], -1)[None].float()
    f["sc_ca_t"] = x[None].float()
    f["t"] = torch.tensor([0.4])
    for shift in (0, 500, 3):
        ff = dict(f)
        ridx = f["residue_idx"].clone()
        ridx[:, L // 2:] += shift
        ff["residue_idx"] = ridx
        nm, node, edge = module_inputs(B, L, 5, 0)
        net.translator.trunk["ipa_0"](node.cuda(), edge.cuda(), Rigid.from_tensor_7(f["rigids_t"].cuda()), nm.cuda())
        with torch.no_grad():
            out = net(cuda(ff), as_tensor_7=True)["rigids"].cpu()
        ref = O.denoising_net(params, ff)["rigids"]
        assert rel(out[..., 4:], ref[..., 4:]) < 1e-4, shift


def test_wrong_dtype_is_an_error_not_garbage(params):
    from str2str_b200 import _lib

    with pytest.raises(TypeError):
        _lib.ptr(torch.zeros(4, device="cuda", dtype=torch.float64))
    with pytest.raises(TypeError):
        _lib.ptr_i64(torch.zeros(4, device="cuda", dtype=torch.int32))


# ---- module-level drop-ins ------------------------------------------------------------------------------------------------
def test_trunk_submodules_forward_vs_oracle(params):
    """NodeTransition / TorsionAngleHead / BackboneUpdate / EdgeTransition called on their own, with the reference's forward
    signatures (layers.py:138-145,170-185,199-213,232-241), against the oracle's restatement of the same layers."""
    import torch.nn.functional as F

    B, L = 2, 40
    g = torch.Generator().manual_seed(8)
    s = torch.randn(B, L, 256, generator=g)
    net = make_net(params)
    tr = net.translator.trunk
    p = params
    nt = "translator.trunk.node_transition_2."
    h = F.relu(O.lin(p, nt + "linear_2", F.relu(O.lin(p, nt + "linear_1", s))))
    ref = O.lnorm(p, nt + "ln", O.lin(p, nt + "linear_3", h) + s)
    assert rel(tr["node_transition_2"](s.cuda()), ref) < 3e-5
    tp = "translator.torsion_pred."
    u = O.lin(p, tp + "linear_final", O.lin(p, tp + "linear_2", F.relu(O.lin(p, tp + "linear_1", s))) + s)
    ref = u / torch.sqrt(torch.clamp((u ** 2).sum(-1, keepdim=True), min=1e-8))
    assert rel(net.translator.torsion_pred(s.cuda()), ref) < 1e-4
    ref = O.lin(p, "translator.trunk.bb_update_1.linear", s)
    assert rel(tr["bb_update_1"](s.cuda()), ref) < 2e-6
    edge = torch.randn(B, L, L, 128, generator=g)
    out = tr["edge_transition_0"](s.cuda(), edge.cuda())
    ref = O.edge_transition(p, "translator.trunk.edge_transition_0.", s, edge.bfloat16().float())
    assert out.dtype == edge.dtype and rel(out, ref) < 4.5e-3


# ---- the embedding table (pair_tc4.cu MODE 2) ------------------------------------------------------------------------------
@pytest.mark.parametrize("L,case", [(128, "plain"), (256, "plain"), (256, "fixed+break"), (160, "fixed+break"), (128, "three-valued fixed mask"),
                                    (256, "float residue mask")])
def test_edge_embed_table_is_bit_identical_to_direct(params, L, case):
    """The edge embedder's table mode (per decoy: (fixed_i, fixed_j, index offset, distogram bin) -> embedding row, then a copy
    per pair) must give the direct kernel's pair tensor bit for bit; inputs the table cannot represent (a third fixed value, a
    non-binary residue mask) must fall back to the direct kernel on the device."""
    B = 3
    f = synthetic.make_features(B, L, seed=500 + L, n_pad=4, n_fixed=0)
    ridx, fixed, rm = f["residue_idx"].clone(), f["fixed_mask"].float().clone(), f["residue_mask"].float().clone()
    if case == "fixed+break":
        ridx[:, L // 3:] += 21
        fixed[0, 3:9] = 1.0
        fixed[2, L - 20:] = 1.0
    if case == "three-valued fixed mask":
        fixed[1, 5] = 1.0
        fixed[1, 6] = 0.5
    if case == "float residue mask":
        rm[0, 7] = 0.25
    q, x = synthetic.make_backbone(L, seed=500 + L)
    g = torch.Generator().manual_seed(L)
    sc = (x[None] + 3.0 * torch.randn(B, L, 3, generator=g)).float()
    t = torch.tensor([0.2, 0.5, 0.9])
    net = make_net(params)
    eng = net.native("cuda")
    eng.reserve(B, L, ridx)
    outs = []
    for table in (0, 1):
        eng.set_option("embed_table", table)
        node, z = eng.embed(t.cuda(), ridx.cuda(), fixed.cuda(), sc.cuda(), rm.cuda())
        outs.append(z.clone())
    assert torch.equal(outs[0], outs[1]), f"table mode differs from the direct kernel ({case}, L={L})"
    _, edge_o = O.embedder(params, ridx, t, fixed, sc)
    edge_o = edge_o * (rm[..., None] * rm[..., None, :])[..., None]
    assert rel(outs[1].float(), edge_o) < 4.5e-3


# ---- EdgeTransition on CTA pairs (pair_tc5.cu, cta_group::2) -----------------------------------------------------------------
@pytest.mark.parametrize("L,B", [(128, 2), (256, 2), (160, 3), (57, 2), (384, 1), (256, 7)])
def test_edge_transition_cta_pair_kernel(params, L, B):
    """The CTA-pair EdgeTransition kernel (M = 256 MMAs issued once per pair of SMs, half of the weight stream per SM) against
    the oracle and against the single-CTA kernel: same arithmetic per row, so the two kernels must agree bit for bit."""
    nm, node, edge = module_inputs(B, L, 700 + L, 4)
    net = make_net(params)
    eng = net.native("cuda")
    eng.reserve(B, L)
    outs = []
    for pair in (0, 1):
        eng.set_option("et_pair", pair)
        outs.append(eng.edge_transition(2, node.cuda().contiguous(), edge.cuda().contiguous(), nm.cuda().contiguous()).clone())
    torch.cuda.synchronize()
    pre = "translator.trunk.edge_transition_2."
    ref = O.edge_transition(params, pre, node, edge.float()) * (nm[..., None] * nm[..., None, :])[..., None]
    r = rel(outs[1].float(), ref)
    print(f"EdgeTransition CTA-pair kernel L={L} B={B}: vs fp32 oracle rel {r:.2e}; equal to the single-CTA kernel: {torch.equal(outs[0], outs[1])}")
    assert r < 4.5e-3
    assert torch.equal(outs[0], outs[1])


# ---- the caller of the path: continuous batching of (delta, replica) trajectories ------------------------------------------
@pytest.mark.parametrize("sde", [False, True])
def test_trajectory_scheduler_vs_oracle(params, sde):
    """TrajectoryScheduler (str2str_b200/scheduler.py): trajectories of two deltas (6 and 10 denoising steps) stream through a 2-row
    persistent batch, ODE and SDE sampler; every trajectory is compared with the CPU ORACLE's forward_backward of its delta
    (diffusion_module.py:260-334) from the same perturbed start frames — and, for the SDE sampler, the same per-step noise, which
    both sides key by (seed, job-wide decoy id, iteration)."""
    from str2str_b200.rigid import Rigid
    from str2str_b200.sampler import ForwardBackwardSampler, InferenceConfig
    from str2str_b200.scheduler import TrajectoryScheduler

    L, seed = 64, 77
    feats = synthetic.make_features(1, L, seed=21, n_pad=3, random_aatype=True)
    q, x = synthetic.make_backbone(L, seed=21)
    gt = torch.zeros(1, L, 8, 4, 4)
    gt[0, :, 0, :3, :3] = O.quat_to_rotmat(torch.nn.functional.normalize(q, dim=-1))
    gt[0, :, 0, :3, 3] = x
    gt[0, :, 0, 3, 3] = 1.0
    batch = cuda(dict(feats, rigidgroups_gt_frames=gt))
    net = make_net(params)
    cfg = InferenceConfig(num_timesteps=20, min_t=0.01, probability_flow=not sde, noise_scale=0.6)
    d = make_diffuser()
    smp = ForwardBackwardSampler(net, d, cfg, use_cuda_graph=True)
    work = [(0.3, 3), (0.5, 2)]
    gen = torch.Generator().manual_seed(9)
    start = {}
    for delta, n in work:
        r0 = Rigid.from_tensor_7(torch.cat([q, x], -1)[None].repeat(n, 1, 1).cuda(), normalize_quats=True)
        noise = (torch.randn(n, L, 3, generator=gen), torch.rand(n, L, generator=gen), torch.randn(n, L, 3, generator=gen))
        start[delta] = d.forward_marginal(r0, delta * torch.ones(n), diffuse_mask=torch.ones(n, L, dtype=torch.float64), noise=noise)["rigids_t"]
    sched = TrajectoryScheduler(smp, slots=2)
    atom37, rig = sched.run(batch, work, rigids_t=start, return_rigids=True, seed=seed)
    assert sched.iterations == 25 and sched.row_iterations == 3 * 7 + 2 * 11   # row 0: 7 + 7 + 11 phases, row 1: 7 + 11 (1 + n each)
    first = 0
    for delta, n in work:
        steps = int(20 * delta)
        noises = None
        if sde:
            noises = [(d.decoy_noise((n, L, 3), "cuda", seed, first, 16 + 2 * k).cpu(), d.decoy_noise((n, L, 3), "cuda", seed, first, 17 + 2 * k).cpu())
                      for k in range(steps)]
        fB = synthetic.make_features(n, L, seed=21, n_pad=3, random_aatype=True)
        fin, _, a37 = O.forward_backward(params, fB, start[delta].cpu(), delta, 20, noise_scale=0.6, probability_flow=not sde, noises=noises)
        valid = fB["residue_mask"].bool()
        r = rel(rig[delta].cpu()[..., 4:][valid], fin[..., 4:][valid])
        print(f"scheduler {'SDE' if sde else 'ODE'} delta={delta}: C-alpha rel vs oracle {r:.2e}")
        assert r < 1e-4
        assert np.abs(atom37[delta][valid.numpy()][:, :5] - a37.numpy()[valid.numpy()][:, :5]).max() < 5e-3
        # the same trajectories from the per-delta sampler (own batch, same decoy ids): equal up to fp32 reordering noise
        r0 = Rigid.from_tensor_7(torch.cat([q, x], -1)[None].repeat(n, 1, 1).cuda(), normalize_quats=True)
        _, rig_ref, _ = smp.forward_backward(batch, r0, delta, rigids_t=start[delta], return_rigids=True, seed=seed, first_decoy=first)
        assert rel(rig[delta][..., 4:], rig_ref[..., 4:]) < 2e-5
        first += n


# ---- node-track chains (gemm_chain.cu): row-local layers fused into one launch ----------------------------------------------
@pytest.mark.parametrize("L,B", [(64, 1), (57, 2), (128, 3), (256, 2), (300, 1)])
def test_node_track_chain_is_bit_identical_to_separate_launches(params, L, B):
    """The chained node track (sequence-transformer tails, post-transformer linear, NodeTransition, the per-residue terms of the
    EdgeTransition and the torsion head as gemm_chain launches; LayerNorm as a step epilogue; the four skip connections as one
    stacked GEMM) performs the same arithmetic in the same order as one launch per layer: outputs must agree BIT FOR BIT, and
    the chained forward is compared with the oracle as well."""
    f = synthetic.make_features(B, L, seed=900 + L, n_pad=3 if L > 8 else 0, n_fixed=1, random_aatype=True)
    q, x = synthetic.make_backbone(L, seed=900 + L)
    g = torch.Generator().manual_seed(L)
    f["rigids_t"] = (torch.cat([q, x], -1)[None].repeat(B, 1, 1) + 0.2 * torch.randn(B, L, 7, generator=g)).float()
    f["sc_ca_t"] = (x[None] + torch.randn(B, L, 3, generator=g)).float()
    f["t"] = torch.linspace(0.3, 0.7, B)
    net = make_net(params)
    outs, launches = [], []
    for chain in (0, 2):  # 2 = always (the default, 1, chains from 8192 residue rows up)
        net.set_option("chain", chain)  # every engine of the module, present and future
        with torch.no_grad(), kernel_log() as kl:
            out = net(cuda(f), as_tensor_7=True)
        outs.append((out["rigids"].clone(), out["psi"].clone()))
        launches.append(sum(kl.names.values()))
        assert sum(v for k, v in kl.names.items() if k.startswith("gemm_chain")) == (8 if chain else 0), kl.names
    print(f"chain L={L} B={B}: profiled launches {launches[0]} -> {launches[1]}")
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    ref = O.denoising_net(params, f)
    valid = f["residue_mask"].bool()
    assert rel(outs[1][0].cpu()[valid][:, 4:], ref["rigids"][valid][:, 4:]) < 1e-4


# ---- the path with its callers on both sides: PDB file -> features -> sampler -> ensemble PDB file --------------------------
def test_pdb_file_to_ensemble_pdb_pipeline_vs_oracle(params, tmp_path):
    """A PDB FILE on disk goes through `featurize_pdb` (parser + ProteinFeatureTransform + collate: SURVEY 8f rank 2), the frames it
    yields start `ForwardBackwardSampler.forward_backward` on the GPU (rank 0), the result is written by `atom37_to_pdb` with the
    protein's own aatype / chain / residue numbering (rank 1) and read back with the same parser.  Against the CPU oracle run on
    the same features and the same perturbed start: what lands in the file equals the oracle's atoms to the 3 decimals of the
    PDB format plus the parity budget, and sequence / numbering survive the round trip."""
    from str2str_b200.featurize import featurize_pdb, parse_pdb
    from str2str_b200.pdb_writer import atom37_to_pdb
    from str2str_b200.rigid import Rigid
    from str2str_b200.sampler import ForwardBackwardSampler, InferenceConfig

    L, B, n = 45, 2, 6   # a length the library pads internally (45 -> 64)
    g = torch.Generator().manual_seed(11)
    aatype = torch.randint(0, 20, (L,), generator=g)
    q, x = synthetic.make_backbone(L, seed=31)
    q = torch.nn.functional.normalize(q, dim=-1)
    psi0 = torch.nn.functional.normalize(torch.randn(L, 2, generator=g), dim=-1)
    native37, _ = O.backbone_atoms(q, x, psi0, aatype)
    src = str(tmp_path / "toy1.pdb")
    resnum = np.arange(L) + 7   # numbering that does not start at 1
    atom37_to_pdb(save_to=src, atom_positions=native37.numpy().astype(np.float32), aatype=aatype.numpy(), residue_index=resnum, overwrite=True)

    batch = featurize_pdb(src)
    assert batch["accession_code"] == ["toy1"] and tuple(batch["aatype"].shape) == (1, L)
    assert torch.equal(batch["aatype"][0], aatype) and torch.equal(batch["residue_index"][0], torch.as_tensor(resnum))
    gt = batch["rigidgroups_gt_frames"][..., 0, :, :].float()   # [1, L, 4, 4] backbone frames recovered from the file's atoms
    assert float((gt[0, :, :3, 3] - x).abs().max()) < 1e-3      # CA of the file = frame origins (PDB rounding)

    net = make_net(params)
    d = make_diffuser()
    smp = ForwardBackwardSampler(net, d, InferenceConfig(num_timesteps=2 * n, min_t=0.01), use_cuda_graph=True)
    r0 = Rigid.from_tensor_4x4(gt.repeat(B, 1, 1, 1).cuda())
    noise = (torch.randn(B, L, 3, generator=g), torch.rand(B, L, generator=g), torch.randn(B, L, 3, generator=g))
    rt = d.forward_marginal(r0, 0.5 * torch.ones(B), diffuse_mask=torch.ones(B, L, dtype=torch.float64), noise=noise)["rigids_t"]
    dev_batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
    atom37 = smp.forward_backward(dev_batch, r0, 0.5, rigids_t=rt)
    out = str(tmp_path / "ensemble.pdb")
    atom37_to_pdb(save_to=out, atom_positions=atom37, aatype=batch["aatype"][0].numpy(), chain_index=batch["chain_index"][0].numpy(),
                  residue_index=batch["residue_index"][0].numpy(), overwrite=True)
    txt = open(out).read()
    assert txt.count("MODEL") == B and txt.count("ENDMDL") == B and txt.endswith("END")

    # the oracle on the same features (repeated per decoy) from the same perturbed start
    f = {k: batch[k].repeat(B, *([1] * (batch[k].dim() - 1))) for k in ("residue_idx", "residue_mask", "fixed_mask", "aatype", "torsion_angles_sin_cos")}
    f = {k: (v.float() if v.is_floating_point() else v) for k, v in f.items()}
    fin_ref, _, a37_ref = O.forward_backward(params, f, rt.cpu().float(), 0.5, 2 * n)
    got = torch.as_tensor(np.asarray(atom37))
    assert rel(got[:, :, 1], a37_ref[:, :, 1]) < 1e-4   # C-alpha, before the file's rounding
    # model 1 of the written file, parsed again
    first = txt.split("ENDMDL")[0] + "ENDMDL\nEND"
    one = str(tmp_path / "model1.pdb")
    open(one, "w").write(first)
    back = parse_pdb(one)
    assert np.array_equal(back["aatype"], aatype.numpy()) and np.array_equal(back["residue_index"], resnum)
    err = np.abs(back["atom_positions"][:, :5] - a37_ref[0, :, :5].numpy()).max()
    print(f"PDB -> features -> sampler -> PDB: max |file - oracle| over N, CA, C, CB, O = {err:.2e} A")
    assert err < 3e-3


# ---- mixed-length jobs on one rank (BASELINE cfg 5's execution path) ------------------------------------------------------
def test_mixed_length_jobs_on_one_rank_vs_oracle(params):
    """`plan_mixed_lengths` cuts a mixed-length request into (length, batch) jobs; one sampler (one engine, one CUDA graph per shape)
    runs them back to back un-padded, as bench.py --workload cfg5 does on every rank.  Each job is compared with the CPU oracle's
    trajectory from the same perturbed start, and the first job is repeated at the end: same bits as the first time, whatever
    shapes the workspace and the graph cache have seen in between."""
    from str2str_b200.rigid import Rigid
    from str2str_b200.sampler import ForwardBackwardSampler, InferenceConfig, plan_mixed_lengths

    n = 4
    plan = plan_mixed_lengths({64: 3, 100: 2, 24: 2}, world=1, replica_per_batch=2)
    jobs = plan[0]
    assert sorted(jobs) == sorted([(64, 2), (64, 1), (100, 2), (24, 2)])
    net = make_net(params)
    d = make_diffuser()
    smp = ForwardBackwardSampler(net, d, InferenceConfig(num_timesteps=2 * n, min_t=0.01), use_cuda_graph=True)

    def run(L, b, seed):
        feats = synthetic.make_features(1, L, seed=seed, n_pad=2, random_aatype=True)
        q, x = synthetic.make_backbone(L, seed=seed)
        g = torch.Generator().manual_seed(seed)
        noise = (torch.randn(b, L, 3, generator=g), torch.rand(b, L, generator=g), torch.randn(b, L, 3, generator=g))
        rt_ref = O.forward_marginal(O.quat_to_rotmat(q[None].repeat(b, 1, 1)), x[None].repeat(b, 1, 1), 0.5 * torch.ones(b),
                                    synthetic.make_features(b, L, seed=seed, n_pad=2)["residue_mask"], *noise)
        r0 = Rigid.from_tensor_7(torch.cat([q, x], -1)[None].repeat(b, 1, 1).cuda(), normalize_quats=True)
        _, fin, _ = smp.forward_backward(cuda(feats), r0, 0.5, rigids_t=rt_ref.cuda(), return_rigids=True)
        return fin.cpu(), rt_ref

    first = None
    for k, (L, b) in enumerate(jobs + [jobs[0]]):
        seed = 50 + (k % len(jobs))
        fin, rt_ref = run(L, b, seed)
        if k == 0:
            first = fin.clone()
        if k == len(jobs):
            assert torch.equal(fin, first), "a repeated job must reproduce its first run bit for bit"
            continue
        featsB = synthetic.make_features(b, L, seed=seed, n_pad=2, random_aatype=True)
        fin_ref, _, _ = O.forward_backward(params, featsB, rt_ref, 0.5, 2 * n)
        valid = featsB["residue_mask"].bool()
        r = rel(fin[..., 4:][valid], fin_ref[..., 4:][valid])
        print(f"mixed-length job L={L} B={b}: C-alpha rel vs oracle {r:.2e}")
        assert r < 1e-4
